"""Multi-GPU plumbing of the layer-parallel prover: which rank proves which layer, and how the proof
elements reach rank 0.  There is no data-path collective: every layer's zkFC / zkReLU proof is independent once the
forward pass has run, so ranks only exchange finished proof elements (a padded gather, ~100 KB).
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def partition_layers(n_layers, world, rank):
    """Layers (fc i, relu i) owned by `rank`: round-robin, so that the 8 fc + 7 relu proofs of the demo MLP spread
    evenly.  Returns (fc_layers, relu_layers)."""
    fc = [i for i in range(n_layers) if i % world == rank]
    relu = [i for i in range(n_layers - 1) if i % world == rank]
    return fc, relu


def gather_proof(flat, world, rank, device, sizes=None):
    """Gathers each rank's flat proof tensor on rank 0.  Returns the list of per-rank tensors on rank 0, None elsewhere.
    `sizes` (elements per rank) is known in advance for a given model shape: passing it avoids the size exchange and
    its host synchronisation."""
    if world == 1:
        return [flat]
    if sizes is None:
        szs = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(szs, torch.tensor([flat.numel()], dtype=torch.int64, device=device))
        sizes = [int(s.item()) for s in szs]
    mx = max(sizes)
    if flat.numel() == mx:
        buf = flat
    else:
        buf = torch.zeros(mx, dtype=flat.dtype, device=device)
        buf[: flat.numel()] = flat
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank != 0:
        return None
    return [o[:s] for o, s in zip(outs, sizes)]


def shard_range(n, world, rank):
    """Contiguous point range [lo, hi) of rank `rank` for an MSM over n (base, scalar) pairs."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def msm_sharded(msm_local, g1_sum, part_like, world):
    """Single large MSM partitioned by point range: every rank computes the MSM of its own (bases, scalars) slice
    (`msm_local()` -> one Jacobian point, [1, 36] limbs), the P partial points (144 B each) are all-gathered and every
    rank adds them locally with `g1_sum` (G1 addition is not an NCCL reduction op).  Returns the full MSM on every rank."""
    part = msm_local()
    if world == 1:
        return part
    parts = [torch.empty_like(part_like if part_like is not None else part) for _ in range(world)]
    dist.all_gather(parts, part)
    return g1_sum(torch.cat(parts))
