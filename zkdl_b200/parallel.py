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


# ------------------------------------------------------------------------------------------------ sub-layer partition
# A layer's proof is itself several independent pieces (include/zkdl_b200.h, ZKDL_FC_* / ZKDL_RELU_*): zkFC = the matmul
# sumcheck + the commitment opening; zkReLU = the binary sumcheck of mag_bin, the one of rem_bin, the Hadamard sumcheck.
# Partitioning these 37 pieces of the demo MLP (instead of 15 whole layers) lets 8 ranks balance to within a piece.
FC_MASKS = (1, 2)
RELU_MASKS = (1, 2, 4)


def _clog2(n):
    return 0 if n <= 1 else (int(n) - 1).bit_length()


def subtask_cost(kind, mask, I, O, B, world=8):
    """Cost (ms) of one piece for the longest-first plan.  With many ranks a rank holds a handful of pieces and is bound by
    their LATENCY (the piece alone on an idle B200, tools/probe_subtasks.py); with two ranks each GPU is as busy as a
    single one and is bound by the pieces' THROUGHPUT share (an opening is mostly bucket accumulation on every SM, a
    sumcheck mostly short launches).  The cost blends the two by world size.  Fitted on the demo shapes with the round-2
    build (profiles/r2_subtask_latencies_b200.json)."""
    n = B * O / 2 ** 20
    if kind == "fc":
        ngens = 1 << ((_clog2(I * O) + 1) // 2)                              # demo.cu:81 on the padded shape
        lat = 0.134 + 0.019 * (I * O / 2 ** 20) if mask == 1 else 0.37 + 0.335 * (ngens / 1024)
        thr = 0.03 + 0.01 * (I * O / 2 ** 20) if mask == 1 else 0.08 + 0.33 * (ngens / 1024)
    else:
        lat = {1: 0.35 + 0.68 * n, 2: 0.31 + 0.56 * n, 4: 0.222 + 0.232 * n}[mask]
        thr = {1: 0.46 * n, 2: 0.33 * n, 4: 0.25 * n}[mask]
    w = min(1.0, max(0.0, (world - 1) / 7.0))
    return w * lat + (1.0 - w) * thr


def partition_subtasks(shapes, B, world, costs=None):
    """shapes: [(I, O)] padded layer shapes.  Returns one dict per rank {("fc"|"relu", layer): part mask}.
    Longest-processing-time greedy over the pieces; among equally loaded ranks a piece joins the rank that already owns
    another piece of the same layer proof (they share a table).  Deterministic: every rank computes the same plan."""
    pieces = []
    for i, (I, O) in enumerate(shapes):
        for m in FC_MASKS:
            pieces.append((("fc", i), m, I, O))
        if i + 1 < len(shapes):
            for m in RELU_MASKS:
                pieces.append((("relu", i), m, I, O))
    cost = lambda pc: (costs or {}).get((pc[0], pc[1]), subtask_cost(pc[0][0], pc[1], pc[2], pc[3], B, world))
    pieces.sort(key=lambda pc: (-cost(pc), pc[0][1], pc[0][0], pc[1]))
    load = [0.0] * world
    plan = [dict() for _ in range(world)]
    for pc in pieces:
        r = min(range(world), key=lambda q: (round(load[q], 6), 0 if pc[0] in plan[q] else 1, q))
        plan[r][pc[0]] = plan[r].get(pc[0], 0) | pc[1]
        load[r] += cost(pc)
    return plan


def task_segments(kind, mask, I, ngens, n):
    """[(buf, lo, hi)]: rows of the proof buffers a part mask writes.  buf 0 = Fr rows (8 limbs), buf 1 = G1 rows (36)."""
    segs = []
    if kind == "fc":
        nip = 3 * _clog2(I) + 2
        if mask & 1:
            segs.append((0, 0, nip + 1))
        if mask & 2:
            segs += [(0, nip + 1, nip + 2), (1, 0, 3 * _clog2(ngens) + 2)]
        return segs
    L = _clog2(n)
    a = 3 * (L + 5) + 1 + 32
    b = a + 3 * (L + 4) + 1 + 16
    for bit, lo, hi in ((1, 0, a), (2, a, b), (4, b, b + 3 * L + 2)):
        if mask & bit:
            segs.append((0, lo, hi))
    return segs


def pack_owned(results, plan_rank, meta):
    """results: [(kind, layer, proof_fr[, proof_g1])] as MLPProver.prove(parts=plan_rank) returns them (full-size
    buffers, only the owned segments written).  Returns the owned segments as one flat tensor.
    meta[(kind, layer)] = (I, ngens, n)."""
    out = []
    for res in results:
        key = (res[0], res[1])
        for buf, lo, hi in task_segments(key[0], plan_rank[key], *meta[key]):
            out.append(res[2 + buf][lo:hi].reshape(-1))
    return torch.cat(out)


def assemble(flats, plans, meta, order):
    """Inverse of pack_owned over all ranks: flats[r] = rank r's flat tensor, plans = partition_subtasks(...),
    order = [(kind, layer)] in proving order.  Returns {(kind, layer): [proof_fr] or [proof_fr, proof_g1]}."""
    width = (8, 36)
    full = {}
    for key in order:
        I, ngens, n = meta[key]
        every = 3 if key[0] == "fc" else 7
        rows = [0, 0]
        for buf, lo, hi in task_segments(key[0], every, I, ngens, n):
            rows[buf] = max(rows[buf], hi)
        full[key] = [flats[0].new_zeros((rows[b], width[b])) for b in range(2 if key[0] == "fc" else 1)]
    for r, plan in enumerate(plans):
        off = 0
        for key in order:
            if key not in plan:
                continue
            for buf, lo, hi in task_segments(key[0], plan[key], *meta[key]):
                cnt = (hi - lo) * width[buf]
                full[key][buf][lo:hi] = flats[r][off: off + cnt].reshape(hi - lo, width[buf])
                off += cnt
        assert off == flats[r].numel(), "flat proof of rank %d does not match its plan" % r
    return full


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


def commit_sharded(commit, gens, W, ngens, world, rank, all_gather=None):
    """Commitment::commit partitioned by row (/root/reference/commitment.cu:29-41): the m = |W| / |G| rows
    are independent commitments, so rank r computes rows [lo, hi) = shard_range(m, world, r) with the replicated generators
    and the row commitments (144 B each) are all-gathered; no reduction.  commit(gens, t) -> [rows, 36] Jacobian limbs.
    all_gather(x) -> list of every rank's x (default: torch.distributed.all_gather, padded to the largest shard)."""
    if world == 1:
        return commit(gens, W)
    m = W.shape[0] // ngens
    lo, hi = shard_range(m, world, rank)
    part = commit(gens, W[lo * ngens: hi * ngens]) if hi > lo else W.new_zeros((0, 36))
    if all_gather is not None:
        return torch.cat(list(all_gather(part)))
    mx = -(-m // world)
    buf = part if part.shape[0] == mx else torch.cat([part, part.new_zeros((mx - part.shape[0], 36))])
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf.contiguous())
    return torch.cat([o[: shard_range(m, world, r)[1] - shard_range(m, world, r)[0]] for r, o in enumerate(outs)])


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


# ------------------------------------------------------------------------------------------------ sharded sumcheck
# A 2^k table split by its LEADING log2(P) variables: rank r holds the contiguous slice [r * 2^kl, (r+1) * 2^kl),
# kl = k - log2 P.  The folds bind the least-significant variable first (fr-tensor.cu:404-406), so the first kl rounds
# are purely local.  Round j < kl of the global sumcheck is
#     S_j = sum_r eq(u[kl:], r) * S_j^(r)      (binary / Hadamard: the eq weight factorises over the rank bits)
#     S_j = sum_r S_j^(r)                      (inner product: no weights)
# where S_j^(r) is what the single-GPU routine emits on rank r's slice with challenges u[:kl], v[:kl]; the remaining
# log2 P rounds run on the P-vector of the slices' final values.  Communication: ONE all-gather of 3*kl + 1 (or 2) field
# elements per rank (all challenges are known up front), then a few dozen local modular operations.
class CapiOps:
    """The sharded-sumcheck arithmetic through libzkdl_b200 (GPU)."""

    def __init__(self):
        from . import capi
        self.zk = capi

    def local(self, kind, tables, u, v):
        zk = self.zk
        if kind == "bin":
            return zk.bin_sumcheck(tables[0], u, v)
        if kind == "hp":
            return zk.hp_sumcheck(tables[0], tables[1], u, v)
        return zk.ip_sumcheck(tables[0], tables[1], u)

    def eq_weights(self, u_hi, world):
        """eq(u_hi, r) for r < world: the multilinear extension of the r-th unit vector evaluated at u_hi."""
        import numpy as np
        zk = self.zk
        one = np.array([4294967294, 1, 215042, 1485092858, 3971764213, 2576109551, 2898593135, 405057881], dtype=np.uint32)
        out = []
        for r in range(world):
            e = np.zeros((world, 8), np.uint32); e[r] = one
            out.append(zk.to_host(zk.fr_me(zk.to_device(e), u_hi))[0])
        return out

    def weighted_sum(self, parts, weights):
        zk = self.zk
        acc = None
        for p, w in zip(parts, weights):
            t = p if w is None else zk.fr_broadcast(zk.OP_MUL, p, w)
            acc = t if acc is None else zk.fr_elementwise(zk.OP_ADD, acc, t)
        return acc

    def stack(self, rows):
        return torch.stack(list(rows))

    def cat(self, parts):
        return torch.cat(list(parts))


def sumcheck_sharded(kind, ops, tables_local, u, v, world, rank, all_gather):
    """kind in {"bin", "hp", "ip"}; tables_local: this rank's slice(s); u, v: the FULL challenge vectors (host arrays,
    v ignored for "ip"); all_gather(x) -> list of every rank's x.  Returns the proof, identical to the single-GPU one."""
    k = len(u)
    lp = (world - 1).bit_length()
    assert (1 << lp) == world and k >= lp, "world must be a power of two no larger than the table"
    kl = k - lp
    nfin = 1 if kind == "bin" else 2
    local = ops.local(kind, tables_local, u[:kl], None if kind == "ip" else v[:kl])     # [3*kl + nfin]
    if world == 1:
        return local
    parts = all_gather(local)
    weights = [None] * world if kind == "ip" else ops.eq_weights(u[kl:], world)
    rounds = ops.weighted_sum([p[: 3 * kl] for p in parts], weights)
    finals = [ops.stack([p[3 * kl + i] for p in parts]) for i in range(nfin)]             # P-vectors of a (and b)
    tail = ops.local(kind, finals, u[kl:], None if kind == "ip" else v[kl:])
    return ops.cat([rounds, tail])
