"""world_size-2 gloo test (CPU) of the multi-GPU host logic: layer partition + proof gather (zkdl_b200/parallel.py)."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zkdl_b200 import parallel


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fc, relu = parallel.partition_layers(8, world, rank)
    # a fake "proof": layer ids repeated, different length per rank
    flat = torch.tensor([100 * rank + i for i in fc for _ in range(3 + rank)] + [1000 + i for i in relu], dtype=torch.int32)
    got = parallel.gather_proof(flat, world, rank, "cpu")
    if rank == 0:
        q.put([g.tolist() for g in got])
    else:
        assert got is None
        q.put((fc, relu, flat.tolist()))
    dist.destroy_process_group()


def test_partition_covers_every_layer_once():
    for world in (1, 2, 3, 4, 8):
        fcs, relus = [], []
        for r in range(world):
            fc, relu = parallel.partition_layers(8, world, r)
            fcs += fc; relus += relu
        assert sorted(fcs) == list(range(8)) and sorted(relus) == list(range(7))
        sizes = [len(parallel.partition_layers(8, world, r)[0]) + len(parallel.partition_layers(8, world, r)[1]) for r in range(world)]
        assert max(sizes) - min(sizes) <= 2


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    items = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(120); assert p.exitcode == 0
    gathered = next(i for i in items if isinstance(i, list))
    other = next(i for i in items if isinstance(i, tuple))
    assert gathered[1] == other[2]                     # rank 1's proof arrived intact, unpadded
    assert gathered[0][:3] == [0, 0, 0] and 1000 in gathered[0]


def _msm_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(10, world, rank)
    vals = torch.arange(10, dtype=torch.int64)
    # stand-in group: integers under addition; "partial MSM" = sum of the slice, "g1_sum" = sum of the gathered partials
    res = parallel.msm_sharded(lambda: vals[lo:hi].sum().reshape(1, 1), lambda p: p.sum().reshape(1, 1), None, world)
    q.put((rank, lo, hi, int(res.item())))
    dist.destroy_process_group()


def test_shard_ranges_and_sharded_msm_world2_gloo():
    for n in (1, 7, 10, 1 << 20):
        for world in (1, 2, 3, 8):
            rs = [parallel.shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_msm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    items = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(120); assert p.exitcode == 0
    assert all(it[3] == 45 for it in items)          # every rank ends with the full sum
