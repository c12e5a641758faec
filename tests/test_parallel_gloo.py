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


class OracleOps:
    """Same interface as parallel.CapiOps with the CPU oracle as the local primitive (tests only)."""

    def local(self, kind, tables, u, v):
        import numpy as np
        from oracle import oracle as orc
        t = [np.asarray(x, dtype=np.uint32) for x in tables]
        if kind == "bin":
            return orc.bin_sumcheck(t[0], u, v)
        if kind == "hp":
            return orc.hp_sumcheck(t[0], t[1], u, v)
        return orc.ip_sumcheck(t[0], t[1], u)

    def eq_weights(self, u_hi, world):
        import numpy as np
        from oracle import oracle as orc
        one = orc.fr_mont(orc.to_limbs([1]))[0]
        out = []
        for r in range(world):
            e = np.zeros((world, 8), np.uint32); e[r] = one
            out.append(orc.fr_me(e, u_hi))
        return out

    def weighted_sum(self, parts, weights):
        from oracle import oracle as orc
        acc = None
        for p, w in zip(parts, weights):
            t = p if w is None else orc.fr_bcast(p, w, "mul")
            acc = t if acc is None else orc.fr_add(acc, t)
        return acc

    def stack(self, rows):
        import numpy as np
        return np.stack(list(rows))

    def cat(self, parts):
        import numpy as np
        return np.concatenate(list(parts))


def _sc_worker(rank, world, port, q):
    import numpy as np
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    k = 6; n = 1 << k
    def rnd(m):
        x = rng.integers(0, 1 << 32, size=(m, 8), dtype=np.uint64).astype(np.uint32); x[:, 7] %= 1944954707; return x
    a, b, u, v = rnd(n), rnd(n), rnd(k), rnd(k)
    lo, hi = parallel.shard_range(n, world, rank)

    def all_gather(x):
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.int32))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.numpy().view(np.uint32) for o in outs]

    ops = OracleOps()
    res = {}
    for kind, tabs in (("bin", [a[lo:hi]]), ("hp", [a[lo:hi], b[lo:hi]]), ("ip", [a[lo:hi], b[lo:hi]])):
        got = parallel.sumcheck_sharded(kind, ops, tabs, u, v, world, rank, all_gather)
        full = {"bin": lambda: orc.bin_sumcheck(a, u, v), "hp": lambda: orc.hp_sumcheck(a, b, u, v), "ip": lambda: orc.ip_sumcheck(a, b, u)}[kind]()
        res[kind] = bool(np.array_equal(got, full))
    q.put((rank, res))
    dist.destroy_process_group()


def test_sharded_sumcheck_world2_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 1000
    procs = [ctx.Process(target=_sc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    items = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(120); assert p.exitcode == 0
    for rank, res in items:
        assert res == {"bin": True, "hp": True, "ip": True}, (rank, res)


# ------------------------------------------------------------------------------------------------ sub-layer partition
DEMO_SHAPES = [(1024, 1024), (1024, 2048)] + [(2048, 2048)] * 5 + [(2048, 1024)]


def _demo_meta(B=256):
    meta, order = {}, []
    nl = len(DEMO_SHAPES)
    ngens = [1024, 2048, 2048, 2048, 2048, 2048, 2048, 2048]
    order.append(("fc", nl - 1))
    for i in range(nl - 2, -1, -1):
        order += [("relu", i), ("fc", i)]
    for i, (I, O) in enumerate(DEMO_SHAPES):
        meta[("fc", i)] = (I, ngens[i], B * O)
        meta[("relu", i)] = (I, ngens[i], B * O)
    return meta, order


def test_subtask_partition_covers_every_piece_once_and_balances():
    for world in (1, 2, 3, 4, 8):
        plan = parallel.partition_subtasks(DEMO_SHAPES, 256, world)
        assert plan == parallel.partition_subtasks(DEMO_SHAPES, 256, world)          # deterministic on every rank
        seen = {}
        for p in plan:
            for key, mask in p.items():
                assert mask and not (seen.get(key, 0) & mask)
                seen[key] = seen.get(key, 0) | mask
        assert all(seen[("fc", i)] == 3 for i in range(8)) and all(seen[("relu", i)] == 7 for i in range(7))
        assert ("relu", 7) not in seen
        loads = [sum(parallel.subtask_cost(k[0], m, *DEMO_SHAPES[k[1]], 256, world) for k, mk in p.items() for m in (1, 2, 4) if mk & m)
                 for p in plan]
        assert max(loads) <= 1.15 * (sum(loads) / world) + 1e-9


def _fake_results(plan_rank, meta, order, full):
    """What MLPProver.prove(parts=plan_rank) would return: full-size buffers with only the owned rows valid."""
    res = []
    for key in order:
        if key not in plan_rank:
            continue
        bufs = [torch.full_like(b, -1) for b in full[key]]
        for buf, lo, hi in parallel.task_segments(key[0], plan_rank[key], *meta[key]):
            bufs[buf][lo:hi] = full[key][buf][lo:hi]
        res.append((key[0], key[1]) + tuple(bufs))
    return res


def _full_proofs(meta, order):
    g = torch.Generator().manual_seed(5)
    full = {}
    for key in order:
        rows = [0, 0]
        for buf, lo, hi in parallel.task_segments(key[0], 3 if key[0] == "fc" else 7, *meta[key]):
            rows[buf] = max(rows[buf], hi)
        full[key] = [torch.randint(0, 2 ** 31 - 1, (rows[b], (8, 36)[b]), dtype=torch.int32, generator=g)
                     for b in range(2 if key[0] == "fc" else 1)]
    return full


def test_pack_assemble_roundtrip_local():
    meta, order = _demo_meta()
    full = _full_proofs(meta, order)
    for world in (1, 3, 8):
        plans = parallel.partition_subtasks(DEMO_SHAPES, 256, world)
        flats = [parallel.pack_owned(_fake_results(plans[r], meta, order, full), plans[r], meta) for r in range(world)]
        got = parallel.assemble(flats, plans, meta, order)
        for key in order:
            for a, b in zip(got[key], full[key]):
                assert torch.equal(a, b)


def _subtask_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meta, order = _demo_meta()
    full = _full_proofs(meta, order)                                   # same seed on every rank
    plans = parallel.partition_subtasks(DEMO_SHAPES, 256, world)
    flat = parallel.pack_owned(_fake_results(plans[rank], meta, order, full), plans[rank], meta)
    got = parallel.gather_proof(flat, world, rank, "cpu")
    ok = True
    if rank == 0:
        asm = parallel.assemble(got, plans, meta, order)
        ok = all(torch.equal(a, b) for key in order for a, b in zip(asm[key], full[key]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_subtask_gather_assemble_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 1000
    procs = [ctx.Process(target=_subtask_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    items = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(120); assert p.exitcode == 0
    assert all(ok for _, ok in items)


# ------------------------------------------------------------------------------------------------ commit sharded by row
def _commit_worker(rank, world, port, q):
    """Commitment::commit by row ranges with the CPU oracle as the local committer, default
    torch.distributed all-gather (gloo), m = 5 rows over 2 ranks (uneven shards: 3 + 2, padded gather)."""
    import numpy as np
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    ng, m = 4, 5
    G = orc.g1_mul(orc.g1_generator(), orc.to_limbs([int(v) for v in rng.integers(1, 1 << 60, size=ng)]), fast=True)
    W = orc.fr_from_ints([int(v) for v in rng.integers(-3000, 3000, size=ng * m)], mont=True)

    def commit(gens, t):                         # torch int32 limbs in/out, like capi.commit
        out = orc.commit(gens, t.numpy().view(np.uint32), fast=True)
        return torch.from_numpy(np.ascontiguousarray(out).view(np.int32)).reshape(-1, 36)

    Wt = torch.from_numpy(W.view(np.int32))
    got = parallel.commit_sharded(commit, G, Wt, ng, world, rank).numpy().view(np.uint32)
    full = orc.commit(G, W, fast=True)
    q.put((rank, got.shape[0] == m and bool(orc.g1_eq(got, full).all())))
    dist.destroy_process_group()


def test_commit_sharded_world2_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 1000
    procs = [ctx.Process(target=_commit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    items = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(120); assert p.exitcode == 0
    assert all(ok for _, ok in items), items
