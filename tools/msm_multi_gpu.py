"""MSM partitioned by point range across GPUs: torchrun --nproc-per-node P tools/msm_multi_gpu.py [log_n]
Every rank builds the same seeded bases/scalars for its slice, computes its partial MSM, all-gathers the P partial points
and adds them.  Rank 0 checks the result against the single-GPU MSM over all points (as points, via the oracle)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from zkdl_b200 import capi as zk, mlp, parallel

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/zkdl_nccl_%h_%p.log")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = 1 << lg
lo, hi = parallel.shard_range(N, world, rank)
ks_all = zk.random_vec(7, N); sc_all = zk.random_vec(8, N)            # same seeded streams on every rank
gen = zk.to_device(mlp._generator())
G = zk.g1_mul(gen, zk.to_device(ks_all[lo:hi]))
tab = zk.G1Table(G, full=False)
sc = zk.to_device(sc_all[lo:hi])
def step():
    return parallel.msm_sharded(lambda: zk.msm(tab, sc, 1, False), zk.g1_sum, None, world)
for _ in range(2): res = step()
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): res = step()
e1.record()
if world > 1: dist.barrier()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if world > 1:
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    ok = None
    if lg <= 20:
        Gf = zk.g1_mul(gen, zk.to_device(ks_all)); tf = zk.G1Table(Gf, full=False)
        full = zk.msm(tf, zk.to_device(sc_all), 1, False)
        from oracle import oracle as orc
        ok = bool(orc.g1_eq(zk.to_host(res), zk.to_host(full)).all())
    print(json.dumps({"config": f"msm 2^{lg} by point range", "n_gpus": world, "ms": ms, "Mpts_per_s": N / ms / 1e3, "matches_single_gpu": ok}))
if world > 1: dist.destroy_process_group()
