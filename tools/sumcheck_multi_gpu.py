"""Sumcheck partitioned by the leading hypercube variables across GPUs:
   torchrun --nproc-per-node P tools/sumcheck_multi_gpu.py [log_n]
Rank 0 also runs the single-GPU sumcheck on the whole table and checks bit-equality of the proofs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from zkdl_b200 import capi as zk, parallel

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/zkdl_nccl_%h_%p.log")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
lo, hi = parallel.shard_range(n, world, rank)
g = torch.Generator(device="cuda").manual_seed(11)                      # same table on every rank, each keeps its slice
def table():
    t = torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g); t[:, 7] &= 0x3FFFFFFF
    return t
A, Bt = table(), table()
a, b = A[lo:hi].contiguous(), Bt[lo:hi].contiguous()
u, v = zk.random_vec(1, k), zk.random_vec(2, k)
ops = parallel.CapiOps()
def all_gather(x):
    outs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(outs, x.contiguous())
    return outs
out = {}
for kind, tabs, fulltabs in (("bin", [a], [A]), ("hp", [a, b], [A, Bt]), ("ip", [a, b], [A, Bt])):
    fn = lambda: parallel.sumcheck_sharded(kind, ops, tabs, u, v, world, rank, all_gather)
    for _ in range(2): res = fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): res = fn()
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank == 0:
        full = ops.local(kind, fulltabs, u, v)
        out[kind] = {"ms": ms, "bit_identical_to_single_gpu": bool(torch.equal(res, full))}
if rank == 0:
    print(json.dumps({"config": f"sumchecks on 2^{k} entries split by leading variables", "n_gpus": world, **out}))
if world > 1: dist.destroy_process_group()
