import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk, mlp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(1)
v = torch.randint(-(1 << 15), 1 << 15, (n * n, 1), generator=g, device="cuda", dtype=torch.int32).float() / 65536.0
W = zk.float_to_fr(v, n * n, 1); W = zk.fr_elementwise(zk.OP_MONT, W, out=W)
G = zk.g1_mul(zk.to_device(mlp._generator()), zk.to_device(zk.random_vec(6, n)))
gens = zk.G1Table(G, full=True)
for r in range(3):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); zk.commit(gens, W); e1.record(); torch.cuda.synchronize()
    print(f"commit {n}x{n}: {e0.elapsed_time(e1):.2f} ms")
