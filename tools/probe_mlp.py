"""Timing probe for the demo MLP path (not a bench): setup / forward / prove, per phase, CUDA events."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk
from zkdl_b200 import mlp

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dims = mlp.demo_layer_dims()
ws, x = mlp.synthetic_mlp(dims, batch)
torch.cuda.synchronize()
t0 = time.time()
P = mlp.MLPProver(ws)
torch.cuda.synchronize()
print(f"setup (generators+tables+quantise+commit): {time.time()-t0:.3f} s, params {P.n_params}")
def timed(fn, name, reps=3):
    for r in range(reps):
        torch.cuda.synchronize(); l0 = zk.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1):.3f} ms device, {1e3*(time.time()-t0):.3f} ms wall, {zk.launch_count()-l0} launches")
    return out
timed(lambda: P.forward(x), "forward")
proof = timed(lambda: P.prove(seed=100, streams=1), "prove 1 stream", reps=3)
for ns in (2, 4, 8, 15):
    timed(lambda: P.prove(seed=100, streams=ns), f"prove {ns} streams", reps=3)
# per-call breakdown of one layer
L = P.layers[2]; B = P.B
import numpy as np
kb, ki, ko = mlp.ceil_log2(B), mlp.ceil_log2(L.I), mlp.ceil_log2(L.O)
u_bs, u_in, u_out = zk.random_vec(1, kb), zk.random_vec(2, ki), zk.random_vec(3, ko)
timed(lambda: zk.fr_partial_me(P.A[1], u_bs, L.I), "  X.partial_me(u_bs, I)")
timed(lambda: zk.fr_partial_me(L.W, u_out, 1), "  W.partial_me(u_out, 1)")
timed(lambda: zk.fr_me(P.Z[2], np.concatenate([u_out, u_bs])), "  Z(u)")
timed(lambda: zk.open_(L.gens, L.com_table, L.W, np.concatenate([u_out, u_in])), "  open")
timed(lambda: zk.me_open(L.gens, zk.fr_partial_me(L.W, u_in, L.gens.n), u_out), "  partial_me + me_open")
n = B * L.O; Lg = mlp.ceil_log2(n)
sign, magp, remp = P.aux[2]
ch = [zk.random_vec(20 + i, k) for i, k in enumerate((Lg + 5, Lg + 5, Lg + 4, Lg + 4, Lg, Lg, Lg))]
timed(lambda: zk.zkrelu_prove_packed(P.Z[2], sign, magp, remp, *ch), "  zkrelu_prove_packed")
timed(lambda: zk.hp_sumcheck(P.Z[2], sign, zk.random_vec(4, Lg), zk.random_vec(5, Lg)), "  hp_sumcheck")
timed(lambda: zk.commit(L.gens, L.W), "  commit 2048x2048")
timed(lambda: zk.fr_matmul(P.A[1], L.W, B, L.I, L.O), "  matmul")
timed(lambda: zk.relu_packed(P.Z[2]), "  relu_packed")
