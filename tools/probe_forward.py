"""Time of the quantised forward pass (float -> Fr, 8 matmuls, 7 ReLU decompositions) of the demo MLP, batch 256."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from zkdl_b200 import capi as zk, mlp

dims = mlp.demo_layer_dims()
ws, x = mlp.synthetic_mlp(dims, 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
for _ in range(3):
    P.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    P.forward(x)
e1.record(); torch.cuda.synchronize()
print(f"forward {e0.elapsed_time(e1) / 20:.3f} ms")
L = P.layers[2]
X = P.A[1]
e0.record()
for _ in range(20):
    zk.fr_matmul(X, L.W, 256, L.I, L.O)
e1.record(); torch.cuda.synchronize()
print(f"fr_matmul 256x2048x2048 {e0.elapsed_time(e1) / 20:.3f} ms")
Z = P.Z[2]
e0.record()
for _ in range(20):
    zk.relu_packed(Z)
e1.record(); torch.cuda.synchronize()
print(f"relu_packed 2^19 {e0.elapsed_time(e1) / 20:.3f} ms")
e0.record()
for _ in range(20):
    zk.fr_matmul_prepared(X, L.mm, 256)
e1.record(); torch.cuda.synchronize()
print(f"fr_matmul_prepared 256x2048x2048 (quantise + route + tensor-core product) {e0.elapsed_time(e1) / 20:.3f} ms")
