"""One eager forward pass of the demo MLP inside an NVTX range, for an ncu launch list (tools/gpu_r2ad.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk, mlp
ws, x = mlp.synthetic_mlp(mlp.demo_layer_dims(), 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
for _ in range(2):
    P.forward(x)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("fwd")
P.forward(x)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
