"""The quantised forward pass of the demo MLP at batch 256 (float -> Fr, 8 products, 7 fused zkReLU epilogues): eager and
graph-replay times, and one eager pass inside the NVTX range "fwd" for ncu (launch list: tools/gpu_r2ad.sh; full capture
of the tcgen05 kernel: tools/gpu_r2ag.sh)."""
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
if os.environ.get("PROBE_TIMES"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("eager", lambda: P.forward(x)), ("graph", lambda: P.forward(x, graph=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"forward ({name}) {e0.elapsed_time(e1) / 20:.3f} ms")
