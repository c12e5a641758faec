"""One hidden layer (2048x2048, batch 256) of the demo MLP: zkReLU::prove + zkFC::prove, for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk, mlp
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dims = [(2048, 2048)]
ws, x = mlp.synthetic_mlp([(1773, 1773), (1773, 1773)], 256)
P = mlp.MLPProver(ws)
P.forward(x)
torch.cuda.synchronize()
for r in range(reps):
    torch.cuda.nvtx.range_push(f"prove{r}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = zk.launch_count(); e0.record()
    P.prove(seed=10 + r, fc_layers=[0], relu_layers=[0])
    e1.record(); torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(f"layer prove (relu+fc): {e0.elapsed_time(e1):.3f} ms, {zk.launch_count() - l0} launches")
