"""8-layer batch-256 proof time at a given stream count (tuning probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk, mlp
ws, x = mlp.synthetic_mlp(mlp.demo_layer_dims(), 256)
P = mlp.MLPProver(ws); P.forward(x)
import time
for ns in [int(v) for v in sys.argv[1:]] or [1, 8]:
  for th in ((False, True) if ns > 1 else (False,)):
    for r in range(3): P.prove(seed=r, streams=ns, threads=th)
    torch.cuda.synchronize(); t0 = time.time(); P.prove(seed=5, streams=ns, threads=th); t_issue = time.time() - t0; torch.cuda.synchronize()
    print(f"streams={ns} threads={th}: host issue time {1e3*t_issue:.2f} ms")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(10): P.prove(seed=10 + r, streams=ns, threads=th)
    e1.record(); torch.cuda.synchronize()
    print(f"streams={ns} threads={th}: {e0.elapsed_time(e1)/10:.3f} ms per proof (ZKDL_MSM_C={os.environ.get('ZKDL_MSM_C','auto')})")
