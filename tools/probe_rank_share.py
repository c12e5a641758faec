"""One rank's share of the 8-GPU layer-parallel proof, emulated on ONE GPU (no NCCL): where does the time of a rank go?"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from zkdl_b200 import capi as zk, mlp

dims = mlp.demo_layer_dims()
ws, x = mlp.synthetic_mlp(dims, 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
P.forward(x)
for s in range(3):
    P.prove(seed=s)
torch.cuda.synchronize()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, t_issue


def step(kw):
    parts = P.prove(seed=7, **kw)
    return torch.cat([t.reshape(-1) for p in parts for t in p[2:]])


for name, kw in [
    ("fc2+relu2 streams=8 threads", dict(fc_layers=[2], relu_layers=[2])),
    ("fc2+relu2 streams=8 nothreads", dict(fc_layers=[2], relu_layers=[2], threads=False)),
    ("fc2+relu2 streams=2 threads", dict(fc_layers=[2], relu_layers=[2], streams=2)),
    ("fc2+relu2 streams=1", dict(fc_layers=[2], relu_layers=[2], streams=1)),
    ("fc2 only streams=1", dict(fc_layers=[2], relu_layers=[], streams=1)),
    ("relu2 only streams=1", dict(fc_layers=[], relu_layers=[2], streams=1)),
    ("fc2,fc3+relu2,relu3 (N=4 share)", dict(fc_layers=[2, 3], relu_layers=[2, 3])),
    ("open2 + mag2 + hp3 + sc4 (mixed parts)", dict(parts={("fc", 2): 2, ("relu", 2): 1, ("relu", 3): 4, ("fc", 4): 1})),
]:
    ms, issue = timeit(lambda: step(kw))
    print(f"{name:45s} gpu {ms:.3f} ms/step   host issue {issue:.3f} ms/step", flush=True)
