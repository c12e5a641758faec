"""Emulates every rank's share of the N-GPU piece plan on ONE GPU (no NCCL): per-rank device time of
P.prove(parts=plan[r]) for N = 2, 4, 8.  max over ranks ~ what bench.py --gpus N measures minus the proof gather.
Usage: python tools/probe_plan.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from zkdl_b200 import capi as zk, mlp, parallel

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dims = mlp.demo_layer_dims()
ws, x = mlp.synthetic_mlp(dims, 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
P.forward(x)
for s in range(3):
    P.prove(seed=s)
torch.cuda.synchronize()
shapes = [(L.I, L.O) for L in P.layers]


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"n1_ms": timeit(lambda: P.prove(seed=7))}
for world in (2, 4, 8):
    plans = parallel.partition_subtasks(shapes, P.B, world)
    per = [timeit(lambda: P.prove(seed=7, parts=plans[r])) for r in range(world)]
    out[f"n{world}_per_rank_ms"] = [round(v, 3) for v in per]
    out[f"n{world}_max_ms"] = max(per)
    out[f"n{world}_efficiency"] = out["n1_ms"] / (world * max(per))
print(json.dumps(out))
