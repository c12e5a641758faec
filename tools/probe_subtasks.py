"""Latency of every independent part of the per-layer proofs, each run ALONE on an idle GPU (cost model of
zkdl_b200.parallel.partition_subtasks).  Usage: python tools/probe_subtasks.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from zkdl_b200 import capi as zk, mlp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dims = mlp.demo_layer_dims()
ws, x = mlp.synthetic_mlp(dims, 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
P.forward(x)
P.prove(seed=1); P.prove(seed=2)
torch.cuda.synchronize()
nl = len(P.layers)
out = {}


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


for i in range(nl):
    for kind, masks in (("fc", (1, 2, 3)), ("relu", (1, 2, 4, 7))):
        if kind == "relu" and i == nl - 1:
            continue
        for m in masks:
            sub = {(kind, i): m}
            ms = timeit(lambda: P.prove(seed=5, parts=sub, streams=1))
            out[f"{kind}{i}:{m}"] = round(ms, 3)
            print(kind, i, "I,O=", P.layers[i].I, P.layers[i].O, "parts", m, f"{ms:.3f} ms", flush=True)
print(json.dumps(out))
