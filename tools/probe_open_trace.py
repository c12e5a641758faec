"""One commitment opening (zkFC piece 2) of a 2048x2048 layer inside the NVTX range "timed" — for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv python tools/probe_open_trace.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from zkdl_b200 import capi as zk, mlp

dims = [(1773, 1773)]
ws, x = mlp.synthetic_mlp(dims, 256, seed=0)
P = mlp.MLPProver(ws, gen_seed=1)
P.forward(x)
mask = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for s in range(3):
    P.prove(seed=s, parts={("fc", 0): mask}, streams=1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
P.prove(seed=9, parts={("fc", 0): mask}, streams=1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
