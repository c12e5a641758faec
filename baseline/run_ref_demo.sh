#!/bin/bash
# Baseline A smoke run: the UNMODIFIED reference demo (oracle/_ref/demo, built from /root/reference by oracle/build_ref.sh
# with nvcc -arch=sm_100 -std=c++17 -dc -dlto) on the reference's own workload, batch 256 and batch 1.
# Usage (from repo root):  gpurun --timeout 900 -- bash baseline/run_ref_demo.sh
# (bench.py --impl reference is the maintained way to time the reference; this script is kept for manual runs.)
OUT=$PWD/gpurun_out/ref_demo; mkdir -p "$OUT"
DEMO=$PWD/oracle/_ref/demo
WORK=$(mktemp -d); cd "$WORK" || exit 1
nvidia-smi > "$OUT/nvidia-smi.txt" 2>&1
python - > "$OUT/model_py.log" 2>&1 <<'PY'
import torch, torch.nn as nn
torch.manual_seed(0)
def save_tensor(t, fn):
    m = nn.Module(); m.register_parameter("0", nn.Parameter(t)); torch.jit.script(m).save(fn)
d = [784, 1000, 1773, 1773, 1773, 1773, 1773, 1124, 1000]
layers = []
for i in range(8):
    layers.append(nn.Linear(d[i], d[i + 1], bias=False))
    if i < 7: layers.append(nn.ReLU())
model = nn.Sequential(*layers).to("cuda").eval()
x = torch.randn(256, 784).to("cuda")
save_tensor(x, "sample_input.pt"); save_tensor(x[:1].contiguous(), "sample_input_b1.pt")
torch.jit.trace(model, x[:1]).save("traced_model.pt")
PY
for inp in sample_input.pt sample_input_b1.pt; do
  for i in 1 2; do
    /usr/bin/env bash -c "time $DEMO traced_model.pt $inp" > "$OUT/demo_${inp%.pt}_run$i.log" 2>&1
    sha256sum demo.out >> "$OUT/demo_${inp%.pt}_run$i.log"; cat "$OUT/demo_${inp%.pt}_run$i.log"
  done
done
