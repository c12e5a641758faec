#!/bin/bash
# Baseline A smoke run: the UNMODIFIED reference demo (vendored under baseline/_ref, prebuilt with
# nvcc -arch=sm_100 -std=c++17 -dc -dlto) on the reference's own workload, batch 256 and batch 1.
# Usage (from repo root):  gpurun --timeout 900 -- bash baseline/run_ref_demo.sh
OUT=$PWD/gpurun_out/ref_demo; mkdir -p "$OUT"
cd baseline/_ref || exit 1
nvidia-smi > "$OUT/nvidia-smi.txt" 2>&1
nproc > "$OUT/nproc.txt"; lscpu | grep -E "Model name|^CPU\(s\)" >> "$OUT/nproc.txt"
# seeded fixture (model.py itself is unseeded): same script, RNG fixed first
python - > "$OUT/model_py.log" 2>&1 <<'PY'
import torch
torch.manual_seed(0)
exec(open('model.py').read())
m = torch.jit.load('sample_input.pt')
x = dict(m.named_parameters())['0'].detach()
save_tensor(x[:1].contiguous(), 'sample_input_b1.pt')
print('B1 input', x[:1].shape)
PY
tail -3 "$OUT/model_py.log"
for i in 1 2 3; do
  /usr/bin/env bash -c "time ./demo traced_model.pt sample_input.pt" > "$OUT/demo_b256_run$i.log" 2>&1
  sha256sum demo.out >> "$OUT/demo_b256_run$i.log"; wc -l demo.out >> "$OUT/demo_b256_run$i.log"
  cat "$OUT/demo_b256_run$i.log"
done
head -c 700 demo.out > "$OUT/demo_b256_out_head.txt"
for i in 1 2 3; do
  /usr/bin/env bash -c "time ./demo traced_model.pt sample_input_b1.pt" > "$OUT/demo_b1_run$i.log" 2>&1
  sha256sum demo.out >> "$OUT/demo_b1_run$i.log"; wc -l demo.out >> "$OUT/demo_b1_run$i.log"
  cat "$OUT/demo_b1_run$i.log"
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv >> "$OUT/nvidia-smi.txt" 2>&1
