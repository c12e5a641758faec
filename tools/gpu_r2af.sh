mkdir -p gpurun_out/r2af
SECONDS=0
timeout 600 python -m pytest tests/test_forward_graph_gpu.py tests/test_gpu_parity.py -x -q -k "forward or fused or matmul" > gpurun_out/r2af/pytest.log 2>&1; echo "pytest rc=$? wall=${SECONDS}s"
tail -12 gpurun_out/r2af/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/r2af/forward_launches.csv python tools/probe_forward.py > gpurun_out/r2af/ncu.log 2>&1; echo "ncu rc=$? wall=${SECONDS}s"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2af/bench_quick.json 2> gpurun_out/r2af/bench_quick.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2af/bench_quick.json'))
print(d['ms_per_step'], d['e2e']['value'], d['extra'].get('forward_ms'), d['extra'].get('forward_graph_ms'))
PY
