mkdir -p gpurun_out/r2ae
SECONDS=0
timeout 900 python -m pytest tests/test_forward_graph_gpu.py tests/test_gpu_parity.py tests/test_demo_gpu.py tests/test_fullsize_properties.py -x -q > gpurun_out/r2ae/pytest.log 2>&1; echo "pytest rc=$? wall=${SECONDS}s"
tail -25 gpurun_out/r2ae/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ae/bench_ours_n1.json 2> gpurun_out/r2ae/bench_ours_n1.err; echo "bench rc=$? wall=${SECONDS}s"
tail -5 gpurun_out/r2ae/bench_ours_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ae/bench_ours_n1.json'))
print(d['ms_per_step'], d['e2e']['value'], d['extra'].get('forward_ms'), d['extra'].get('forward_graph_ms'), d['extra'].get('demo_cpp_ms'))
PY
