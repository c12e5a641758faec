mkdir -p gpurun_out/r2ah
SECONDS=0
timeout 600 python -m pytest tests/test_forward_graph_gpu.py tests/test_gpu_parity.py "tests/test_demo_gpu.py::test_demo_out_identical_to_reference[model.py-b256]" -x -q -k "forward or fused or matmul or model" > gpurun_out/r2ah/pytest.log 2>&1; echo "pytest rc=$? wall=${SECONDS}s"
tail -5 gpurun_out/r2ah/pytest.log
PROBE_TIMES=1 timeout 300 python tools/probe_forward.py > gpurun_out/r2ah/forward_times.log 2>&1; tail -2 gpurun_out/r2ah/forward_times.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/r2ah/forward_launches.csv python tools/probe_forward.py > gpurun_out/r2ah/ncu.log 2>&1; echo "ncu rc=$? wall=${SECONDS}s"
grep umma gpurun_out/r2ah/forward_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
