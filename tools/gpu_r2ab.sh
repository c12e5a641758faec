mkdir -p gpurun_out/r2ab
SECONDS=0
timeout 600 python -m pytest tests/test_linked_gpu.py -x -q > gpurun_out/r2ab/pytest_linked.log 2>&1; echo "pytest rc=$? wall=${SECONDS}s"
tail -25 gpurun_out/r2ab/pytest_linked.log
timeout 600 python tools/run_configs.py linked > gpurun_out/r2ab/linked_full.jsonl 2> gpurun_out/r2ab/linked_full.err; echo "linked rc=$? wall=${SECONDS}s"
tail -5 gpurun_out/r2ab/linked_full.err; cat gpurun_out/r2ab/linked_full.jsonl
