mkdir -p gpurun_out/r2f
timeout 600 bash tools/run_demo_full.sh > gpurun_out/r2f/demo_full.log 2>&1; cat gpurun_out/r2f/demo_full.log
timeout 600 python -m pytest tests/test_demo_gpu.py tests/test_golden.py -x -q -k "toy or golden or host" > gpurun_out/r2f/tests.log 2>&1; tail -3 gpurun_out/r2f/tests.log
