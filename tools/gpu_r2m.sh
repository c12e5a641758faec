mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_fiat_shamir_gpu.py tests/test_fullsize_properties.py tests/test_golden.py -x -q > gpurun_out/r2m/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m/tests.log
tail -6 gpurun_out/r2m/tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-200 > gpurun_out/r2m/bench_short.log; cat gpurun_out/r2m/bench_short.log
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2m/subtasks.log 2>&1; tail -1 gpurun_out/r2m/subtasks.log
