mkdir -p gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_golden.py tests/test_fullsize_properties.py -x -q > gpurun_out/r2p/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2p/tests.log
tail -6 gpurun_out/r2p/tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-200 > gpurun_out/r2p/bench_coop.log; cat gpurun_out/r2p/bench_coop.log
ZKDL_MSM_NO_COOP=1 timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-200 > gpurun_out/r2p/bench_nocoop.log; cat gpurun_out/r2p/bench_nocoop.log
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2p/subtasks.log 2>&1; tail -1 gpurun_out/r2p/subtasks.log | cut -c1-400
timeout 400 python tools/probe_plan.py 10 > gpurun_out/r2p/plan.json 2>&1; cat gpurun_out/r2p/plan.json
