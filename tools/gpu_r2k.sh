mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_fiat_shamir_gpu.py tests/test_proof_file_gpu.py -x -q > gpurun_out/r2k/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k/tests.log
tail -4 gpurun_out/r2k/tests.log
timeout 300 python tools/probe_forward.py > gpurun_out/r2k/forward.log 2>&1; cat gpurun_out/r2k/forward.log
timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-200 > gpurun_out/r2k/bench_short.log; cat gpurun_out/r2k/bench_short.log
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2k/subtasks.log 2>&1; tail -1 gpurun_out/r2k/subtasks.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2k/smoke.log 2>&1; tail -2 gpurun_out/r2k/smoke.log
