mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_proof_file_gpu.py tests/test_golden.py -x -q > gpurun_out/r2c/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c/tests.log
tail -5 gpurun_out/r2c/tests.log
for w in 2 1 0.5; do ZKDL_MSM_WAVES=$w timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-200; done > gpurun_out/r2c/bench_waves.log 2>&1
cat gpurun_out/r2c/bench_waves.log
ZKDL_MM_NO_UMMA=1 timeout 300 python tools/probe_forward.py > gpurun_out/r2c/forward_noumma.log 2>&1; tail -3 gpurun_out/r2c/forward_noumma.log
timeout 300 python tools/probe_forward.py > gpurun_out/r2c/forward_umma.log 2>&1; tail -3 gpurun_out/r2c/forward_umma.log
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2c/subtasks.log 2>&1; tail -1 gpurun_out/r2c/subtasks.log
