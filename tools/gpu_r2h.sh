mkdir -p gpurun_out/r2h
timeout 600 python -m pytest tests/test_fiat_shamir_gpu.py tests/test_gpu_parity.py -x -q > gpurun_out/r2h/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h/tests.log
tail -25 gpurun_out/r2h/tests.log
