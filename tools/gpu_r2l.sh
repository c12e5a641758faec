mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_fiat_shamir_gpu.py tests/test_proof_file_gpu.py tests/test_fullsize_properties.py -x -q > gpurun_out/r2l/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l/tests.log
tail -4 gpurun_out/r2l/tests.log
