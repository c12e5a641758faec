mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_golden.py -x -q > gpurun_out/r2x/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2x/tests.log; tail -4 gpurun_out/r2x/tests.log
timeout 300 python tools/run_configs.py msm 22 > gpurun_out/r2x/msm_coop.jsonl 2>&1; cut -c1-200 gpurun_out/r2x/msm_coop.jsonl
ZKDL_MSM_NO_COOP=1 timeout 300 python tools/run_configs.py msm 22 > gpurun_out/r2x/msm_nocoop.jsonl 2>&1; cut -c1-200 gpurun_out/r2x/msm_nocoop.jsonl
