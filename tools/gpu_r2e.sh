mkdir -p gpurun_out/r2e
timeout 600 bash tools/run_demo_full.sh ref > gpurun_out/r2e/demo_full.log 2>&1; cat gpurun_out/r2e/demo_full.log
