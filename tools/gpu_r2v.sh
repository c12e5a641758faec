mkdir -p gpurun_out/r2v
timeout 900 python tools/run_configs.py fc4096 deep fs > gpurun_out/r2v/configs.jsonl 2> gpurun_out/r2v/configs.err; echo "rc=$?"; cat gpurun_out/r2v/configs.jsonl; tail -3 gpurun_out/r2v/configs.err
