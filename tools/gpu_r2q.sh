mkdir -p gpurun_out/r2q
timeout 400 python tools/probe_plan.py 10 > gpurun_out/r2q/plan.json 2>&1; cat gpurun_out/r2q/plan.json
