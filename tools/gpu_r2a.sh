mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/gputests.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2a/bench_c8.log 2>&1
ZKDL_MSM_C=12 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2a/bench_c12.log 2>&1
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2a/subtasks_c8.log 2>&1
ZKDL_PROVE_THREADS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2a/ncu.log 2>&1
tail -5 gpurun_out/r2a/gputests.log; cat gpurun_out/r2a/bench_c8.log | tail -2; cat gpurun_out/r2a/bench_c12.log | tail -2
