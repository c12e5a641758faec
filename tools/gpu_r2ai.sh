mkdir -p gpurun_out/r2ai
SECONDS=0
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai/gputests.log 2>&1; echo "pytest rc=$? wall=${SECONDS}s" | tee -a gpurun_out/r2ai/gputests.log
tail -3 gpurun_out/r2ai/gputests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2ai/smoke.log 2>&1; tail -1 gpurun_out/r2ai/smoke.log
timeout 200 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/r2ai/bench_ours_n1_nocpu.json 2> gpurun_out/r2ai/bench.err; echo "bench rc=$? wall=${SECONDS}s"
cut -c1-200 gpurun_out/r2ai/bench_ours_n1_nocpu.json
