mkdir -p gpurun_out/r2u
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2u/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u/gputests.log
tail -12 gpurun_out/r2u/gputests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2u/smoke.log 2>&1; tail -1 gpurun_out/r2u/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u/bench_ours_n1.json 2> gpurun_out/r2u/bench_ours_n1.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r2u/bench_ours_n1.json
