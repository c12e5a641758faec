mkdir -p gpurun_out/r2aa
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2aa/bench_ours_n1.json 2> gpurun_out/r2aa/bench_ours_n1.err; echo "bench rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2aa/bench_ours_n1.err; cut -c1-200 gpurun_out/r2aa/bench_ours_n1.json
