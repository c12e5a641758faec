mkdir -p gpurun_out/r2d
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d/bench_n2.json 2> gpurun_out/r2d/bench_n2.err
echo "rc=$?"; tail -c 600 gpurun_out/r2d/bench_n2.err; cut -c1-400 gpurun_out/r2d/bench_n2.json
