mkdir -p gpurun_out/r2s
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2s/bench_ours_n2.json 2> gpurun_out/r2s/bench_ours_n2.err
echo "rc=$?"; grep -c "^{" gpurun_out/r2s/bench_ours_n2.json; grep "^{" gpurun_out/r2s/bench_ours_n2.json | cut -c1-250
