mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2i/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/gputests.log
tail -14 gpurun_out/r2i/gputests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i/bench_ours_n1.json 2> gpurun_out/r2i/bench_ours_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2i/bench_reference_n1.json 2> gpurun_out/r2i/bench_reference_n1.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r2i/bench_reference_n1.json
timeout 300 python tools/probe_kernels_r2.py > gpurun_out/r2i/probe_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate|k_msm_reduce_scan|k_msm_combine|k_umma_matmul|k_bin_r34|k_bin_packed3|k_sc_tail|k_fr_fold_multi|k_sc_round" -c 60 -o gpurun_out/r2i/kernels python tools/probe_kernels_r2.py > gpurun_out/r2i/ncu_full.log 2>&1; echo "ncu full rc=$?"
ZKDL_PROVE_THREADS=0 timeout 300 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2i/launch_plain.log 2>&1 && ZKDL_PROVE_THREADS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/r2i/launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2i/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/r2i
