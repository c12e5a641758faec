mkdir -p gpurun_out/r2z
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z/bench_ours_n1.json 2> gpurun_out/r2z/bench_ours_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z/bench_reference_n1.json 2> gpurun_out/r2z/bench_reference_n1.err; echo "ref rc=$?"
P="python tools/probe_kernels_r2.py"
N="ncu --set full --clock-control none --import-source on"
timeout 300 $P > gpurun_out/r2z/probe_plain.log 2>&1 && {
timeout 600 $N -k regex:"k_msm_accumulate|k_msm_combine|k_msm_reduce_coop|k_msm_reduce_scan" -c 4 -o gpurun_out/r2z/k_msm $P > gpurun_out/r2z/ncu_msm.log 2>&1; echo "ncu msm rc=$?"
timeout 600 $N -k regex:"k_umma_matmul" -c 1 -o gpurun_out/r2z/k_umma $P > gpurun_out/r2z/ncu_umma.log 2>&1; echo "ncu umma rc=$?"
timeout 600 $N -k regex:"k_bin_packed3|k_bin_r34" -c 4 -o gpurun_out/r2z/k_bin $P > gpurun_out/r2z/ncu_bin.log 2>&1; echo "ncu bin rc=$?"
timeout 600 $N -k regex:"k_sc_tail|k_sc_round" -c 3 -o gpurun_out/r2z/k_sc $P > gpurun_out/r2z/ncu_sc.log 2>&1; echo "ncu sc rc=$?"
timeout 600 $N -k regex:"k_fr_fold_multi" -c 1 -o gpurun_out/r2z/k_fold $P > gpurun_out/r2z/ncu_fold.log 2>&1; echo "ncu fold rc=$?"
}
for f in k_msm k_umma k_bin k_sc k_fold; do ncu -i gpurun_out/r2z/$f.ncu-rep --page raw --csv > gpurun_out/r2z/$f.raw.csv 2>/dev/null; done
ZKDL_PROVE_THREADS=0 timeout 300 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2z/launch_plain.log 2>&1 && ZKDL_PROVE_THREADS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/r2z/launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2z/ncu_launches.log 2>&1; echo "ncu list rc=$?"
du -sh gpurun_out/r2z; ls -la gpurun_out/r2z
# keep the merge under 64 MiB: drop the biggest reports if needed (the raw CSVs stay)
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/r2z/k_bin.ncu-rep gpurun_out/r2z/k_sc.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/r2z/*.ncu-rep; fi
du -sh gpurun_out
