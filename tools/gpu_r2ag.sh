mkdir -p gpurun_out/r2ag
SECONDS=0
PROBE_TIMES=1 timeout 300 python tools/probe_forward.py > gpurun_out/r2ag/forward_times.log 2>&1; echo "plain rc=$? wall=${SECONDS}s"; cat gpurun_out/r2ag/forward_times.log | tail -3
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_umma_matmul" -s 3 -c 1 -o gpurun_out/r2ag/k_umma_fused python tools/probe_forward.py > gpurun_out/r2ag/ncu.log 2>&1; echo "ncu rc=$? wall=${SECONDS}s"
ncu -i gpurun_out/r2ag/k_umma_fused.ncu-rep --page raw --csv > gpurun_out/r2ag/k_umma_fused.raw.csv 2>/dev/null
ls -la gpurun_out/r2ag
