mkdir -p gpurun_out/r2ad
SECONDS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/r2ad/forward_launches.csv python tools/probe_forward.py > gpurun_out/r2ad/ncu.log 2>&1; echo "ncu rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2ad/ncu.log; wc -l gpurun_out/r2ad/forward_launches.csv
