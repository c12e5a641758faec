mkdir -p gpurun_out/r2r
for m in 0 6 7 8; do
  if [ $m -eq 0 ]; then L=zkdl_b200/libzkdl_b200.so; else L=tools/_build/libzkdl_m$m.so; fi
  echo "min_ctas=$m"; for c in 16 24 32 48; do ZKDL_LIB=$PWD/$L ZKDL_FOLD_CAP=$c timeout 100 python tools/probe_fold.py; done
done > gpurun_out/r2r/fold_occupancy.log 2>&1
cat gpurun_out/r2r/fold_occupancy.log
