mkdir -p gpurun_out/r2t
for L in zkdl_b200/libzkdl_b200.so tools/_build/libzkdl_sc3.so; do
  echo "lib=$L"
  ZKDL_LIB=$PWD/$L timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-170
  ZKDL_LIB=$PWD/$L timeout 300 python tools/probe_subtasks.py 10 2>&1 | tail -1 | cut -c1-330
done > gpurun_out/r2t/sc_occupancy.log 2>&1
cat gpurun_out/r2t/sc_occupancy.log
