mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py tests/test_golden.py tests/test_fullsize_properties.py -x -q > gpurun_out/r2g/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/tests.log
tail -4 gpurun_out/r2g/tests.log
for f in 0 131072 524288 4194304; do echo "fuse_max=$f"; ZKDL_SC_FUSE_MAX=$f timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-extras 2>&1 | tail -1 | cut -c1-170; done > gpurun_out/r2g/bench_fuse.log 2>&1
cat gpurun_out/r2g/bench_fuse.log
for b in 64 128 256; do for c in 8 16 32; do ZKDL_FOLD_BLOCK=$b ZKDL_FOLD_CAP=$c timeout 100 python tools/probe_fold.py; done; done > gpurun_out/r2g/fold_sweep.log 2>&1
cat gpurun_out/r2g/fold_sweep.log
timeout 300 python tools/probe_subtasks.py 10 > gpurun_out/r2g/subtasks.log 2>&1; tail -1 gpurun_out/r2g/subtasks.log
timeout 400 python tools/probe_plan.py 10 > gpurun_out/r2g/plan.json 2>&1; cat gpurun_out/r2g/plan.json
