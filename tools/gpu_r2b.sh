mkdir -p gpurun_out/r2b
./tools/_build/addlat1 > gpurun_out/r2b/addlat.jsonl 2>&1; ./tools/_build/addlat0 >> gpurun_out/r2b/addlat.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_reference.py -x -q > gpurun_out/r2b/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b/bench_n1.json 2> gpurun_out/r2b/bench_n1.err
timeout 400 python tools/probe_plan.py 10 > gpurun_out/r2b/plan.json 2>&1
tail -3 gpurun_out/r2b/tests.log; cat gpurun_out/r2b/addlat.jsonl; cat gpurun_out/r2b/plan.json; tail -c 1500 gpurun_out/r2b/bench_n1.err
