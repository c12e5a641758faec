mkdir -p gpurun_out/r2w
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "forward_and_prove or subtask" > gpurun_out/r2w/tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w/tests.log; tail -5 gpurun_out/r2w/tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/r2w/bench.json 2> gpurun_out/r2w/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2w/bench.json')); print(d['ms_per_step'], d['e2e'])"
