mkdir -p gpurun_out/san
timeout 300 python __graft_entry__.py smoke > gpurun_out/san/plain.log 2>&1 && timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san/memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/san/memcheck_smoke.log
