mkdir -p gpurun_out
timeout 35 compute-sanitizer --tool memcheck --print-limit 5 python tools/memcheck_workload.py > gpurun_out/memcheck_workload.txt 2>&1; tail -5 gpurun_out/memcheck_workload.txt
