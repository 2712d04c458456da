mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_batcher.py -m gpu -x -q 2>&1 | tail -15; done > gpurun_out/batcher_repeat2.txt 2>&1
grep -c " passed" gpurun_out/batcher_repeat2.txt; grep -B2 -A12 "libgpumqtt batcher\|FAILED\|Error" gpurun_out/batcher_repeat2.txt | head -60
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/memcheck_workload.py > gpurun_out/memcheck_workload.txt 2>&1; tail -4 gpurun_out/memcheck_workload.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/memcheck_workload.py > gpurun_out/racecheck_workload.txt 2>&1; tail -4 gpurun_out/racecheck_workload.txt
