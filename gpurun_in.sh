mkdir -p gpurun_out
python tools/c4_once.py stats 2>&1 | grep "kernel ms\|retain stats" > gpurun_out/c4_locality.txt; cat gpurun_out/c4_locality.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu7.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu7.txt; tail -14 gpurun_out/pytest_gpu7.txt
python tools/ab_round2.py tok 2>/dev/null | head -2
