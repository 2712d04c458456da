mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_final.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.txt; tail -12 gpurun_out/pytest_gpu_final.txt
