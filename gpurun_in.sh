mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retain.py tests/test_gpu_hypothesis.py tests/test_gpu_parity.py tests/test_gpu_router.py -x -q -m gpu 2>&1 | tail -6
