mkdir -p gpurun_out
python tools/c4_once.py stats 2>&1 | grep "kernel ms\|retain stats" > gpurun_out/c4_inline.txt; cat gpurun_out/c4_inline.txt
python tools/c4_once.py 2>&1 | grep "kernel ms" >> gpurun_out/c4_inline.txt; tail -1 gpurun_out/c4_inline.txt
timeout 900 python -m pytest tests/test_gpu_retain.py tests/test_gpu_hypothesis.py tests/test_gpu_batcher.py "tests/test_gpu_fullsize.py::test_retained_full_size_every_filter_bit_exact" -m gpu -x -q > gpurun_out/pytest_gpu8.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.txt; tail -6 gpurun_out/pytest_gpu8.txt
