mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retain.py tests/test_gpu_hypothesis.py tests/test_gpu_parity.py tests/test_gpu_desc_comm.py "tests/test_gpu_fullsize.py::test_retained_full_size_every_filter_bit_exact" -m gpu -x -q > gpurun_out/pytest_gpu2.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.txt
tail -15 gpurun_out/pytest_gpu2.txt
timeout 600 python tools/ab_round2.py c4 tok e2e > gpurun_out/ab_round2.jsonl 2> gpurun_out/ab_round2.err; echo "ab exit $?"; tail -3 gpurun_out/ab_round2.err
cat gpurun_out/ab_round2.jsonl
