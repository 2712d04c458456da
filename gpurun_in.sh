mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_router.py tests/test_gpu_retain.py tests/test_gpu_hypothesis.py tests/test_gpu_batcher.py "tests/test_gpu_fullsize.py::test_retained_full_size_every_filter_bit_exact" -m gpu -x -q > gpurun_out/pytest_gpu4.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu4.txt; tail -12 gpurun_out/pytest_gpu4.txt
python tools/c4_once.py stats > gpurun_out/c4_stats2.txt 2>&1; tail -4 gpurun_out/c4_stats2.txt
python - <<'PY' > gpurun_out/relations_leg.json 2> gpurun_out/relations_leg.err
import json, bench
print(json.dumps(bench._relations_leg(False)))
PY
cat gpurun_out/relations_leg.json; tail -3 gpurun_out/relations_leg.err
