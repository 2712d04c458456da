mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_router.py "tests/test_gpu_parity.py::test_extra_trees_ride_in_the_same_batch" tests/test_gpu_batcher.py -m gpu -x -q > gpurun_out/pytest_gpu5.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu5.txt; tail -12 gpurun_out/pytest_gpu5.txt
for g in 32 64 128; do echo "== L2 fetch $g"; GM_L2_FETCH=$g python tools/c4_once.py 2>&1 | grep "kernel ms"; GM_L2_FETCH=$g python tools/ab_round2.py tok 2>/dev/null | head -2; done > gpurun_out/ab_l2fetch.txt 2>&1
cat gpurun_out/ab_l2fetch.txt
python - <<'PY' > gpurun_out/relations_leg2.json 2> gpurun_out/relations_leg2.err
import json, bench
print(json.dumps(bench._relations_leg(False)))
PY
cat gpurun_out/relations_leg2.json; tail -3 gpurun_out/relations_leg2.err
