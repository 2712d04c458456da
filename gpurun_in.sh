mkdir -p gpurun_out
python tools/c4_once.py stats > gpurun_out/c4_stats.txt 2>&1; tail -5 gpurun_out/c4_stats.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_retain_round --launch-skip 7 --launch-count 7 -o gpurun_out/r2_retain_round python tools/c4_once.py > gpurun_out/c4_ncu.log 2>&1; tail -3 gpurun_out/c4_ncu.log
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_batcher.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu3.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu3.txt; tail -8 gpurun_out/pytest_gpu3.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; echo "bench exit $?"; tail -5 gpurun_out/bench_n1_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_b.json'))
for k in ('value','value_descriptor_mode','e2e','multi_gpu','parity_check','c4','latency','churn'):
    print(k, json.dumps(d.get(k))[:1500])
PY
