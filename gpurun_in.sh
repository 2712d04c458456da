mkdir -p gpurun_out
python tools/c4_once.py 2>&1 | grep "kernel ms" > gpurun_out/c4_dyn.txt; cat gpurun_out/c4_dyn.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu6.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu6.txt; tail -16 gpurun_out/pytest_gpu6.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_c.json'))
for k in ('value','value_descriptor_mode','e2e','c4','relations'):
    print(k, json.dumps(d.get(k))[:900])
print('latency', [(r['offered_burst'], round(r['p50_us'],1), round(r['p99_us'],1), int(r['topics_per_s'])) for r in d['latency']['table']])
PY
