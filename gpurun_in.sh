mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("$1 failed", e)
PY
tail -2 gpurun_out/b_$1.err | cut -c1-300
}
GM_SORTED_ROWS=1 run rows1
GM_SORTED_ROWS=0 run rows0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batches 2 --e2e-steps 1 > /dev/null 2>&1
