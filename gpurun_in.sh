mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print("$1 failed", e)
PY
tail -1 gpurun_out/b_$1.err | cut -c1-200
}
run prefetch
