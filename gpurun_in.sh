mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/b_$1.json").read().strip().splitlines()[-1])
print("$1", "value %.3e"%d["value"], d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "tables %.2f GB"%(d["config"]["trie"]["device_bytes"]/1e9))
PY
}
GM_K2_HINTS=0 run nohint
GM_K2_HINTS=1 run hint
GM_K2_HINTS=1 GM_EDGE_SLOTS_PER_FILTER=13 run hint_lf
GM_K2_HINTS=0 GM_EDGE_SLOTS_PER_FILTER=13 run nohint_lf
