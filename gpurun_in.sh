mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 8 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["trie"]["values"])
except Exception as e: print("N=$N failed", e)
PY
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
done
