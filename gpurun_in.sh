mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_b.json").read().strip().splitlines()[-1])
print("N=1 value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
PY
