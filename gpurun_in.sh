mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_b.json 2> gpurun_out/bench_n8_b.err; echo "bench8 exit $?"; tail -4 gpurun_out/bench_n8_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8_b.json'))
for k in ('value','value_with_gather','value_descriptor_mode','e2e','multi_gpu','parity_check'):
    print(k, json.dumps(d.get(k))[:2200])
PY
