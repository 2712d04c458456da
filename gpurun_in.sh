mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi4.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi4.txt; tail -5 gpurun_out/pytest_multi4.txt

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_c.json 2> gpurun_out/bench_n2_c.err; echo "bench2 exit $?"; tail -5 gpurun_out/bench_n2_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_c.json'))
for k in ('value','value_with_gather','e2e','multi_gpu','parity_check'):
    print(k, json.dumps(d.get(k))[:1800])
PY
