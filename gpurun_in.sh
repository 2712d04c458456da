mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.txt; tail -8 gpurun_out/pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit $?"; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
for k in ('value','value_with_gather','value_descriptor_mode','e2e','multi_gpu','parity_check'):
    print(k, json.dumps(d.get(k))[:1500])
PY
./tools/randbench6 > gpurun_out/randbench6.txt 2>&1; cat gpurun_out/randbench6.txt
