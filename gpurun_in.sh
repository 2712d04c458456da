mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -5 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | head -c 6000
