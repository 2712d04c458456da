mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python tools/bench_configs.py C1 C2 C4 2>&1 | tee gpurun_out/configs_r1.jsonl | cut -c1-900
