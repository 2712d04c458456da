mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py C3Z C3 2>&1 | tee gpurun_out/configs_c3z.jsonl | cut -c1-700
