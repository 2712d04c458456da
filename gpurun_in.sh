mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
timeout 600 python tools/bench_configs.py C1 C2 C4 2>&1 | tee gpurun_out/configs.jsonl | cut -c1-250
