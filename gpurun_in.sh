mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; head -c 330 gpurun_out/bench_n1.json; echo; grep -o '"roofline": {[^}]*}' gpurun_out/bench_n1.json | head -c 600; echo
AB_CHUNKS=1 AB_REPS=30 AB_KNOBS="diag_flags=0,diag_flags=2,diag_flags=0" timeout 600 python tools/ab_windows.py 8 2>&1 | tee gpurun_out/ab_walk_only.jsonl | cut -c1-220
