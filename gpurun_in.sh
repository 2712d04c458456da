mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
AB_CHUNKS=1,8,16 AB_REPS=40 timeout 900 python tools/ab_windows.py 0 8 2>&1 | tee gpurun_out/ab_windows2.jsonl | cut -c1-200
