mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
AB_CHUNKS=1 AB_REPS=30 timeout 600 python tools/ab_windows.py 8 2>&1 | tee gpurun_out/ab_final.jsonl | cut -c1-330
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
grep -i -E "AnonHugePages|MemFree" /proc/meminfo; cat /sys/kernel/mm/transparent_hugepage/enabled
