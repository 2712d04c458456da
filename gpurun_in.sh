mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 600 gpurun_out/bench_final_n1.json; tail -2 gpurun_out/bench_final_n1.err
timeout 900 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 700 gpurun_out/bench_final_ref.json
timeout 900 python tools/bench_configs.py C1 C2 C4 > gpurun_out/configs_final.jsonl 2>gpurun_out/configs_final.err; cut -c1-400 gpurun_out/configs_final.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match_fast -s 5 -c 1 -o gpurun_out/prof_k2_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batches 2 --e2e-steps 1 > gpurun_out/ncu_k2_final.log 2>&1; tail -1 gpurun_out/ncu_k2_final.log | cut -c1-80
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batches 3 --e2e-steps 1 > /dev/null 2>&1
