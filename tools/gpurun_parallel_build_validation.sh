# The command list of the last gpurun call of round 2 (validation of the parallel host builds on the GPU box): kept for the record.
mkdir -p gpurun_out
nproc > gpurun_out/r2_hostbuild_box.txt; lscpu | grep -i "model name\|socket\|numa node(s)" >> gpurun_out/r2_hostbuild_box.txt
GM_BULK_PROFILE=1 timeout 220 python -m pytest tests/test_gpu_fullsize.py -x -q -s --durations=5 > gpurun_out/r2_pytest_fullsize_parallel_build.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_fullsize_parallel_build.txt
grep -a "insert_batch\|retained bulk\|sync:\|reserve:\|passed\|failed\|exit" gpurun_out/r2_pytest_fullsize_parallel_build.txt | tail -20
GM_BULK_PROFILE=1 timeout 170 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_parallel_build.json 2> gpurun_out/r2_bench_n1_parallel_build.err; echo "bench exit $?"
grep -a "insert_batch\|retained bulk\|sync:\|reserve:" gpurun_out/r2_bench_n1_parallel_build.err | tail; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1_parallel_build.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d.get("e2e"), d["details"].get("build_s"), d["details"].get("build"), d.get("c4",{}).get("build_s") if isinstance(d.get("c4"),dict) else None)
    print("parity", d.get("parity_check"))
except Exception as e: print("no bench line", e)
PY
