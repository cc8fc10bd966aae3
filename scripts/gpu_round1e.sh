#!/bin/bash
# Emission through a shared list, AC entries = row byte offset, smaller WM tables; bytes path + c4 first look
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1e.log) 2>&1
echo "=== sanity (hang check) ==="; timeout 300 python scripts/sanity_small.py; echo "exit $?"
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_c2.json
echo "=== bench c1 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --workload c1 --no-cpu | tee gpurun_out/bench_c1.json
echo "=== tune ==="; rm -f gpurun_out/tune.csv; TUNE_WL=c2,c1,c2ac,c1wm,c4 timeout 1500 python scripts/tune.py
echo "=== ncu full (scan kernel) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c1 -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c1 > gpurun_out/ncu_full_c1.log 2>&1; echo "exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c4 -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c4 > gpurun_out/ncu_full_c4.log 2>&1; echo "exit $?"
ls -la gpurun_out
