#!/bin/bash
# Fused arrival (count + barrier + ticket in one atomic), smem tile counts, programmatic dependent launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1d.log) 2>&1
echo "=== sanity (hang check) ==="; timeout 300 python scripts/sanity_small.py; echo "exit $?"
echo "=== compute-sanitizer memcheck + racecheck (small) ==="
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanity_small.py c1_ac_dna_p100_m8 c2_wm_dna_p1000_m16 wm_ascii_p1000_m8 ac_dna_depth5 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_c2.json
echo "=== bench c2 no overlap ==="; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-overlap | tee gpurun_out/bench_c2_noovl.json
echo "=== bench c1 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --workload c1 --no-cpu | tee gpurun_out/bench_c1.json
echo "=== tune ==="; rm -f gpurun_out/tune.csv; TUNE_WL=c2,c1 timeout 1200 python scripts/tune.py
echo "=== ncu full (scan kernel) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c2 -f python bench.py --steps 4 --warmup 3 --no-cpu --no-overlap > gpurun_out/ncu_full_c2.log 2>&1; echo "exit $?"
ls -la gpurun_out
