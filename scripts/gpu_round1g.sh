#!/bin/bash
# L2-resident tables for large pattern sets (c3wm, c4), regression check of c1/c2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1g.log) 2>&1
echo "=== sanity (hang check) ==="; timeout 300 python scripts/sanity_small.py c1_ac_dna_p100_m8 c2_wm_dna_p1000_m16 c3_wm_dna_p100000_m32_l2 wm_dna_p10000_m16_l2 c4_wm_ascii_p10000_mixed_l2 c3_ac_dna_p100000_m32; echo "exit $?"
echo "=== compute-sanitizer memcheck (small) ==="
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanity_small.py wm_dna_p10000_m16_l2 c4_wm_ascii_p10000_mixed_l2 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu | tee gpurun_out/bench_c2.json
echo "=== tune ==="; rm -f gpurun_out/tune.csv; TUNE_WL=c2,c1,c3wm,c4,c3 timeout 1800 python scripts/tune.py
echo "=== ncu full ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c3wm -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c3wm > gpurun_out/ncu_full_c3wm.log 2>&1; echo "exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c4 -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c4 > gpurun_out/ncu_full_c4.log 2>&1; echo "exit $?"
ls -la gpurun_out
