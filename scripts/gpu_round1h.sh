#!/bin/bash
# AC automaton in global memory (L2) for large sets; placement thresholds; regression
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1h.log) 2>&1
echo "=== sanity (hang check) ==="; timeout 300 python scripts/sanity_small.py c1_ac_dna_p100_m8 c3_ac_dna_p100000_m32 ac_dna_p10000_m16_f2_l2 c4_wm_ascii_p10000_mixed_l2 c3_wm_dna_p100000_m32_l2; echo "exit $?"
echo "=== compute-sanitizer memcheck (small) ==="
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanity_small.py ac_dna_p10000_m16_f2_l2 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== tune ==="; rm -f gpurun_out/tune.csv; TUNE_WL=c3,ac10k16,c4,c3wm,c2ac timeout 1800 python scripts/tune.py
echo "=== ncu full ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o /tmp/prof_c3 -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c3 > gpurun_out/ncu_full_c3.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/prof_c3.ncu-rep gpurun_out/ncu_full_c3_summary.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o /tmp/prof_c3wm -f python bench.py --steps 4 --warmup 3 --no-cpu --workload c3wm > gpurun_out/ncu_full_c3wm.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/prof_c3wm.ncu-rep gpurun_out/ncu_full_c3wm_summary.csv
ls -la gpurun_out
