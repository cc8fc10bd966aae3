#!/bin/bash
# Round 2, session t: compute-sanitizer (memcheck, racecheck on shared memory) over small cases of every front end
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02t.log) 2>&1
nvidia-smi -L
echo "=== memcheck ==="
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanity_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "=== racecheck (c1, c2, a bytes case) ==="
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanity_small.py c1_ac_dna_p100_m8 c2_wm_dna_p1000_m16 wm_ascii_p1000_m8 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log
echo "=== host path + packer tests ==="; timeout 1200 python -m pytest tests -m gpu -x -q -k "host or hybrid or packed or sharded or overflow or lookback" 2>&1 | tail -3
