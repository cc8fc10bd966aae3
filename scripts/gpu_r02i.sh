#!/bin/bash
# Round 2, session i: per-warp areas at constant shared addresses, lane-local threshold A/B, bucket pre-check in front of
# the automaton of filtered AC; source-level profiles of the c1 / c2 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02i.log) 2>&1
nvidia-smi -L
echo "=== parity subset ==="
timeout 1500 python -m pytest tests -m gpu -x -q -k "random_cases or edge_cases or overlapped or dense_matches or launch_shape or unaligned or sweep or planted" 2>&1 | tail -5
rm -f gpurun_out/probe_warps.csv
echo "=== step times by lane-local threshold (ACWM_TUNE) ==="
for t in 0x0001 0x0301 0x0801 0x1401; do echo "tune $t"; ACWM_TUNE=$t PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c1,c2,c3,c3wm,c4 100; done
cap() { # name workload opts
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:scan_kernel -s 5 -c 1 -o /tmp/prof_$1 -f python scripts/one_scan.py $2 "$3" 2>&1 | tail -1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep gpurun_out/ncu_$1_summary.csv
}
cap c2_r02i c2 '{}'
cap c1_r02i c1 '{}'
cap c4_r02i c4 '{}'
ls -la gpurun_out/ncu_*r02i*
