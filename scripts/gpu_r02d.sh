#!/bin/bash
# Round 2, session d: ncu --set full of the c2 kernel (one CTA per SM) with source counters
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02d.log) 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c2_single -f python scripts/one_scan.py c2 '{"force_ctas": 1}' 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
