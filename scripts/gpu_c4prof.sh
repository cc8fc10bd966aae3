#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/prof_c4 -f python bench.py --steps 4 --warmup 3 --no-cpu --no-overlap --workload c4 > gpurun_out/ncu_full_c4.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py gpurun_out/prof_c4.ncu-rep gpurun_out/ncu_full_c4_summary.csv
timeout 300 python scripts/trace.py c4 overlap 2>&1 | grep -v Warning | tail -22
