#!/bin/bash
# Round 2, session f: ncu --set full of the c2 kernel (one CTA per SM) and the c1 kernel (two CTAs per SM);
# the reports stay on the box (> 64 MiB together), their raw and source pages come back as CSV
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02f.log) 2>&1
cap() { # name workload opts
  timeout 900 ncu --set full --clock-control none -k regex:scan_kernel -s 5 -c 1 -o /tmp/prof_$1 -f python scripts/one_scan.py $2 "$3" 2>&1 | tail -2
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep gpurun_out/ncu_$1_summary.csv
}
cap c2_single c2 '{"force_ctas": 1}'
cap c1_dual c1 '{}'
ls -la gpurun_out/ncu_*
