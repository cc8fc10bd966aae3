#!/bin/bash
# Round 2, session s: two bits per entry in hashed stage-1 bitmaps (blocked Bloom filter): configs[2] / configs[3]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02s.log) 2>&1
nvidia-smi -L
echo "=== pytest -m gpu (all) ==="; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
rm -f gpurun_out/probe_warps.csv
echo "=== step times ==="
PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c1,c2,c3,c3wm,c4 100
echo "=== counters ==="
for wl in c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum --clock-control none -k regex:scan_kernel -s 4 -c 1 --csv python scripts/one_scan.py $wl 2>/dev/null | grep -E "scan_kernel" | awk -F'","' -v wl=$wl '{print wl, $(NF-2), $(NF)}' | tr -d '"'
done
