#!/bin/bash
# Round 2, session h: lean scan loop (32-bit shared addresses, funnel-shift hit words, one-select DFA step, lane-local
# candidate checks): parity first, then step times, instruction / wavefront counts of the headline kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02h.log) 2>&1
nproc; nvidia-smi -L
echo "=== pytest -m gpu (all) ==="; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
rm -f gpurun_out/probe_warps.csv
echo "=== step times: default / one CTA per SM ==="
PROBE_OPTS='[{}, {"force_ctas": 1}]' timeout 600 python scripts/probe_warps.py c1,c2,c2ac,c3,c3wm,c4 100
echo "=== counters ==="
for wl in c1 c2 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum --clock-control none -k regex:scan_kernel -s 4 -c 1 --csv python scripts/one_scan.py $wl 2>/dev/null | grep -E "scan_kernel" | awk -F'","' -v wl=$wl '{print wl, $(NF-2), $(NF)}' | tr -d '"'
done
echo "=== bench (default) ==="; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; tail -c 400 gpurun_out/bench_default.err; head -c 1500 gpurun_out/bench_default.json; echo
