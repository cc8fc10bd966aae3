#!/bin/bash
# EXPERIMENT: programmatic dependent launch between consecutive scans when two grids can be co-resident
# ACWM_PDL=1: cooperative + programmatic attributes; ACWM_PDL=2: programmatic attribute only
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1i.log) 2>&1
for shape in '{"force_threads":256,"force_stages":1}' '{"force_threads":384,"force_stages":1}' '{}'; do
  for pdl in 0 2; do
    echo "=== c2 shape=$shape ACWM_PDL=$pdl ==="
    ACWM_PDL=$pdl timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --matcher-opts "$shape" | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('us/step', round(d['ms_per_step']*1000,2), 'GB/s', round(d['value'],1), 'isolated', round(d['roofline']['kernel_ms_isolated_launch']*1000,2), d['config']['kernel']['threads'], d['config']['kernel']['smem_bytes'], 'matches', d['matches_last_step'])"
  done
done
