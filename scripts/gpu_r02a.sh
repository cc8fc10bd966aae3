#!/bin/bash
# Round 2, session a: scan rate against warps per CTA with the round-1 kernel (input to the two-CTAs-per-SM design)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02a.log) 2>&1
nproc; nvidia-smi -L
rm -f gpurun_out/probe_warps.csv
timeout 600 python scripts/probe_warps.py c1,c2,c3wm 100
