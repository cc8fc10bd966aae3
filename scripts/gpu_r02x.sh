#!/bin/bash
# Round 2, session x: the configs[4] sweep (p 10..100k x m 8..64, AC and WM) with the final build, one GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02x.log) 2>&1
nvidia-smi -L
timeout 2400 python scripts/sweep.py 2>&1 | grep -v Warning
