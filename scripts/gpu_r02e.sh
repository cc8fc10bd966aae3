#!/bin/bash
# Round 2, session e: leaner tile pipeline (registers, no proxy fence, LDS table refs), C-ABI device-sharded path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02e.log) 2>&1
nproc; nvidia-smi -L
echo "=== parity subset ==="
timeout 1500 python -m pytest tests -m gpu -x -q -k "random_cases or edge_cases or overlapped or dense_matches or launch_shape or unaligned or device_sharded or shims or sharded" 2>&1 | tail -15
rm -f gpurun_out/probe_warps.csv
echo "=== step times: default / one CTA per SM ==="
PROBE_OPTS='[{}, {"force_ctas": 1}]' timeout 600 python scripts/probe_warps.py c1,c2,c2ac,c3,c3wm,c4 100
