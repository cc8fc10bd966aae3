#!/bin/bash
# Round 2, session u: stand-alone scans with and without the cooperative launch attribute
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02u.log) 2>&1
nvidia-smi -L
rm -f gpurun_out/probe_warps.csv
echo "=== cooperative (default) ==="; PROBE_OPTS='[{}, {"force_ctas": 1}]' timeout 600 python scripts/probe_warps.py c1,c2 100
echo "=== plain launch ==="; ACWM_PLAIN_LAUNCH=1 PROBE_OPTS='[{}, {"force_ctas": 1}]' timeout 600 python scripts/probe_warps.py c1,c2 100
