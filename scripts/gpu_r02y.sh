#!/bin/bash
# Round 2, session y: c2 with stride 4 (13-symbol blocks, hashed two-bit bitmap) against the planner's stride 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02y.log) 2>&1
rm -f gpurun_out/probe_warps.csv
PROBE_OPTS='[{}, {"force_stride": 4}, {"force_stride": 4, "force_ctas": 1}, {"force_stride": 2}]' timeout 600 python scripts/probe_warps.py c2,c2ac 100
