#!/bin/bash
# Round 2, session c: timelines of the two-CTAs-per-SM overlap mode (who shares an SM, launches in flight)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02c.log) 2>&1
nvidia-smi -L
for wl in c2 c1; do
  timeout 300 python scripts/trace.py $wl overlap 2>&1 | grep -v Warning
done
