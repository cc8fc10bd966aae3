#!/bin/bash
# Kernel timelines from in-kernel %globaltimer stamps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for wl in c2 c1; do for mode in overlap coop; do
  timeout 300 python scripts/trace.py $wl $mode 2>&1 | grep -v Warning
done; done
