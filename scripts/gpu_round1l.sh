#!/bin/bash
# Session l: deep overlap (scratch arrays by launch parity, Work ring of 3, wait at arrival), late-arriver ordering
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1l.log) 2>&1
echo "=== pytest -m gpu ==="; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
rm -f gpurun_out/ab.csv
for t in 3 1; do
  ACWM_TUNE=$t timeout 600 python scripts/ab.py c2,c1,c2ac,c1wm,c4,c3wm 100 128 2>&1 | grep -v Warning
done
for wl in c2 c1; do timeout 300 python scripts/trace.py $wl overlap 2>&1 | grep -v Warning; done
