#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/sweep.log) 2>&1
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
rm -f gpurun_out/ab.csv
timeout 600 python scripts/ab.py c2,c1 100 128 2>&1 | grep -v Warning
echo "=== sweep ==="; timeout 1500 python scripts/sweep.py 2>&1 | grep -v Warning
