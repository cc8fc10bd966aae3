#!/bin/bash
# Session k: warp-cooperative candidate verification + early retire of sparse scans -- parity, A/B, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1k.log) 2>&1
echo "=== pytest -m gpu ==="; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== A/B (ACWM_TUNE: 1 = cooperative verification, 2 = early retire) ==="
rm -f gpurun_out/ab.csv
for t in 0 1 2 3; do
  ACWM_TUNE=$t timeout 600 python scripts/ab.py c2,c1,c2ac,c1wm 100 128 2>&1 | grep -v Warning
done
for t in 0 3; do
  ACWM_TUNE=$t timeout 600 python scripts/ab.py c4,c3wm 50 128 2>&1 | grep -v Warning
done
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu | tee gpurun_out/bench_c2.json
echo "=== bench c1 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --workload c1 | tee gpurun_out/bench_c1.json
ls -la gpurun_out
