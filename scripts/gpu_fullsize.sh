#!/bin/bash
# BASELINE configs[2] / configs[3] at their full sizes, one GPU: c3 shard (1e9 B of the 8 GB text), c4 (2e9 B)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/fullsize.log) 2>&1
free -g | head -2
echo "=== bench c4 (2e9 bytes, 256 symbols, 10000 patterns m = 8..64) ==="
timeout 900 python bench.py --workload c4 --text-mib 1908 --steps 10 --warmup 3 --no-cpu | tee gpurun_out/bench_c4_full.json
echo "=== bench c3 shard (AC, 1e9 bytes of DNA, 100000 patterns m = 32) ==="
timeout 900 python bench.py --workload c3 --text-mib 954 --steps 10 --warmup 3 --no-cpu | tee gpurun_out/bench_c3_shard.json
echo "=== bench c3wm shard (WM on the same set) ==="
timeout 900 python bench.py --workload c3wm --text-mib 954 --steps 10 --warmup 3 --no-cpu | tee gpurun_out/bench_c3wm_shard.json
echo "=== bench c2 at 1 GiB ==="
timeout 900 python bench.py --workload c2 --text-mib 1024 --steps 20 --warmup 3 --no-cpu | tee gpurun_out/bench_c2_1g.json
