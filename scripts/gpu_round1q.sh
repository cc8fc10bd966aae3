#!/bin/bash
# Session q: single-process multi-GPU host search (acwm_search_host_sharded, smatcher_main -gpus G)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1q.log) 2>&1
nproc; nvidia-smi -L
echo "=== new tests ==="; timeout 900 python -m pytest tests -m gpu -x -q -k "sharded or multi_gpu_mode" 2>&1 | tail -8
echo "=== sharded host bench (c2) ==="; timeout 600 python scripts/sharded_host_bench.py --workload c2 --mib-per-gpu 1024 --reps 4 | tee gpurun_out/sharded_host_c2.jsonl
echo "=== pytest -m gpu (all) ==="; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_c2_q.json
echo "=== smoke ==="; timeout 600 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?"
