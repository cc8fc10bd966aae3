#!/bin/bash
# What the driver runs at round end, on N GPUs of one box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
exec > >(tee gpurun_out/driver_like_$N.log) 2>&1
echo "=== smoke ==="; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning
echo "=== bench.py --impl reference (defaults) ==="; timeout 900 python bench.py --impl reference | cut -c1-400
echo "=== bench.py (defaults) ==="; timeout 900 python bench.py | tee gpurun_out/driver_like_bench_n1.json | cut -c1-300
for n in $N; do
echo "=== torchrun bench.py --impl reference --gpus $n ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --impl reference --gpus $n --steps 2 --warmup 3 2>/dev/null | cut -c1-300
echo "=== torchrun bench.py --gpus $n ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $n --steps 50 --warmup 5 2>/dev/null | tee gpurun_out/driver_like_bench_n$n.json | cut -c1-300
done
