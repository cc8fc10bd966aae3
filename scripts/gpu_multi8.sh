#!/bin/bash
# 8 x B200: weak-scaling bench line + sharded parity with the in-kernel count exchange
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
exec > >(tee gpurun_out/multi_${N}.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv | head -3
for n in $N 4; do
echo "=== bench c2 --gpus $n ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu 2>/dev/null | tee gpurun_out/scale_c2_n$n.json
done
echo "=== bench c2 --gpus $N --nccl-count ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --nccl-count 2>/dev/null | tee gpurun_out/scale_c2_n${N}_nccl.json
echo "=== sharded parity ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 scripts/sharded_parity.py 2>/dev/null
