#!/bin/bash
# Multi-GPU: one process per GPU (torchrun), weak scaling, NCCL all-reduce of the 8-byte count.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
exec > >(tee gpurun_out/multi_${N}.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv
for n in 1 $N; do
  echo "=== bench c2 --gpus $n ==="
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu | tee gpurun_out/scale_c2_n$n.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu | tee gpurun_out/scale_c2_n$n.json; fi
done
echo "=== bench c1 --gpus $N ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --workload c1 | tee gpurun_out/scale_c1_n$N.json
echo "=== sharded parity (torchrun, NCCL) ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/sharded_parity.py
