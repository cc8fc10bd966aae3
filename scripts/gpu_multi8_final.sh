#!/bin/bash
# 8 x B200: BASELINE configs[2] at full size (AC, 8e9 bytes of DNA over 8 GPUs, 100000 patterns m=32), the c2 weak-scaling
# line of the final build, sharded parity with the in-kernel count exchange
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
exec > >(tee gpurun_out/multi${N}_final.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv | head -12; free -g | head -2
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2> gpurun_out/stderr_$1.log; echo "exit $?"; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/stderr_$1.log | tail -5; }
echo "=== bench c2 --gpus $N ==="; run 29541 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu | tee gpurun_out/final_c2_n$N.json
echo "=== bench c3 --gpus $N (8e9 bytes) ==="; run 29542 bench.py --gpus $N --workload c3 --text-mib 954 --steps 10 --warmup 3 --no-cpu | tee gpurun_out/final_c3_n$N.json
echo "=== bench c1 --gpus $N ==="; run 29543 bench.py --gpus $N --workload c1 --steps 50 --warmup 5 --no-cpu | tee gpurun_out/final_c1_n$N.json
echo "=== sharded parity ==="; run 29544 scripts/sharded_parity.py
