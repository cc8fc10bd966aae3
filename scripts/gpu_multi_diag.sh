#!/bin/bash
# Multi-GPU diagnosis: one short bench with stderr kept
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv | head -12
nvidia-smi topo -m 2>/dev/null | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/diag_n$N.json 2> gpurun_out/diag_n$N.err; echo "exit $?"
cat gpurun_out/diag_n$N.json | cut -c1-400
grep -v "Warning\|warn" gpurun_out/diag_n$N.err | tail -40
