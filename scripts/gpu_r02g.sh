#!/bin/bash
# Round 2, session g: the bench as the driver runs it (both arms), smoke, the whole GPU suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02g.log) 2>&1
nproc; nvidia-smi -L
echo "=== bench (ours, default) ==="; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; tail -c 600 gpurun_out/bench_default.err; head -c 3000 gpurun_out/bench_default.json; echo
echo "=== bench (reference arm) ==="; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; head -c 1500 gpurun_out/bench_ref.json; echo
echo "=== smoke ==="; timeout 600 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?"
echo "=== pytest -m gpu (all) ==="; timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
