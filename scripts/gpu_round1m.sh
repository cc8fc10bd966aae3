#!/bin/bash
# Session m: barrier-free ordering + deep overlap -- parity, sanitizer, bench lines, launch list, ncu captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1m.log) 2>&1
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== memcheck (overlap + dense tests) ==="
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "overlapped or edge or empty" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_c2.json
echo "=== bench c2 --no-overlap ==="; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-overlap | tee gpurun_out/bench_c2_noovl.json
echo "=== bench c1 ==="; timeout 600 python bench.py --steps 50 --warmup 5 --workload c1 | tee gpurun_out/bench_c1.json
echo "=== bench reference ==="; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 | tee gpurun_out/bench_ref.json
echo "=== smoke ==="; timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== ncu launch list (c2) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "exit $?"
echo "=== ncu full (c2, c1) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o /tmp/prof_c2 -f python bench.py --steps 4 --warmup 3 --no-cpu --no-overlap > gpurun_out/ncu_full_c2.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/prof_c2.ncu-rep gpurun_out/ncu_full_c2_summary.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o /tmp/prof_c1 -f python bench.py --steps 4 --warmup 3 --no-cpu --no-overlap --workload c1 > gpurun_out/ncu_full_c1.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/prof_c1.ncu-rep gpurun_out/ncu_full_c1_summary.csv
cp /tmp/prof_c2.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out
