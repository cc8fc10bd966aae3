#!/bin/bash
# Session o: final evidence of the round (final build)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1o.log) 2>&1
nproc; grep -m1 "model name" /proc/cpuinfo
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== memcheck (host-packed, overlap, dense) ==="
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "overlapped or edge or empty or shims" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "=== A/B rows ==="; rm -f gpurun_out/ab.csv
timeout 600 python scripts/ab.py c2,c1,c2ac,c1wm,c4,c3wm,c3 100 128 2>&1 | grep -v Warning
echo "=== bench c2 ==="; timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_c2.json
echo "=== bench c2 (raw H2D) ==="; ACWM_HOST_PACK=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu | tee gpurun_out/bench_c2_rawh2d.json
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
ls -la gpurun_out | head -40
echo "=== bench c4 (2e9 bytes) ==="
timeout 900 python bench.py --workload c4 --text-mib 1908 --steps 10 --warmup 3 --no-cpu | tee gpurun_out/bench_c4_full.json
