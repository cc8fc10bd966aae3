#!/bin/bash
# Round 2, session l: sibling shims on the GPU, per-launch shape for long scans, host transfer choice at N = 1, c3 against c3wm under ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02l.log) 2>&1
nvidia-smi -L
echo "=== tests ==="; timeout 1800 python -m pytest tests/test_siblings.py tests/test_c_dropin.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
rm -f gpurun_out/ab.csv
echo "=== c1 / c2 at 128, 256, 1024 MiB per launch (shape picked per launch) ==="
for mib in 128 256 1024; do timeout 600 python scripts/ab.py c1,c2 30 $mib | grep overlap; done
cap() { # name workload
  timeout 900 ncu --set full --clock-control none -k regex:scan_kernel -s 5 -c 1 -o /tmp/prof_$1 -f python scripts/one_scan.py $2 2>&1 | tail -1
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep gpurun_out/ncu_$1_summary.csv
}
cap c3_r02l c3
cap c3wm_r02l c3wm
paste -d, gpurun_out/ncu_c3_r02l_summary.csv gpurun_out/ncu_c3wm_r02l_summary.csv | cut -d, -f1,3,6 | head -45
echo "=== bench (default, no big legs) ==="; ACWM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-big-legs > gpurun_out/bench_nb.json 2> gpurun_out/bench_nb.err; echo "exit $?"
grep -h "hybrid:" gpurun_out/bench_nb.err | tail -3
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/bench_nb.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"],1), v["e2e"]["h2d_bytes_per_step"])
P
