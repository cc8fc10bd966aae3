#!/bin/bash
# Round 2, session k: whole GPU suite (sibling shims included), launch shapes at 1 GiB per launch, c3 / c3wm order, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02k.log) 2>&1
nvidia-smi -L
echo "=== pytest -m gpu (all) ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
rm -f gpurun_out/ab.csv
echo "=== 1 GiB per launch: two CTAs per SM (default) / one ==="
timeout 600 python scripts/ab.py c1,c2 20 1024 | grep -v "^$"
AB_OPTS='{"force_ctas": 1}' timeout 600 python scripts/ab.py c1,c2 20 1024 | grep -v "^$"
echo "=== 512 MiB ==="
timeout 600 python scripts/ab.py c1,c2 20 512 | grep -v "^$"
AB_OPTS='{"force_ctas": 1}' timeout 600 python scripts/ab.py c1,c2 20 512 | grep -v "^$"
rm -f gpurun_out/probe_warps.csv
echo "=== c3wm before c3 ==="
PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c3wm,c3,c3wm 60
echo "=== bench (default) ==="; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; tail -c 300 gpurun_out/bench_default.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/bench_default.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "isolated us", round(v["roofline"]["kernel_ms_isolated_launch"]*1e3,1), "e2e", round(v["e2e"]["value"],1), "cpu", round(v["cpu_baseline"]["value"],3))
for l in d.get("north_star_legs",[]): print("   big", l["workload"], l["text_bytes_per_gpu"], round(l["value"]), "frac", round(l["roofline"]["frac"],3), l["kernel"]["threads"], l["kernel"]["ctas_per_sm"])
P
