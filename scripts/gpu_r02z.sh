#!/bin/bash
# Round 2, session z: last regression of the committed build -- smoke, whole GPU suite, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02z.log) 2>&1
nvidia-smi -L
echo "=== smoke ==="; timeout 600 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?"
echo "=== pytest -m gpu (all) ==="; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench (ours, default) ==="; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "exit $?"; tail -c 300 gpurun_out/r02z_bench.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r02z_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"],1))
for l in d.get("north_star_legs",[]): print("   big", l["workload"], l["text_bytes_per_gpu"], round(l["value"]), "frac", round(l["roofline"]["frac"],3))
P
