#!/bin/bash
# Round 2, session ab: configs[2] -- what the access-policy window should cover now that candidates are rare
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02ab.log) 2>&1
for w in all front all front; do
ACWM_L2_WINDOW=$w ACWM_DEBUG_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-big-legs --no-cpu --workload c3 --text-mib 954 > gpurun_out/bench_l2$w.json 2> gpurun_out/bench_l2$w.err; echo "window=$w exit $?"
grep -h "L2 window" gpurun_out/bench_l2$w.err | sort | uniq -c | head -2
python - $w <<'P'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_l2%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print("   value", round(d["value"]), "us", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3))
P
done
