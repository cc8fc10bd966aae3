#!/bin/bash
# Round 2, session aa: e2e with one full-size CTA per SM for every scan of a matcher with a verification stage (ACWM_BIG_SHAPE=1)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02aa.log) 2>&1
for big in "" 1; do
ACWM_BIG_SHAPE=$big ACWM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-big-legs --no-cpu --workload c2 > gpurun_out/bench_big$big.json 2> gpurun_out/bench_big$big.err; echo "big='$big' exit $?"
grep -h "acwm host-packed search" gpurun_out/bench_big$big.err | tail -2 | cut -c60-
python - "$big" <<'P'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_big%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print("   value", round(d["value"]), "us", round(d["ms_per_step"]*1e3,2), "isolated", round(d["roofline"]["kernel_ms_isolated_launch"]*1e3,1), "e2e", round(d["e2e"]["value"],1))
P
done
