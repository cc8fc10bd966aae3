#!/bin/bash
# Round 2, session q: one full-size CTA per SM at 128 MiB per launch (c2, c2ac, c1), the e2e leg after the packer / share changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02q.log) 2>&1
nvidia-smi -L
rm -f gpurun_out/probe_warps.csv
echo "=== default shapes ==="; PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c1,c2,c2ac 100
echo "=== ACWM_BIG_SHAPE=1 ==="; ACWM_BIG_SHAPE=1 PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c1,c2,c2ac 100
echo "=== e2e: packer implementations ==="
for impl in avx512 bmi2; do
ACWM_PACK_IMPL=$impl ACWM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-big-legs --no-cpu > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err; echo "$impl exit $?"
grep -h "hybrid:" gpurun_out/bench_$impl.err | tail -2; grep -h "acwm host-packed search" gpurun_out/bench_$impl.err | tail -2
python - $impl <<'P'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
for k,v in d["per_algo"].items(): print("   ", sys.argv[1], k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "e2e", round(v["e2e"]["value"],1), v["e2e"]["h2d_bytes_per_step"])
P
done
