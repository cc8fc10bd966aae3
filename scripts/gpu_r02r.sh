#!/bin/bash
# Round 2, session r: e2e leg after the share controller was stabilised; host-path tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02r.log) 2>&1
nvidia-smi -L
for impl in avx512 bmi2; do
ACWM_PACK_IMPL=$impl ACWM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-big-legs --no-cpu > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err; echo "$impl exit $?"
grep -h "hybrid:" gpurun_out/bench_$impl.err | awk "{print \$2,\$7,\$NF}" | tr "\n" ";" | cut -c1-700; echo
grep -h "acwm host-packed search" gpurun_out/bench_$impl.err | tail -2
python - $impl <<'P'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
for k,v in d["per_algo"].items(): print("   ", sys.argv[1], k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "e2e", round(v["e2e"]["value"],1), v["e2e"]["h2d_bytes_per_step"])
P
done
echo "=== host-path tests ==="; timeout 1200 python -m pytest tests -m gpu -x -q -k "host or hybrid or packed or sharded or overflow" 2>&1 | tail -3
