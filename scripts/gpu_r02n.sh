#!/bin/bash
# Round 2, session n: bytes path with one raw slot per warp and 20-24 warps (BASELINE configs[3]); pipelined fetch of long
# position lists (e2e of the AC leg)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02n.log) 2>&1
nvidia-smi -L
echo "=== parity (whole suite) ==="; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
rm -f gpurun_out/probe_warps.csv
echo "=== c4 by launch shape ==="
PROBE_OPTS='[{}, {"force_threads": 640, "force_stages": 1}, {"force_threads": 512, "force_stages": 1}, {"force_threads": 384, "force_stages": 2}]' timeout 600 python scripts/probe_warps.py c4 60
echo "=== bench (default, no big legs) ==="; ACWM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-big-legs > gpurun_out/bench_nb.json 2> gpurun_out/bench_nb.err; echo "exit $?"
grep -h "positions after" gpurun_out/bench_nb.err | tail -3
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/bench_nb.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"],1), v["e2e"]["h2d_bytes_per_step"])
P
