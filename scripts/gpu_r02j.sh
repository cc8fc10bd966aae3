#!/bin/bash
# Round 2, session j (4 GPUs): e2e through acwm_search_host under torchrun -- adaptive raw share of the hybrid host transfer
# against the plain one-byte-per-symbol copy; multi-GPU tests on real peers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02j.log) 2>&1
N=$(nvidia-smi -L | wc -l); nproc; echo "$N GPUs"
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-big-legs > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; echo "$tag exit $?"
  python - <<P
import json
d=json.load(open("gpurun_out/r02j_$tag.json"))
print("$tag", "value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","per_gpu_text_GBps_this_rank","per_gpu_link_GBps_this_rank","pinned_copy_GBps_per_rank","leg")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), "h2d", v["e2e"]["h2d_bytes_per_step"])
P
  grep -h "hybrid:" gpurun_out/r02j_$tag.err | tail -4
}
run adaptive ACWM_DEBUG_TIMING=1
run raw ACWM_HOST_PACK=0
run packall ACWM_HOST_RAW_PERCENT=0
echo "=== multi-GPU tests ==="
timeout 900 python -m pytest tests -m gpu -x -q -k "sharded or peers or device_sharded" 2>&1 | tail -5
