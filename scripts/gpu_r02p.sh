#!/bin/bash
# Round 2, session p (N GPUs of one box): the bench as the driver's scaling run launches it (default workload, big legs
# included: 1 GiB per GPU for c1 / c2, configs[2] at 10^9 B per GPU), configs[2] at its full 8e9 B, sharded parity with the
# in-kernel count exchange, the multi-GPU tests on real peers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
exec > >(tee gpurun_out/r02p_n$N.log) 2>&1
nproc; free -g | head -2; echo "$N GPUs"
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2> gpurun_out/stderr_$1.log; echo "exit $?" >&2; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/stderr_$1.log | tail -3 >&2; }
summ() { python - "$1" <<'P'
import json, sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", round(d["value"]), "n_gpus", d["n_gpus"], "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank","pinned_copy_GBps_all_ranks")})
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"],1), v["e2e"]["h2d_bytes_per_step"], v["count_exchange"][:50])
for l in d.get("north_star_legs",[]): print("   big", l["workload"], l["text_bytes_per_gpu"], round(l["value"]), "frac", round(l["roofline"]["frac"],3))
P
}
echo "=== bench --gpus $N (default) ==="; run 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02p_bench_n$N.json; summ gpurun_out/r02p_bench_n$N.json
echo "=== bench c3 --gpus $N (954 MiB per GPU: 8e9 B over 8) ==="; run 29542 bench.py --gpus $N --workload c3 --text-mib 954 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02p_c3_n$N.json; summ gpurun_out/r02p_c3_n$N.json
echo "=== sharded parity ==="; run 29544 scripts/sharded_parity.py
echo "=== multi-GPU tests ==="; timeout 900 python -m pytest tests -m gpu -x -q -k "sharded or peers or device_sharded" 2>&1 | tail -3
