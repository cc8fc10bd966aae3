#!/bin/bash
# Round 2, session v: the evidence run of the final build -- whole GPU suite, step times of every BASELINE config, the launch list
# of the bench command, ncu --set full summaries of the c1 / c2 / c3 kernels, both bench arms as the driver runs them
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02v.log) 2>&1
nproc; nvidia-smi -L
echo "=== pytest -m gpu (all) ==="; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
rm -f gpurun_out/probe_warps.csv
echo "=== step times (128 MiB per launch; us per step in overlap mode, plain launches; matches) ==="
PROBE_OPTS='[{}]' timeout 600 python scripts/probe_warps.py c1,c2,c2ac,c1wm,c3,c3wm,c4 100
echo "=== launch list of the bench command ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "exit $?"
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02v_launches.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4][:70]; agg[k][0]+=1; agg[k][1]+=float(r[-1].replace(",",""))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"{v[0]:5d} launches {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f} %  {k}")
P
cap() { # name workload
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:scan_kernel -s 5 -c 1 -o /tmp/prof_$1 -f python scripts/one_scan.py $2 2>&1 | tail -1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep gpurun_out/ncu_$1_summary.csv
}
cap c1_r02v c1
cap c2_r02v c2
cap c3_r02v c3
cap c4_r02v c4
echo "=== bench (reference arm) ==="; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; echo "exit $?"
echo "=== bench (ours, default) ==="; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "exit $?"; tail -c 300 gpurun_out/r02v_bench.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r02v_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k in ("value","h2d_bytes_per_step","leg","pinned_copy_GBps_per_rank")}, "clocks", d["clocks"])
for k,v in d["per_algo"].items(): print("   ", k, round(v["value"]), "us", round(v["ms_per_step"]*1e3,2), "frac", round(v["roofline"]["frac"],3), "isolated us", round(v["roofline"]["kernel_ms_isolated_launch"]*1e3,1), "e2e", round(v["e2e"]["value"],1), "cpu", round(v["cpu_baseline"]["value"],3))
for l in d.get("north_star_legs",[]): print("   big", l["workload"], l["text_bytes_per_gpu"], round(l["value"]), "frac", round(l["roofline"]["frac"],3), l["kernel"]["threads"], l["kernel"]["ctas_per_sm"])
r=json.loads([l for l in open("gpurun_out/r02v_bench_ref.json") if l.startswith("{")][-1])
print("reference arm", round(r["value"],3), {k:round(v["value"],3) for k,v in r["per_algo"].items()})
P
