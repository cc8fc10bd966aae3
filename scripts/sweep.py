"""BASELINE configs[4]: pattern count 10..100k x length 8..64, AC and WM, on a 128 MiB DNA text (one GPU).

For every point: back-to-back step time (bench.py's timed loop: 4 texts cycled, overlap mode), text GB/s, the
kernel the table compiler chose, the match count checked against the oracle on a 16 MiB prefix, and -- where the
reference's preprocessing finishes in reasonable time (p <= 1000) -- the unmodified reference on all host cores
on the same prefix.  Rows -> gpurun_out/sweep.csv (+ sweep.md).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import acwm_pkg
import oracle

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
torch.cuda.set_device(0)
N = 128 << 20
CHECK = 16 << 20
STEPS = 50
ps = [int(x) for x in os.environ.get("SWEEP_P", "10,100,1000,10000,100000").split(",")]
ms = [int(x) for x in os.environ.get("SWEEP_M", "8,16,32,64").split(",")]
algos = os.environ.get("SWEEP_ALGO", "AC,WM").split(",")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "sweep.csv"), "w")
md = open(os.path.join(ROOT, "gpurun_out", "sweep.md"), "w")
cores = os.cpu_count() or 1


def log(*a):
    s = ",".join(str(x) for x in a)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


text0 = dg.text_host(N, 4, 1)
bufs = [dg.text_device(N, 4, 100 + k) for k in range(4)]
bufs[0].copy_(torch.from_numpy(text0))
st = torch.cuda.current_stream().cuda_stream
log("algo,p,m,us_per_step,GBps,frac_of_measured_peak,kernel,table_bytes,count_128MiB,oracle_check_16MiB,ref_cpu_GBps")
md.write(f"| algo | p | m | us / 128 MiB | text GB/s | kernel (stride, depth, tables) | matches | reference CPU GB/s ({cores} threads) |\n|---|---|---|---|---|---|---|---|\n")
peak = 6420.7
try:
    import json
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for algo_name in algos:
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    for p in ps:
        for m in ms:
            pats = dg.patterns_with_hits(text0, p, m, 4, 2)
            t0 = time.perf_counter()
            try:
                mt = acwm.Matcher(algo, pats, 4)
            except acwm.AcwmError as e:
                log(algo_name, p, m, "build-failed", e.code)
                continue
            build_s = time.perf_counter() - t0
            inf = mt.info
            mt.upload(0, N // 4)
            mt.set_overlap(True)
            res = []
            for rep in range(2):
                for i in range(8):
                    mt.scan_tensor(bufs[i % 4])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(STEPS):
                    mt.scan_tensor(bufs[i % 4])
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / STEPS * 1e3)
            mt.set_overlap(False)
            us = min(res)
            mt.scan_tensor(bufs[0])
            cnt, _, _ = mt.fetch(cap=0, stream=st)
            # parity on a prefix: count and positions against the oracle
            mt.scan_tensor(bufs[0][:CHECK])
            c16, pos16, _ = mt.fetch(cap=N // 4, stream=st, allow_overflow=True)
            ref = oracle.set_search(pats, text0[:CHECK])
            ok = c16 == ref["count"] and (c16 > N // 4 or np.array_equal(pos16, ref["positions"]))
            ref_gbps = ""
            if p <= 1000 and oracle.ref_available():
                r = (oracle.ref_ac if algo_name == "AC" else oracle.ref_wu)(pats, 4, text0[:CHECK], threads=cores)
                ref_gbps = f"{CHECK / r['search_s'] / 1e9:.3f}"
                ok = ok and r["count"] == ref["count"]
            gbps = N / us / 1e3
            kern = f"s{inf['stride']} d{inf['depth']} {'exact' if inf['exact_front'] else 'verify'} {'smem' if inf['table_in_smem'] else 'L2'}"
            log(algo_name, p, m, f"{us:.1f}", f"{gbps:.0f}", f"{gbps / peak:.3f}", kern.replace(",", ";"), inf["table_bytes"], cnt,
                "ok" if ok else "MISMATCH", ref_gbps)
            md.write(f"| {algo_name} | {p} | {m} | {us:.1f} | {gbps:.0f} | {kern}, {inf['table_bytes'] >> 10} KiB | {cnt} | {ref_gbps or 'n/a'} |\n")
            md.flush()
            mt.close()
