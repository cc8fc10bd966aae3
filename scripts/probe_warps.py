"""Step time of the scan kernel against the number of warps per CTA (one CTA per SM): how much of the scan
rate survives with 12-16 warps, the budget of a CTA that shares its SM with the next launch's CTA.
    python scripts/probe_warps.py c1,c2 [steps] -> gpurun_out/probe_warps.csv"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import acwm_pkg
import bench

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
wls = (sys.argv[1] if len(sys.argv) > 1 else "c1,c2").split(",")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mib = 128
torch.cuda.set_device(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "probe_warps.csv"), "a")
n = mib << 20
texts = {}
for wl in wls:
    algo_name, alphabet, p, m, _ = bench.WORKLOADS[wl]
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    text0 = dg.text_host(n, alphabet, bench.TEXT_SEED)
    pats, _ = bench.make_patterns(dg, text0, wl)
    if alphabet not in texts:
        texts.clear()
        texts[alphabet] = [dg.text_device(n, alphabet, 100 + k) for k in range(4)]
        texts[alphabet][0].copy_(torch.from_numpy(text0))
    bufs = texts[alphabet]
    for opts in json.loads(os.environ.get("PROBE_OPTS", '[{}, {"force_threads": 768, "force_stages": 1}, '
                                          '{"force_threads": 512, "force_stages": 1}, {"force_threads": 384, "force_stages": 1}, '
                                          '{"force_threads": 512, "force_stages": 2}, {"force_threads": 384, "force_stages": 2}]')):
        try:
            mt = acwm.Matcher(algo, pats, alphabet, **opts)
        except acwm.AcwmError as e:
            print(wl, opts, "build failed", e, flush=True)
            continue
        mt.upload(0, max(1 << 20, n // 16))
        st = torch.cuda.current_stream().cuda_stream
        row = [wl, json.dumps(opts).replace(",", ";"), mt.info["threads"], mt.info["stages"], mt.info["smem_bytes"]]
        for overlap in (True, False):
            mt.set_overlap(overlap)
            res = []
            for rep in range(3):
                for i in range(10):
                    mt.scan_tensor(bufs[i % 4])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    mt.scan_tensor(bufs[i % 4])
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / steps * 1e3)
            row.append(f"{min(res):.2f}")
        mt.set_overlap(False)
        cnt, _, _ = mt.fetch(cap=0, stream=st)
        row.append(cnt)
        s = ",".join(str(x) for x in row)
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()
        mt.close()
