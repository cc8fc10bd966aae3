"""Back-to-back step time of the scan kernel for A/B runs (one process per ACWM_TUNE value; the
library reads the variable once).  Same timed loop as bench.py (4 texts cycled, CUDA events around K
steps, overlap mode on/off), no e2e / profiling / CPU legs.

    ACWM_TUNE=3 python scripts/ab.py c2,c1 [steps] [text_mib] -> rows appended to gpurun_out/ab.csv
"""
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
wls = (sys.argv[1] if len(sys.argv) > 1 else "c2,c1").split(",")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mib = int(sys.argv[3]) if len(sys.argv) > 3 else 128
opts = json.loads(os.environ.get("AB_OPTS", "{}"))
tune = os.environ.get("ACWM_TUNE", "default")
torch.cuda.set_device(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "ab.csv"), "a")


def log(*a):
    s = ",".join(str(x) for x in a)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


n = mib << 20
texts = {}
for wl in wls:
    algo_name, alphabet, p, m, _ = bench.WORKLOADS[wl]
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    text0 = dg.text_host(128 << 20, alphabet, bench.TEXT_SEED)
    pats, _ = bench.make_patterns(dg, text0, wl)
    if alphabet not in texts:
        texts.clear()
        texts[alphabet] = [dg.text_device(n, alphabet, 100 + k) for k in range(4)]
        texts[alphabet][0][: min(n, 128 << 20)].copy_(torch.from_numpy(text0)[: min(n, 128 << 20)])
    bufs = texts[alphabet]
    mt = acwm.Matcher(algo, pats, alphabet, **opts)
    mt.upload(0, max(1 << 20, n // 16))
    st = torch.cuda.current_stream().cuda_stream
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
        mt.set_overlap(False)
        cnt, _, _ = mt.fetch(cap=0, stream=st)
        us = float(np.min(res))
        log(wl, mib, f"tune={tune}", json.dumps(opts).replace(",", ";"), "overlap" if overlap else "coop",
            f"{us:.2f}", f"{np.median(res):.2f}", f"{n / us / 1e3:.1f}", cnt)
    mt.close()
