"""A handful of plain scans of one workload (for ncu captures): python scripts/one_scan.py c2 ['{"force_ctas": 1}'] [n_scans]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import acwm_pkg
import bench

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
opts = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
torch.cuda.set_device(0)
algo_name, alphabet, p, m, _ = bench.WORKLOADS[wl]
n = 128 << 20
text0 = dg.text_host(n, alphabet, bench.TEXT_SEED)
pats, _ = bench.make_patterns(dg, text0, wl)
bufs = [dg.text_device(n, alphabet, 100 + k) for k in range(4)]
mt = acwm.Matcher(acwm.AC if algo_name == "AC" else acwm.WM, pats, alphabet, **opts)
mt.upload(0, max(1 << 20, n // 16))
for i in range(reps):
    mt.scan_tensor(bufs[i % 4])
torch.cuda.synchronize()
c, _, _ = mt.fetch(cap=0, stream=torch.cuda.current_stream().cuda_stream)
print(wl, opts, mt.info, "count", c)
