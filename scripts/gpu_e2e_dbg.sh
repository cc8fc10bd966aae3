#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, '.')
os.environ["ACWM_DEBUG_TIMING"] = "1"
import torch, acwm_pkg, bench
acwm = acwm_pkg.load(); dg = acwm_pkg.submodule("datagen")
n = 128 << 20
text = dg.text_host(n, 4, 1)
pats, _ = bench.make_patterns(dg, text, "c2")
mt = acwm.Matcher(acwm.WM, pats, 4); mt.upload(0, 1 << 23)
pinned = [torch.from_numpy(dg.text_host(n, 4, 1 + k)).pin_memory() for k in range(2)]
for frac in (1.0, 0.5):
    m = int(n * frac)
    for i in range(3):
        t0 = time.perf_counter(); c, pos = mt.search_host(pinned[i % 2][:m], cap=1 << 23); dt = time.perf_counter() - t0
        print(f"n={m>>20} MiB: search_host {dt*1e3:.3f} ms = {m/dt/1e9:.1f} GB/s count {c}", flush=True)
PY
