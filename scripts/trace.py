"""Timeline of the scan kernel from its own %globaltimer stamps (acwm_set_trace): where a step's time goes.

    python scripts/trace.py c2 [overlap|coop] -> gpurun_out/trace_<wl>_<mode>.txt (+ .npy of the raw stamps)
"""
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
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "overlap"
torch.cuda.set_device(0)
algo_name, alphabet, p, m, _ = bench.WORKLOADS[wl]
n = 128 << 20
text0 = dg.text_host(n, alphabet, bench.TEXT_SEED)
pats, _ = bench.make_patterns(dg, text0, wl)
bufs = [dg.text_device(n, alphabet, 100 + k) for k in range(4)]
mt = acwm.Matcher(acwm.AC if algo_name == "AC" else acwm.WM, pats, alphabet)
mt.upload(0, max(1 << 20, n // 16))
W = int(acwm.lib().acwm_trace_words_per_cta())
G = 148
NL = 6  # traced launches
traces = [torch.zeros(256 * W, dtype=torch.int64, device="cuda") for _ in range(NL)]
mt.set_overlap(mode == "overlap")
for i in range(10):
    mt.scan_tensor(bufs[i % 4])
for k in range(NL):
    mt.set_trace(traces[k].data_ptr())
    mt.scan_tensor(bufs[k % 4])
mt.set_trace(None)
torch.cuda.synchronize()
mt.set_overlap(False)
T = np.stack([t.cpu().numpy().reshape(256, W)[:G] for t in traces]).astype(np.float64)  # [launch, cta, word]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{wl}_{mode}.npy"), T)
out = open(os.path.join(ROOT, "gpurun_out", f"trace_{wl}_{mode}.txt"), "w")


def log(s=""):
    print(s)
    out.write(s + "\n")


names = ["entry", "first copies issued", "CTA set up", "tables ready (warp 0)", "all warps done (sync)",
         "span total published", "arrived", "look-back done", "exit"]
t0 = T[0, :, 0].min()
log(f"# {wl} {mode}: {NL} consecutive launches, 148 CTAs; times in us relative to the first CTA entry of launch 0")
for k in range(NL):
    L = T[k]
    log(f"launch {k}: entry min/med/max {(L[:,0].min()-t0)/1e3:8.2f} {(np.median(L[:,0])-t0)/1e3:8.2f} {(L[:,0].max()-t0)/1e3:8.2f}"
        f" | exit min/med/max {(L[:,8].min()-t0)/1e3:8.2f} {(np.median(L[:,8])-t0)/1e3:8.2f} {(L[:,8].max()-t0)/1e3:8.2f}"
        f" | span {(L[:,8].max()-L[:,0].min())/1e3:6.2f}")
log()
log("per-CTA phase durations (us), median / p10 / p90 / max over CTAs, averaged over launches 2..")
for a_, b_ in ((0, 1), (1, 2), (2, 3), (0, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (0, 8)):
    d = (T[2:, :, b_] - T[2:, :, a_]) / 1e3
    log(f"  {names[a_]:>28s} -> {names[b_]:<28s} med {np.median(d):7.2f}  p10 {np.percentile(d,10):7.2f}  p90 {np.percentile(d,90):7.2f}  max {d.max():7.2f}")
log(f"  cost of one stamp (two back-to-back pairs): {np.median(T[2:, :, 9] - T[2:, :, 2])/1e3:.2f} {np.median(T[2:, :, 10] - T[2:, :, 9])/1e3:.2f}")
first = (T[2:, :, 16:48] - T[2:, :, 0:1]) / 1e3
end = (T[2:, :, 48:80] - T[2:, :, 0:1]) / 1e3
cta_scan_end = T[2:, :, 4] - T[2:, :, 0]
log(f"  CTA scan time (entry -> all warps done): med {np.median(cta_scan_end)/1e3:6.2f} min {cta_scan_end.min()/1e3:6.2f} max {cta_scan_end.max()/1e3:6.2f}")
gap = (T[3:, :, 0].min(1) - T[2:-1, :, 8].max(1)) / 1e3
log(f"  next launch's first entry minus this launch's last exit: {np.round(gap, 2)}")
step = np.diff(T[:, :, 8].max(1)) / 1e3
log(f"  last-exit to last-exit (step time): {np.round(step, 2)}")
# who shares an SM: for every SM the [entry, exit] intervals of the CTAs of all traced launches, in time order
sm = T[:, :, 11].astype(np.int64)
log()
log("per-SM occupancy (first 4 SMs): launch:cta [entry .. scan done .. exit] in us")
for s_ in sorted(set(sm.reshape(-1).tolist()))[:4]:
    rows = sorted((T[k, c, 0], k, c) for k in range(NL) for c in range(G) if sm[k, c] == s_)
    log(f"  SM {s_}: " + "  ".join(f"{k}:{c} [{(T[k,c,0]-t0)/1e3:.1f} .. {(T[k,c,4]-t0)/1e3:.1f} .. {(T[k,c,8]-t0)/1e3:.1f}]" for _, k, c in rows))
conc = []
for k in range(1, NL):
    for c in range(G):
        others = [(kk, cc) for kk in range(NL) for cc in range(G) if (kk, cc) != (k, c) and sm[kk, cc] == sm[k, c]
                  and T[kk, cc, 0] < T[k, c, 8] and T[kk, cc, 8] > T[k, c, 0]]
        conc.append(len(others))
log(f"  CTAs of other traced launches that overlap a CTA's lifetime on its SM: mean {np.mean(conc):.2f} max {max(conc)}")
inflight = []
for k in range(NL):
    lo, hi = T[k, :, 0].min(), T[k, :, 8].max()
    inflight.append(sum(1 for kk in range(NL) if T[kk, :, 0].min() < hi and T[kk, :, 8].max() > lo))
log(f"  launches in flight during each launch (incl. itself): {inflight}")
