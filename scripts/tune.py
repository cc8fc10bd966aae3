"""Kernel-time sweep on one GPU: workloads x thread counts x text sizes -> gpurun_out/tune.csv."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import acwm_pkg
acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
import bench

out = open(os.path.join(ROOT, "gpurun_out", "tune.csv"), "a")
def log(*a):
    s = ",".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()

torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in os.environ.get("TUNE_MIB", "128,1024").split(",")]
wls = os.environ.get("TUNE_WL", "c2,c1,c2ac,c1wm").split(",")
def V(t, st, **kw):
    d = dict(force_threads=t, force_stages=st)
    d.update(kw)
    return d

variants = {  # workload -> list of option dicts ({} = what the builder picks)
    "c2": [{}, V(768, 1), V(512, 1)],
    "c1": [{}, V(768, 1), V(512, 1)],
    "c2ac": [{}],
    "c1wm": [{}],
    "c3": [{}, dict(force_stride=1), dict(force_smem_tables=1)],
    "ac10k16": [{}, dict(force_smem_tables=1)],
    "c3wm": [{}, dict(force_smem_tables=1), V(768, 1), V(512, 1), dict(force_stride=8)],
    "c4": [{}, dict(force_smem_tables=1), V(384, 2), V(256, 2), dict(force_stride=2)],
}
log("workload,text_mib,opts,stride,depth,exact,threads,smem,scan_us,finalize_us,GBps,frac_measured,count")
texts = {}
for wl in wls:
    algo_name, alphabet, p, m, desc = bench.WORKLOADS[wl]
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    text0 = dg.text_host(128 << 20, alphabet, bench.TEXT_SEED)
    pats, m_max = bench.make_patterns(dg, text0, wl)
    for mib in sizes:
        key = (alphabet, mib)
        if key not in texts:
            texts.clear()
            nrot = 4 if mib <= 256 else 2
            texts[key] = [dg.text_device(mib << 20, alphabet, 100 + k) for k in range(nrot)]
            texts[key][0][: 128 << 20].copy_(torch.from_numpy(text0)[: min(mib, 128) << 20])
        bufs = texts[key]
        for opts in variants[wl]:
            try:
                mt = acwm.Matcher(algo, pats, alphabet, **opts)
            except acwm.AcwmError as e:
                log(wl, mib, json.dumps(opts).replace(",", ";"), "build-failed", e.code)
                continue
            inf = mt.info
            mt.upload(pos_capacity=max(1 << 22, (mib << 20) // 16))
            mt.set_profiling(True)
            ss, ff, cnt = [], [], 0
            for i in range(12):
                mt.scan_tensor(bufs[i % len(bufs)])
                a, b = mt.profiled_seconds()
                cnt, _, _ = mt.fetch(cap=0, stream=st)
                if i >= 4:
                    ss.append(a); ff.append(b)
            scan = float(np.median(ss)); fin = float(np.median(ff))
            gbps = (mib << 20) / scan / 1e9
            log(wl, mib, json.dumps(opts).replace(",", ";"), inf["stride"], inf["depth"], inf["exact_front"], f'{inf["threads"]}x{inf["stages"]}',
                f'{inf["smem_bytes"]}/l2={1 - inf["table_in_smem"]}', f"{scan*1e6:.1f}", f"{fin*1e6:.1f}", f"{gbps:.1f}", f"{gbps/6543.1:.3f}", cnt)
            mt.close()
