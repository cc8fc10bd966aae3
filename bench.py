#!/usr/bin/env python
"""Bench harness for the AC / WM scan path (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c1|...] [--text-mib M]

One JSON line on rank 0.  A "step" is one pass of the hot path over one batch of
synthetic text: default workload = BASELINE.json configs[1] (Wu-Manber, 128 MiB 4-symbol
DNA text, 1000 patterns of m = 16), per GPU (weak scaling: every rank scans its own
128 MiB shard + (m-1)-byte halo, counts are all-reduced over NCCL).

 value   text GB/s, text resident in HBM, CUDA events on the launching stream, max over ranks
 e2e     same metric through acwm_search_host with the text in PINNED HOST memory:
         H2D of the text and D2H of count + positions inside the timed region
 roofline  the scan kernel (CUDA events inside acwm_scan_device), algorithmic bytes =
         1 B per text symbol + 8 B per reported position, vs the measured HBM peak
 cpu_baseline  the unmodified reference search_wu / search_ac (oracle/_ref) on all host cores
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (algo, alphabet, p, m, description)
    "c2": ("WM", 4, 1000, 16, "BASELINE configs[1]: Wu-Manber, 4-symbol DNA text, 1000 patterns m=16"),
    "c1": ("AC", 4, 100, 8, "BASELINE configs[0]: Aho-Corasick, 4-symbol DNA text, 100 patterns m=8"),
    "c2ac": ("AC", 4, 1000, 16, "Aho-Corasick on the configs[1] pattern set (1000 patterns m=16)"),
    "c1wm": ("WM", 4, 100, 8, "Wu-Manber on the configs[0] pattern set (100 patterns m=8)"),
    "c3": ("AC", 4, 100000, 32, "BASELINE configs[2] per-GPU shard: Aho-Corasick, DNA, 100000 patterns m=32"),
    "c3wm": ("WM", 4, 100000, 32, "Wu-Manber on the configs[2] pattern set"),
    "ac10k16": ("AC", 4, 10000, 16, "Aho-Corasick, DNA, 10000 patterns m=16 (sweep point)"),
    "c4": ("WM", 256, 10000, (8, 64), "BASELINE configs[3]: Wu-Manber, 256-symbol text, 10000 patterns m=8..64"),
}
TEXT_SEED, PAT_SEED = 1, 2
N_ROTATE = 4  # distinct text buffers per rank, cycled so that no step finds its text in L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """NVML clocks + throttle reasons every ~5 ms while the GPU is under load."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.stop_flag, self.ok = [], False, False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), clk, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self, t0, t1):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if len(inside) < 3:
            inside, window = self.samples, "warm-up + timed + profiling loops (timed region < 3 samples)"
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        seen = set()
        for _, _, rs in inside:
            for k, bit in names.items():
                if rs & bit:
                    seen.add(k)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": float(self.max_sm),
                "reasons": sorted(seen), "samples": len(inside), "window": window}


def make_patterns(dg, text_np, wl):
    algo, alphabet, p, m, _ = WORKLOADS[wl]
    if isinstance(m, tuple):
        return dg.mixed_patterns_with_hits(text_np, p, m[0], m[1], alphabet, PAT_SEED), m[1]
    return dg.patterns_with_hits(text_np, p, m, alphabet, PAT_SEED), m


# ============================================================== reference arm (CPU)
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref = unmodified ac.c / wu.c compiled
    from /root/reference; else the oracle port) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    import acwm_pkg
    dg = acwm_pkg.submodule("datagen")
    algo, alphabet, p, m, desc = WORKLOADS[args.workload]
    n = args.text_mib << 20
    cores = os.cpu_count() or 1
    text = dg.text_host(n, alphabet, TEXT_SEED)
    pats, m_max = make_patterns(dg, text, args.workload)
    mixed = isinstance(m, tuple)
    use_ref = oracle.ref_available() and not mixed and not (algo == "AC" and p * m_max > 400_000)

    def one_pass(sample):
        t0 = time.perf_counter()
        if use_ref:
            r = (oracle.ref_ac if algo == "AC" else oracle.ref_wu)(pats, alphabet, sample, threads=cores)
            secs, cnt = r["search_s"], r["count"]
        else:  # port: single scalar thread
            r = oracle.set_search(pats, sample, want_positions=False)
            secs, cnt = time.perf_counter() - t0, r["count"]
        return secs, cnt

    # bounded sample: calibrate on 8 MiB, then size each step for a whole run of ~2 minutes
    cal_secs, _ = one_pass(text[: min(n, 8 << 20)])
    rate = min(n, 8 << 20) / max(cal_secs, 1e-6)
    budget = 120.0 / max(1, args.steps + args.warmup)
    sample_n = int(min(n, max(1 << 20, rate * budget)))
    sample = text[:sample_n]
    for _ in range(args.warmup):
        one_pass(sample)
    tot = 0.0
    for _ in range(args.steps):
        s, cnt = one_pass(sample)
        tot += s
    value = sample_n * args.steps / tot / 1e9
    kind = "reference" if use_ref else "port"
    line = {
        "impl": "reference", "metric": "text GB/s scanned", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "text_bytes_per_gpu": n, "algo": algo, "patterns": p,
                   "m": m_max if not mixed else list(m), "alphabet": alphabet},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores if use_ref else 1, "kind": kind,
                         "sample": f"first {sample_n} bytes of the {n}-byte text per step "
                                   f"(search only, preprocessing excluded as in main.c:246-262)"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ============================================================== our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist

    import acwm_pkg
    acwm = acwm_pkg.load()
    dg = acwm_pkg.submodule("datagen")
    sh = acwm_pkg.submodule("sharding")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    algo_name, alphabet, p, m, desc = WORKLOADS[args.workload]
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    n = args.text_mib << 20  # per-GPU shard (weak scaling)

    # the pattern set is the same on every rank (replicated tables): "with hits" from rank 0's text
    text0 = dg.text_host(n, alphabet, TEXT_SEED)
    pats, m_max = make_patterns(dg, text0, args.workload)
    halo = m_max - 1
    mt = acwm.Matcher(algo, pats, alphabet, **json.loads(args.matcher_opts))
    pos_cap = max(1 << 20, n // 16)
    mt.upload(local_rank, pos_cap)

    # N_ROTATE distinct shards per rank.  Buffer 0 of rank r is shard r of the global text
    # = concatenation of the per-rank texts (seed TEXT_SEED + r), followed by the first
    # m_max-1 bytes of shard r+1 (the halo of main.c:467-477); the others only defeat L2.
    def shard_text(seed_base, k):
        own = dg.text_host(n, alphabet, seed_base + rank + 1000 * k)
        if rank + 1 < world and halo:
            nxt = dg.text_host(n, alphabet, seed_base + rank + 1 + 1000 * k)[:halo]
            own = np.concatenate([own, nxt])
        return own

    # texts much larger than L2 (126 MB) need no rotation: every step streams from HBM anyway
    N_ROTATE = globals()["N_ROTATE"] if n <= (256 << 20) else 1
    host_texts = [text0 if (rank == 0 and world == 1) else shard_text(TEXT_SEED, 0)]
    for k in range(1, N_ROTATE):
        host_texts.append(shard_text(TEXT_SEED, k))
    dev_texts = [torch.from_numpy(t).to(dev) for t in host_texts]
    report_from = halo if rank > 0 else 0
    stream = torch.cuda.current_stream().cuda_stream
    count_buf = torch.zeros(1, dtype=torch.int64, device=dev)
    d_count_ptr, _ = mt.result_device_ptrs()

    count_view = _device_count_tensor(torch, d_count_ptr, dev)  # the matcher's device-resident count

    # the only thing that crosses NVLink is the per-GPU match count (8 bytes): exchanged inside the scan
    # kernel through peer-mapped mailboxes; NCCL all_reduce is the fallback
    fused_exchange = world > 1 and not args.nccl_count and sh.connect_peers(mt, dev)

    def step(i, want_positions=True):
        mt.scan_tensor(dev_texts[i % N_ROTATE], want_positions=want_positions, report_from=report_from)
        if world > 1 and not fused_exchange:
            count_buf.copy_(count_view)
            sh.allreduce_count_tensor(count_buf)

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mt.set_overlap(not args.no_overlap)  # consecutive scans of resident texts: programmatic dependent launches
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = mt.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = mt.launch_count - launches0
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([elapsed_ms], dtype=torch.float64, device=dev))
        per_rank_ms = [float(g.item()) / args.steps for g in gathered]
        elapsed_ms = max(per_rank_ms) * args.steps  # max over ranks
    ms_per_step = elapsed_ms / args.steps
    text_bytes = sum(int(dev_texts[(args.warmup + i) % N_ROTATE].numel()) for i in range(args.steps)) / args.steps
    value = world * text_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- results of the last step (parity of the global count is a test, here it is reported)
    mt.set_overlap(False)
    last_count, last_pos, _ = mt.fetch(cap=pos_cap, stream=stream)
    global_count = sh.allreduce_count(last_count, dev)
    if fused_exchange:  # what the kernels exchanged must be what NCCL sums
        fused_global = mt.fetch_global_count(stream)
        assert fused_global == global_count, (fused_global, global_count)
        mt.set_peers(0, 0, None)  # the legs below are per-rank (profiling, e2e): no exchange

    # ---- roofline: the scan kernel alone, CUDA events on the launching stream
    mt.set_profiling(True)
    scan_s, fin_s, matches = [], [], []
    for i in range(max(args.steps, 8)):
        mt.scan_tensor(dev_texts[i % N_ROTATE], want_positions=True, report_from=report_from)
        a, b = mt.profiled_seconds()
        c, _, _ = mt.fetch(cap=0, stream=stream)
        scan_s.append(a)
        fin_s.append(b)
        matches.append(c)
    mt.set_profiling(False)
    scan_isolated = float(np.mean(scan_s))
    alg_bytes = float(np.mean([dev_texts[i % N_ROTATE].numel() + 8 * matches[i] for i in range(len(matches))]))
    peak, peak_src = peaks()
    # the kernel's average launch duration over the timed region: at N = 1 a step IS one launch of the scan
    # kernel and nothing else, so the timed region / K is that average (back-to-back launches); otherwise (a
    # collective per step) the event-bracketed single launches are used
    scan_mean = ms_per_step * 1e-3 if (world == 1 and launches == args.steps) else scan_isolated
    achieved = alg_bytes / scan_mean / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- end to end through the public API: pinned host text -> count + positions on the host
    pinned = [torch.from_numpy(t).pin_memory() for t in host_texts[:2]]
    if len(pinned) == 1:
        pinned = pinned * 2
    e2e_steps = max(3, min(args.steps, 10))
    pos_out = np.empty(pos_cap, np.uint64)  # the caller's position buffer, reused
    for i in range(2):
        mt.search_host(pinned[i % 2], out=pos_out)
    barrier()
    t0 = time.perf_counter()
    e2e_count = 0
    for i in range(e2e_steps):
        c, ppos = mt.search_host(pinned[i % 2], out=pos_out)
        if world > 1:
            c = sh.allreduce_count(c, dev)
        e2e_count = c
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_bytes = int(pinned[0].numel())
    e2e_value = world * e2e_bytes * e2e_steps / e2e_s / 1e9
    sampler.stop_flag = True
    sampler.join(timeout=1)
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- CPU baseline: the reference on the host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        mt.scan_tensor(dev_texts[0], want_positions=False)
        gpu_count0, _, _ = mt.fetch(cap=0, stream=stream)
        cpu = cpu_baseline(args, host_texts[0], pats, algo_name, alphabet, m, gpu_count0)

    if rank == 0:
        info = mt.info
        line = {
            "metric": "text GB/s scanned", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "algo": algo_name, "alphabet": alphabet,
                       "patterns": p, "m": list(m) if isinstance(m, tuple) else m,
                       "text_bytes_per_gpu": n, "halo_bytes": halo,
                       "l2": (f"{N_ROTATE} distinct {args.text_mib} MiB texts cycled (working set > L2)" if N_ROTATE > 1
                              else f"one {args.text_mib} MiB text per GPU (larger than L2)"),
                       "positions": "count + sorted uint64 positions produced every step",
                       "count_exchange": ("none (1 GPU)" if world == 1 else
                                          "in-kernel st.release.sys into NVLink peer mailboxes (torch symmetric memory)"
                                          if fused_exchange else "NCCL all_reduce of the 8-byte count per step"),
                       "launch": "one kernel per step (scan + position ordering + result), no memset / finalize nodes; "
                                 + ("cooperative launches" if args.no_overlap else
                                    "consecutive steps chained as programmatic dependent launches (acwm_set_overlap)"),
                       "kernel": {k: info[k] for k in ("packed2bit", "stride", "depth", "exact_front", "n_rows",
                                                        "table_in_smem", "smem_bytes", "threads", "stages")}},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(mt.last_h2d_bytes),
                    "text_bytes_per_step": e2e_bytes,
                    "h2d": ("hybrid: a prefix of the text copied one byte per symbol by DMA while the host cores pack the "
                            "rest to 2 bits per symbol (csrc/hostpack.cpp)" if mt.last_h2d_bytes < e2e_bytes
                            else "text copied one byte per symbol"),
                    "d2h_bytes_per_step": 8 * int(e2e_count if world == 1 else last_count) + 32,
                    "steps": e2e_steps, "api": "acwm_search_host (pinned host text -> host count + positions)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "scan_kernel (one launch: TMA-fed scan + barrier-free position ordering)",
                         "kernel_ms": scan_mean * 1e3, "kernel_ms_isolated_launch": scan_isolated * 1e3,
                         "duration_source": "timed region / K (one launch per step)" if scan_mean != scan_isolated
                         else "CUDA events around single launches",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "frac_of_8TBps_spec": achieved / 8000.0},
            "matches_last_step": int(global_count),
            "per_rank_ms_per_step": [round(x, 5) for x in per_rank_ms],
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _device_count_tensor(torch, ptr, dev):
    """int64[1] view of the matcher's device-resident count (no host round trip)."""
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(w, device=dev)


def cpu_baseline(args, text, pats, algo_name, alphabet, m, gpu_count):
    import oracle
    cores = os.cpu_count() or 1
    mixed = isinstance(m, tuple)
    n = text.size
    use_ref = oracle.ref_available() and not mixed and not (algo_name == "AC" and len(pats) * pats.shape[1] > 400_000)
    passes, tot, cnt = 0, 0.0, None
    sample = text
    t_start = time.perf_counter()
    while passes < 5 and time.perf_counter() - t_start < 12.0:
        if use_ref:
            r = (oracle.ref_ac if algo_name == "AC" else oracle.ref_wu)(pats, alphabet, sample, threads=cores)
            tot += r["search_s"]
            cnt = r["count"]
        else:
            t0 = time.perf_counter()
            cnt = oracle.set_search(pats, sample, want_positions=False)["count"]
            tot += time.perf_counter() - t0
        passes += 1
    out = {"value": n * passes / tot / 1e9, "unit": "GB/s", "cores": cores if use_ref else 1,
           "kind": "reference" if use_ref else "port",
           "sample": f"{passes} pass(es) over the full {n}-byte text of one step, search only "
                     f"(unmodified search_{'ac' if algo_name == 'AC' else 'wu'} on {cores} threads, MPI-rank shard "
                     f"geometry of main.c:467-477)" if use_ref else f"{passes} pass(es), oracle port, 1 thread",
           "count": int(cnt)}
    if gpu_count is not None:
        out["count_equals_gpu"] = bool(cnt == gpu_count)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--text-mib", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="plain cooperative launches in the timed loop")
    ap.add_argument("--matcher-opts", default="{}", help="JSON of acwm_options overrides (tuning experiments)")
    ap.add_argument("--nccl-count", action="store_true", help="all-reduce the count with NCCL instead of in-kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
