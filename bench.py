#!/usr/bin/env python
"""Bench harness for the AC / WM scan path (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c1+c2|c1|c2|...] [--text-mib M] [--no-big-legs]

One JSON line on rank 0.  BASELINE.json's metric is "text GB/s scanned (AC & WM)", so the default
run measures BOTH algorithms in one invocation: leg AC = BASELINE configs[0] (Aho-Corasick, 100
patterns of m = 8) and leg WM = configs[1] (Wu-Manber, 1000 patterns of m = 16), both on the same
128 MiB 4-symbol DNA text per GPU (weak scaling: every rank scans its own 128 MiB shard + (m-1)-byte
halo; only the 8-byte counts cross NVLink).  A "step" is one pass of one algorithm over one text.
Top-level `value` = the LOWER of the two legs (ms_per_step / roofline / e2e / cpu_baseline are that
leg's); `per_algo` carries every number of both.

 value   text GB/s, text resident in HBM, CUDA events on the launching stream, max over ranks
 e2e     same metric through acwm_search_host with the text in PINNED HOST memory:
         H2D of the text and D2H of count + positions inside the timed region
 roofline  the scan kernel: algorithmic bytes = 1 B per text symbol + 8 B per reported position,
         over the average launch duration in the timed region, vs the measured HBM peak
 cpu_baseline  the unmodified reference search_ac / search_wu (oracle/_ref) on all host cores
 north_star_legs  the same two algorithms on 1 GiB of DNA per GPU and BASELINE configs[2]
         (AC, 100 000 patterns of m = 32, 10^9 bytes per GPU): what the north star scales on
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
DEFAULT_WORKLOAD = "c1+c2"
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


def workload_config(workload, n):
    """The `config` object: names the workload, identical in both arms (ours / reference)."""
    legs = workload.split("+")
    algos = {}
    for wl in legs:
        algo, alphabet, p, m, desc = WORKLOADS[wl]
        algos[algo if len(legs) > 1 else wl] = {"workload": wl, "algo": algo, "patterns": p,
                                               "m": list(m) if isinstance(m, tuple) else m, "alphabet": alphabet,
                                               "what": desc}
    if workload == DEFAULT_WORKLOAD:
        name = (f"c1+c2: BASELINE configs[0] (AC, 100 patterns m=8) and configs[1] (WM, 1000 patterns m=16, B=3), "
                f"each over the same {n >> 20} MiB 4-symbol DNA text per GPU; value = the slower of the two")
    else:
        name = " + ".join(f"{wl}: {WORKLOADS[wl][4]}" for wl in legs)
    return {"workload": name, "text_bytes_per_gpu": n, "legs": algos}


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

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if any(t0 <= s[0] <= t1 for t0, t1 in windows)]
        window = "timed regions"
        if len(inside) < 3:
            inside, window = self.samples, "warm-up + timed + profiling loops (timed regions < 3 samples)"
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


def pick_limiting(per):
    """The leg `value` is quoted on: the slower one."""
    return min(per, key=lambda k: per[k]["value"])


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
    legs = args.workload.split("+")
    n = args.text_mib << 20
    cores = os.cpu_count() or 1
    texts = {}
    per = {}
    for wl in legs:
        algo, alphabet, p, m, desc = WORKLOADS[wl]
        if alphabet not in texts:
            texts[alphabet] = dg.text_host(n, alphabet, TEXT_SEED)
        text = texts[alphabet]
        pats, m_max = make_patterns(dg, text, wl)
        mixed = isinstance(m, tuple)
        use_ref = oracle.ref_available() and not mixed and not (algo == "AC" and p * m_max > 400_000)

        def one_pass(sample):
            t0 = time.perf_counter()
            if use_ref:
                r = (oracle.ref_ac if algo == "AC" else oracle.ref_wu)(pats, alphabet, sample, threads=cores)
                return r["search_s"], r["count"]
            r = oracle.set_search(pats, sample, want_positions=False)  # port: single scalar thread
            return time.perf_counter() - t0, r["count"]

        # bounded sample: calibrate on 8 MiB, then size each step for a whole run of ~2 minutes over all legs
        cal_secs, _ = one_pass(text[: min(n, 8 << 20)])
        rate = min(n, 8 << 20) / max(cal_secs, 1e-6)
        budget = 120.0 / len(legs) / max(1, args.steps + args.warmup)
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
        per[algo if len(legs) > 1 else wl] = {
            "workload": wl, "value": value, "unit": "GB/s", "ms_per_step": tot / args.steps * 1e3,
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores if use_ref else 1, "kind": kind,
                             "sample": f"first {sample_n} bytes of the {n}-byte text per step (search only, "
                                       f"preprocessing excluded as in main.c:246-262; unmodified "
                                       f"search_{'ac' if algo == 'AC' else 'wu'}, {cores} threads, MPI-rank shard "
                                       f"geometry of main.c:467-477)", "count": int(cnt)},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    lim = pick_limiting(per)
    line = {
        "impl": "reference", "metric": "text GB/s scanned", "value": per[lim]["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per[lim]["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, n),
        "value_is": f"the slower leg ({lim})", "per_algo": per,
        "cpu_baseline": per[lim]["cpu_baseline"], "e2e": per[lim]["e2e"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ============================================================== our arm (GPU)
def pin_rank_to_cores(local_rank, local_world):
    """One process per GPU on one box: give every rank its own slice of the host cores, on the NUMA node of its GPU
    when the topology can be read (pinned buffers are then allocated node-local, first touch).  Returns a dict for
    the report."""
    try:
        avail = sorted(os.sched_getaffinity(0))
    except Exception:
        return {"pinned": False}
    if local_world <= 1 or len(avail) < 2 * local_world:
        return {"pinned": False, "cores": len(avail)}
    node_of = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        for g in range(local_world):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(g)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            node_of[g] = max(node, 0)
    except Exception:
        node_of = {g: 0 for g in range(local_world)}

    def cpus_of(node):
        try:
            out = []
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                out.extend(range(int(a), int(b or a) + 1))
            return [c for c in out if c in avail]
        except Exception:
            return []

    node = node_of.get(local_rank, 0)
    mates = [g for g in range(local_world) if node_of.get(g, 0) == node]
    cpus = cpus_of(node)
    if len(cpus) < 2 * len(mates):  # topology unreadable or lopsided: plain equal slices of what we may use
        mates, cpus = list(range(local_world)), avail
    k = len(cpus) // len(mates)
    i = mates.index(local_rank)
    mine = cpus[i * k:(i + 1) * k]
    try:
        os.sched_setaffinity(0, mine)
    except Exception:
        return {"pinned": False, "cores": len(avail)}
    return {"pinned": True, "numa_node": node, "cores": len(mine), "first_core": mine[0]}


def measure_pinned_copy(torch, dev, nbytes=128 << 20, window_s=0.3):
    """The box's own H2D rate for this rank's link (pinned host -> HBM, cudaMemcpyAsync): the ceiling of e2e when the
    text crosses the link one byte per symbol.  Every rank copies for the same wall-clock window (the caller lines them
    up with a barrier), so each figure is the rate with ALL links busy."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    src.zero_()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    copies = 0
    while time.perf_counter() - t0 < window_s:
        for _ in range(4):  # back to back: the DMA engine never waits for the host
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        copies += 4
    return copies * nbytes / (time.perf_counter() - t0) / 1e9


class Rig:
    """What the legs of one process share: the process group, texts by (alphabet, size), the clock sampler."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import acwm_pkg
        self.torch, self.dist = torch, dist
        self.acwm = acwm_pkg.load()
        self.dg = acwm_pkg.submodule("datagen")
        self.sh = acwm_pkg.submodule("sharding")
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.affinity = pin_rank_to_cores(self.local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world))))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.texts = {}
        self.timed_windows = []
        self.sampler = ClockSampler(self.local_rank)
        self.sampler.start()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x), [float(x)]
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        g = [self.torch.zeros(1, dtype=self.torch.float64, device=self.dev) for _ in range(self.world)]
        self.dist.all_gather(g, t)
        per = [float(v.item()) for v in g]
        return max(per), per

    def text_set(self, alphabet, n, halo, host_copies):
        """N_ROTATE distinct shards per rank, resident in HBM (+ `host_copies` of them on the host).  Buffer 0 of rank
        r is shard r of the global text = concatenation of the per-rank texts (seed TEXT_SEED + r), followed by the
        first m_max-1 bytes of shard r+1 (the halo of main.c:467-477); the others only defeat L2.  Texts much larger
        than L2 (126 MB) need no rotation: every step streams from HBM anyway, and are generated on the device."""
        if self.rank + 1 >= self.world:
            halo = 0  # the last shard (and a single GPU) has no successor to borrow a halo from
        key = (alphabet, n, halo)
        if key in self.texts:
            return self.texts[key]
        torch, dg, rank, world = self.torch, self.dg, self.rank, self.world
        if n <= (256 << 20):
            def shard(k):
                own = dg.text_host(n, alphabet, TEXT_SEED + rank + 1000 * k)
                if rank + 1 < world and halo:
                    own = np.concatenate([own, dg.text_host(n, alphabet, TEXT_SEED + rank + 1 + 1000 * k)[:halo]])
                return own
            host = [shard(k) for k in range(N_ROTATE)]
            devt = [torch.from_numpy(t).to(self.dev) for t in host]
            host = host[:max(host_copies, 1)]
        else:
            devt = [dg.text_device(n + (halo if rank + 1 < world else 0), alphabet, TEXT_SEED + rank, self.dev)]
            host = []
        self.texts = {key: (host, devt)}  # one set at a time stays resident
        return host, devt


def run_leg(rig, wl, n, steps, warmup, full=True, pats=None):
    """One workload on every rank's shard: timed device-resident loop (+ isolated launches, e2e and the CPU baseline
    when `full`).  Returns the leg's report (rank 0) -- every rank must call it (collectives inside)."""
    torch, dist, acwm, dg, sh, args = rig.torch, rig.dist, rig.acwm, rig.dg, rig.sh, rig.args
    world, rank, dev, stream = rig.world, rig.rank, rig.dev, rig.stream
    algo_name, alphabet, p, m, desc = WORKLOADS[wl]
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    m_max = m[1] if isinstance(m, tuple) else m
    halo = m_max - 1
    host_texts, dev_texts = rig.text_set(alphabet, n, halo, 2 if full else 1)
    nrot = len(dev_texts)
    # the pattern set is the same on every rank (replicated tables): "with hits" from rank 0's first text
    if pats is None:
        text0 = host_texts[0] if (host_texts and rank == 0 and world == 1) else dg.text_host(min(n, 128 << 20), alphabet, TEXT_SEED)
        pats, _ = make_patterns(dg, text0, wl)
    t_build = time.perf_counter()
    mt = acwm.Matcher(algo, pats, alphabet, **json.loads(args.matcher_opts))
    t_build = time.perf_counter() - t_build
    pos_cap = max(1 << 20, n // 16)
    mt.upload(rig.local_rank, pos_cap)
    report_from = halo if rank > 0 else 0
    count_buf = torch.zeros(1, dtype=torch.int64, device=dev)
    d_count_ptr, _ = mt.result_device_ptrs()
    count_view = _device_count_tensor(torch, d_count_ptr, dev)  # the matcher's device-resident count

    # the only thing that crosses NVLink is the per-GPU match count (8 bytes): exchanged inside the scan
    # kernel through peer-mapped mailboxes; NCCL all_reduce is the fallback
    fused_exchange = world > 1 and not args.nccl_count and sh.connect_peers(mt, dev)

    def step(i, want_positions=True):
        mt.scan_tensor(dev_texts[i % nrot], want_positions=want_positions, report_from=report_from)
        if world > 1 and not fused_exchange:
            count_buf.copy_(count_view)
            sh.allreduce_count_tensor(count_buf)

    mt.set_overlap(not args.no_overlap)  # consecutive scans of resident texts: programmatic dependent launches
    for i in range(warmup):
        step(i)
    rig.barrier()
    launches0 = mt.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        step(warmup + i)
    ev1.record()
    rig.barrier()
    rig.timed_windows.append((t_wall0, time.perf_counter()))
    launches = mt.launch_count - launches0
    ms_per_step, per_rank_ms = rig.max_over_ranks(ev0.elapsed_time(ev1) / steps)  # max over ranks
    text_bytes = sum(int(dev_texts[(warmup + i) % nrot].numel()) for i in range(steps)) / steps
    value = world * text_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- results of the last step (parity of the global count is a test, here it is reported)
    mt.set_overlap(False)
    last_count, _, _ = mt.fetch(cap=pos_cap, stream=stream)
    global_count = sh.allreduce_count(last_count, dev)
    if fused_exchange:  # what the kernels exchanged must be what NCCL sums
        fused_global = mt.fetch_global_count(stream)
        assert fused_global == global_count, (fused_global, global_count)
        mt.set_peers(0, 0, None)  # the legs below are per-rank (profiling, e2e): no exchange

    # ---- the scan kernel alone: CUDA events around single launches on the launching stream
    mt.set_profiling(True)
    scan_s, matches = [], []
    for i in range(max(min(steps, 20), 8)):
        mt.scan_tensor(dev_texts[i % nrot], want_positions=True, report_from=report_from)
        a, _ = mt.profiled_seconds()
        c, _, _ = mt.fetch(cap=0, stream=stream)
        scan_s.append(a)
        matches.append(c)
    mt.set_profiling(False)
    scan_isolated = float(np.mean(scan_s[2:]))
    alg_bytes = float(np.mean([dev_texts[i % nrot].numel() + 8 * matches[i] for i in range(len(matches))]))
    peak, peak_src = peaks()
    # The kernel's average launch duration over the timed region.  A step is ONE launch of the scan kernel and nothing
    # else whenever the count is exchanged inside the kernel (every N) -- then the timed region / K is that average;
    # with a collective per step (--nccl-count, or no peer memory) the event-bracketed single launches are used.
    one_launch_per_step = launches == steps
    scan_mean = ms_per_step * 1e-3 if one_launch_per_step else scan_isolated
    achieved = alg_bytes / scan_mean / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(wl, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    info = mt.info
    rep = {
        "workload": wl, "algo": algo_name, "value": value, "unit": "GB/s", "ms_per_step": ms_per_step,
        "text_bytes_per_gpu": n, "gpu_launches": int(launches), "matches_last_step": int(global_count),
        "per_rank_ms_per_step": [round(x, 5) for x in per_rank_ms],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "scan_kernel (one launch: TMA-fed scan + barrier-free position ordering + count exchange)",
                     "kernel_ms": scan_mean * 1e3, "kernel_ms_isolated_launch": scan_isolated * 1e3,
                     "frac_isolated_launch": alg_bytes / scan_isolated / 1e9 / peak,
                     "duration_source": "timed region / K (one launch per step)" if one_launch_per_step
                     else "CUDA events around single launches (a collective per step)",
                     "algorithmic_bytes_per_launch": alg_bytes, "frac_of_8TBps_spec": achieved / 8000.0},
        "count_exchange": ("none (1 GPU)" if world == 1 else
                           "in-kernel system-scope stores into NVLink peer mailboxes (torch symmetric memory)"
                           if fused_exchange else "NCCL all_reduce of the 8-byte count per step"),
        "kernel": {k: info[k] for k in ("packed2bit", "stride", "depth", "exact_front", "front_kind", "n_rows",
                                         "table_in_smem", "smem_bytes", "threads", "stages", "ctas_per_sm")},
        "table_build_s": round(t_build, 3),
    }
    if not full:
        mt.close()
        return rep

    # ---- end to end through the public API: pinned host text -> count + positions on the host
    if not host_texts:  # a text generated on the device (larger than 256 MiB): its host copy is the e2e input
        host_texts = [dev_texts[0].cpu().numpy()]
    pinned = [torch.from_numpy(t).pin_memory() for t in host_texts[:2]]
    if len(pinned) == 1:
        pinned = pinned * 2
    e2e_steps = max(3, min(steps, 10))
    pos_out = np.empty(pos_cap, np.uint64)  # the caller's position buffer, reused
    for i in range(9):  # acwm_search_host sets up and measures its two transfers and the neighbours of its raw share first
        mt.search_host(pinned[i % 2], out=pos_out)
    rig.barrier()
    t0 = time.perf_counter()
    counts = []
    kernel_s = 0.0
    for i in range(e2e_steps):
        c, _ = mt.search_host(pinned[i % 2], out=pos_out)
        counts.append(c)
        kernel_s += mt.last_kernel_seconds
    own_s = time.perf_counter() - t0
    if world > 1:  # the counts of all steps cross NVLink in ONE all-reduce (8 B per step), not one blocking call per step
        ct = torch.tensor(counts, dtype=torch.int64, device=dev)
        dist.all_reduce(ct)
        counts = [int(x) for x in ct.tolist()]
    rig.barrier()
    e2e_s, _ = rig.max_over_ranks(time.perf_counter() - t0)
    e2e_bytes = int(pinned[0].numel())
    e2e_value = world * e2e_bytes * e2e_steps / e2e_s / 1e9
    h2d = int(mt.last_h2d_bytes)
    rep["e2e"] = {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "text_bytes_per_step": e2e_bytes,
                  "h2d": ("hybrid: a prefix of the text copied one byte per symbol by DMA while the host cores pack the "
                          "rest to 2 bits per symbol (csrc/hostpack.cpp)" if h2d < e2e_bytes
                          else "text copied one byte per symbol"),
                  "d2h_bytes_per_step": 8 * int(last_count) + 32, "steps": e2e_steps,
                  "per_gpu_text_GBps_this_rank": e2e_bytes * e2e_steps / own_s / 1e9,
                  "per_gpu_link_GBps_this_rank": h2d * e2e_steps / own_s / 1e9,
                  "kernel_ms_per_step": kernel_s / e2e_steps * 1e3,
                  "api": "acwm_search_host (pinned host text -> host count + positions)"}

    # ---- CPU baseline: the reference on the host cores (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        mt.scan_tensor(dev_texts[0], want_positions=False)
        gpu_count0, _, _ = mt.fetch(cap=0, stream=stream)
        rep["cpu_baseline"] = cpu_baseline(host_texts[0], pats, algo_name, alphabet, m, gpu_count0)
    mt.close()
    return rep


def run_ours(args):
    rig = Rig(args)
    torch, world, rank = rig.torch, rig.world, rig.rank
    n = args.text_mib << 20  # per-GPU shard (weak scaling)
    legs = args.workload.split("+")
    per = {}
    for wl in legs:
        rep = run_leg(rig, wl, n, args.steps, args.warmup, full=True)
        per[rep["algo"] if len(legs) > 1 else wl] = rep
    rig.barrier()  # every rank copies at the same time: what the box's memory and PCIe root deliver to all links together
    link = measure_pinned_copy(torch, rig.dev)
    _, links = rig.max_over_ranks(link)

    # ---- the sizes the north star scales on: 1 GiB of DNA per GPU for both algorithms, configs[2] at 10^9 B per GPU,
    # configs[3] (bytes, 10 000 mixed-length patterns) at its full 2 * 10^9 B per GPU
    big = []
    if not args.no_big_legs and args.workload == DEFAULT_WORKLOAD:
        for wl, nb in (("c1", 1 << 30), ("c2", 1 << 30), ("c3", 10 ** 9), ("c4", 2 * 10 ** 9)):
            rep = run_leg(rig, wl, nb, max(5, args.steps // 5), 3, full=False)
            big.append({k: rep[k] for k in ("workload", "algo", "text_bytes_per_gpu", "value", "unit", "ms_per_step",
                                            "gpu_launches", "matches_last_step", "roofline", "kernel", "count_exchange",
                                            "table_build_s")})
    rig.sampler.stop_flag = True
    rig.sampler.join(timeout=1)
    clocks = rig.sampler.summary(rig.timed_windows[:len(legs)])

    if rank == 0:
        lim = pick_limiting(per)
        e2e_lim = min(per, key=lambda k: per[k]["e2e"]["value"])
        e2e = dict(per[e2e_lim]["e2e"])
        e2e["leg"] = e2e_lim
        e2e["pinned_copy_GBps_per_rank"] = [round(x, 2) for x in links]
        e2e["pinned_copy_GBps_all_ranks"] = round(float(sum(links)), 1)
        e2e["pinned_copy_note"] = ("plain cudaMemcpyAsync from pinned memory, all ranks at the same time: the box's ceiling for "
                                   "text that crosses the links one byte per symbol (acwm_search_host measures its hybrid and "
                                   "its plain transfer on the first two calls and keeps the faster)")
        e2e["affinity"] = rig.affinity
        line = {
            "metric": "text GB/s scanned", "value": per[lim]["value"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per[lim]["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.workload, n),
            "value_is": f"the slower leg ({lim}); every leg runs {args.steps} timed steps after {args.warmup} warm-up steps",
            "per_algo": per,
            "e2e": e2e,
            "gpu_launches": int(sum(per[k]["gpu_launches"] for k in per)),
            "clocks": clocks,
            "roofline": per[lim]["roofline"],
            "details": {
                "l2": f"{N_ROTATE} distinct {args.text_mib} MiB texts cycled per leg (working set > L2)"
                      if n <= (256 << 20) else f"one {args.text_mib} MiB text per GPU (larger than L2)",
                "positions": "count + sorted uint64 positions produced every step",
                "launch": "one kernel per step (scan + position ordering + result + count exchange), no memset / finalize "
                          "nodes; " + ("cooperative launches" if args.no_overlap else
                                       "consecutive steps chained as programmatic dependent launches (acwm_set_overlap)"),
                "halo_bytes": {k: (WORKLOADS[per[k]["workload"]][3][1] if isinstance(WORKLOADS[per[k]["workload"]][3], tuple)
                                   else WORKLOADS[per[k]["workload"]][3]) - 1 for k in per},
            },
            "matches_last_step": per[lim]["matches_last_step"],
            "per_rank_ms_per_step": per[lim]["per_rank_ms_per_step"],
        }
        if "cpu_baseline" in per[lim]:
            line["cpu_baseline"] = per[lim]["cpu_baseline"]
        if big:
            line["north_star_legs"] = big
        print(json.dumps(line), flush=True)
    if world > 1:
        rig.dist.destroy_process_group()


def _device_count_tensor(torch, ptr, dev):
    """int64[1] view of the matcher's device-resident count (no host round trip)."""
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(w, device=dev)


def cpu_baseline(text, pats, algo_name, alphabet, m, gpu_count):
    import oracle
    cores = os.cpu_count() or 1
    mixed = isinstance(m, tuple)
    n = text.size
    p = len(pats)
    m_max = m[1] if mixed else m
    use_ref = oracle.ref_available() and not mixed and not (algo_name == "AC" and p * m_max > 400_000)
    passes, tot, cnt = 0, 0.0, None
    t_start = time.perf_counter()
    while passes < 5 and time.perf_counter() - t_start < 12.0:
        if use_ref:
            r = (oracle.ref_ac if algo_name == "AC" else oracle.ref_wu)(pats, alphabet, text, threads=cores)
            tot += r["search_s"]
            cnt = r["count"]
        else:
            t0 = time.perf_counter()
            cnt = oracle.set_search(pats, text, want_positions=False)["count"]
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
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD,
                    help="c1+c2 (default: both algorithms, value = the slower) or any of " + ", ".join(sorted(WORKLOADS))
                         + ", or several joined with +")
    ap.add_argument("--text-mib", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-big-legs", action="store_true", help="skip the 1 GiB / configs[2] legs of the default run")
    ap.add_argument("--no-overlap", action="store_true", help="plain cooperative launches in the timed loop")
    ap.add_argument("--matcher-opts", default="{}", help="JSON of acwm_options overrides (tuning experiments)")
    ap.add_argument("--nccl-count", action="store_true", help="all-reduce the count with NCCL instead of in-kernel")
    args = ap.parse_args()
    for wl in args.workload.split("+"):
        if wl not in WORKLOADS:
            ap.error(f"unknown workload {wl}")
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
