"""torchrun entry: every rank scans its shard of one seeded text on its own GPU; the summed count and the
gathered positions must equal the oracle's on the whole text (rank 0 checks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import acwm_pkg

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
sh = acwm_pkg.submodule("sharding")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
ok = True
for algo, alphabet, p, m, n in ((acwm.AC, 4, 100, 8, 64 << 20), (acwm.WM, 4, 1000, 16, 64 << 20),
                                (acwm.WM, 256, 2000, (8, 64), 32 << 20)):
    text = dg.text_host(n, alphabet, 5)
    if isinstance(m, tuple):
        pats = dg.mixed_patterns_with_hits(text, p, m[0], m[1], alphabet, 6)
        m_max = m[1]
    else:
        pats = dg.patterns_with_hits(text, p, m, alphabet, 6)
        m_max = m
    mt = acwm.Matcher(algo, pats, alphabet)
    total, pos = sh.search_sharded_host(mt, text, m_max)
    if rank == 0:
        import oracle
        ref = oracle.set_search(pats, text)
        good = total == ref["count"] and np.array_equal(pos, ref["positions"])
        ok &= good
        print(f"sharded x{world} algo={algo} alphabet={alphabet} p={p} n={n}: count {total} vs oracle {ref['count']}, "
              f"positions {'identical' if good else 'DIFFER'}", flush=True)
    # the same shards again, device-resident, with the count exchanged inside the scan kernel
    start, length, report_from = sh.shard_of(text.size, world, rank, m_max)
    fused = sh.connect_peers(mt)
    shard = torch.from_numpy(text[start:start + length]).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):  # several epochs: the mailbox slots alternate
        mt.scan_tensor(shard, want_positions=bool(rep & 1), report_from=report_from)
        g = mt.fetch_global_count(st) if fused else None
        if rank == 0:
            good = (g == ref["count"]) if fused else True
            ok &= good
            print(f"   fused exchange rep {rep}: global count {g} ({'ok' if good else 'WRONG'}; fused={fused})", flush=True)
    mt.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
