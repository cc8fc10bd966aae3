"""Host-side logic of the multi-GPU layer on CPU: shard geometry (== main.c:467-477),
exactly-once ownership of match ends, and the world_size-2 count all-reduce + position
gather over gloo.  The per-shard scanner here is the oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest

from cases import RANDOM_CASES, make_case


def test_shard_bounds_match_reference_geometry(acwm):
    for n in (1, 7, 100, 12345, 1 << 20, (1 << 33) + 5):
        for world in (1, 2, 3, 4, 8):
            for halo in (0, 7, 31, 63):
                chunk = -(-n // world)
                covered = 0
                for r in range(world):
                    s, l = acwm.shard_bounds(n, world, r, halo)
                    start = min(r * chunk, n)
                    stop = min((r + 1) * chunk + halo, n)
                    assert (s, l) == (start, max(0, stop - start))
                    covered = max(covered, s + l)
                assert covered == n


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_scan_is_exactly_once(acwm, oracle, world):
    sh = __import__("acwm_pkg").submodule("sharding")
    for cname in ("c1_ac_dna_p100_m8", "wm_dna_mixed_8_64"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        pats, text = make_case(case)
        text = text[:90_001]
        m_max = max(q.size for q in pats) if isinstance(pats, list) else pats.shape[1]
        whole = oracle.set_search(pats, text)
        got = []
        for r in range(world):
            start, length, report_from = sh.shard_of(text.size, world, r, m_max)
            res = oracle.set_search(pats, text[start:start + length])
            pos = res["positions"]
            got.append(pos[pos >= report_from] + np.uint64(start))
        got = np.concatenate(got)
        assert got.size == whole["count"]
        assert np.array_equal(got, whole["positions"])  # already globally sorted


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import acwm_pkg
    import oracle
    sh = acwm_pkg.submodule("sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = next(c for c in RANDOM_CASES if c[0] == "c2_wm_dna_p1000_m16")
    pats, text = make_case(case)
    text = text[:150_000]
    start, length, report_from = sh.shard_of(text.size, world, rank, 16)
    res = oracle.set_search(pats, text[start:start + length])
    pos = res["positions"][res["positions"] >= report_from]
    total = sh.allreduce_count(int(pos.size))
    glob = sh.gather_positions(pos, start)
    if rank == 0:
        whole = oracle.set_search(pats, text)
        q.put((total == whole["count"], bool(np.array_equal(glob, whole["positions"]))))
    dist.destroy_process_group()


def test_gloo_world2_count_allreduce_and_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == (True, True)
