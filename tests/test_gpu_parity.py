"""GPU parity tests proper: every call goes through the C ABI of libacwm_b200.so
(acwm_search_host / acwm_scan_device + acwm_fetch / the reference-shaped shims) and is
compared bit-exactly -- match count and every match position -- with the oracle on the
same seeded inputs, with the committed golden vectors (reference counts), and at
BASELINE sizes through size-independent properties as well."""
import numpy as np
import pytest

from cases import RANDOM_CASES, edge_cases, make_case
from golden_util import load_golden

pytestmark = pytest.mark.gpu
GOLD = load_golden()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def _check(acwm, oracle, algo, pats, alphabet, text, **opts):
    mt = acwm.Matcher(algo, pats, alphabet, **opts)
    count, pos = mt.search_host(text, cap=max(1, text.size * 2))
    ref = oracle.set_search(pats, text)
    assert count == ref["count"], (count, ref["count"], mt.info)
    assert np.array_equal(pos, ref["positions"]), mt.info
    return mt


@pytest.mark.parametrize("case", RANDOM_CASES, ids=lambda c: c[0])
def test_random_cases_match_oracle(acwm, oracle, torch_cuda, case):
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    _check(acwm, oracle, algo, pats, alphabet, text, **opts).close()


@pytest.mark.parametrize("case", [c for c in RANDOM_CASES if c[6].get("force_front") == 2], ids=lambda c: c[0])
def test_filtered_ac_with_the_automaton_as_verifier(acwm, oracle, torch_cuda, case, monkeypatch):
    """ACWM_VERIFY_DFA=1: candidate windows of the filtered AC are decided by a walk of the full-depth automaton
    (verify_dfa) instead of the bucket compare -- same matches."""
    monkeypatch.setenv("ACWM_VERIFY_DFA", "1")
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    mt = _check(acwm, oracle, algo, pats, alphabet, text, **opts)
    assert mt.params().verify_kind == 1
    mt.close()


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("algo_name", ["AC", "WM"])
def test_golden_vectors(acwm, torch_cuda, name, algo_name):
    g = GOLD[name]
    pats, text, alphabet = g["patterns"], g["text"], int(g["alphabet"])
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    mt = acwm.Matcher(algo, pats, alphabet)
    count, pos = mt.search_host(text, cap=max(1, text.size))
    assert count == int(g["ref_ac_count"]) == int(g["ref_wu_count"])
    assert np.array_equal(pos, g["positions"])
    mt.close()


def test_edge_cases(acwm, torch_cuda):
    for name, alphabet, pats, text, exp in edge_cases():
        for algo in (acwm.AC, acwm.WM):
            for opts in ({}, dict(force_bytes_path=1)):
                mt = acwm.Matcher(algo, pats, alphabet, **opts)
                count, pos = mt.search_host(text, cap=64)
                assert pos.tolist() == exp and count == len(exp), (name, algo, opts)
                mt.close()


def test_empty_text(acwm, torch_cuda):
    mt = acwm.Matcher(acwm.AC, np.zeros((1, 4), np.uint8), 4)
    count, pos = mt.search_host(np.zeros(0, np.uint8), cap=4)
    assert count == 0 and pos.size == 0


@pytest.mark.parametrize("threads,stages", [(128, 1), (256, 1), (384, 1), (512, 1), (768, 1), (1024, 1),
                                            (128, 2), (256, 2), (384, 2), (512, 2)])
def test_launch_shape_variants(acwm, oracle, torch_cuda, threads, stages):
    """Every (warps, ring depth) shape of the scan kernel gives the same matches."""
    for cname in ("c1_ac_dna_p100_m8", "c2_wm_dna_p1000_m16", "ac_dna_depth5", "wm_ascii_p1000_m8",
                  "ac_protein_p100_m6"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        if (alphabet > 4) != (stages == 2) or (alphabet > 4 and threads > 512):
            continue  # 2-bit path: one slot per warp; bytes path (walks the raw tile): two slots, <= 16 warps
        try:
            acwm.Matcher(algo, pats, alphabet, force_threads=threads, force_stages=stages, **opts).close()
        except acwm.AcwmError as e:
            assert e.code == acwm.ERR_INVALID  # this shape does not fit next to the case's tables
            continue
        _check(acwm, oracle, algo, pats, alphabet, text, force_threads=threads, force_stages=stages, **opts).close()


def test_device_resident_unaligned_and_report_from(acwm, oracle, torch_cuda):
    torch = torch_cuda
    for cname in ("c1_ac_dna_p100_m8", "c2_wm_dna_p1000_m16", "wm_ascii_p1000_m8", "ac_protein_p100_m6"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        mt = acwm.Matcher(algo, pats, alphabet, **opts).upload(pos_capacity=text.size)
        d_all = torch.from_numpy(text).cuda()
        st = torch.cuda.current_stream().cuda_stream
        for off, ln in ((0, text.size), (1, 70_001), (5, 3584 * 3), (13, 50_000), (16, 3584 * 5 + 3), (31, 777)):
            sub = d_all[off:off + ln]
            mt.scan_tensor(sub)
            count, pos, _ = mt.fetch(cap=text.size, stream=st)
            ref = oracle.set_search(pats, text[off:off + ln])
            assert count == ref["count"], (cname, off, ln)
            assert np.array_equal(pos, ref["positions"]), (cname, off, ln)
        # report_from drops exactly the ends below it
        mt.scan_tensor(d_all, report_from=12345)
        count, pos, _ = mt.fetch(cap=text.size, stream=st)
        ref = oracle.set_search(pats, text)
        keep = ref["positions"][ref["positions"] >= 12345]
        assert count == keep.size and np.array_equal(pos, keep)
        # count-only scan
        mt.scan_tensor(d_all, want_positions=False)
        count, pos, _ = mt.fetch(cap=0, stream=st)
        assert count == ref["count"]
        mt.close()


def test_overlapped_back_to_back_scans(acwm, oracle, torch_cuda):
    """acwm_set_overlap: consecutive scans chained as programmatic dependent launches (the next scan's
    read-only phase runs while the previous one orders its matches) still give the oracle's result."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    for cname in ("c1_ac_dna_p100_m8", "c2_wm_dna_p1000_m16", "wm_ascii_p1000_m8", "ac_dna_depth5"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        big = np.concatenate([text] * 100)  # 15-30 MB: enough tiles for every warp of every CTA (the launches chain)
        other = dg.text_host(big.size - 12345, alphabet, 77)
        other[1000:1000 + pats.shape[1]] = pats[0]
        refs = [oracle.set_search(pats, t) for t in (big, other)]
        d = [torch.from_numpy(t).cuda() for t in (big, other)]
        mt = acwm.Matcher(algo, pats, alphabet, **opts).upload(pos_capacity=big.size)
        mt.set_overlap(True)
        st = torch.cuda.current_stream().cuda_stream
        for rounds in (1, 2, 5):
            for last in (0, 1):
                for k in range(rounds):  # no synchronisation between the launches
                    mt.scan_tensor(d[(last + k + 1) % 2], want_positions=bool(k & 1))
                mt.scan_tensor(d[last])
                count, pos, _ = mt.fetch(cap=big.size, stream=st)
                assert count == refs[last]["count"], (cname, rounds, last)
                assert np.array_equal(pos, refs[last]["positions"]), (cname, rounds, last)
        mt.close()


@pytest.mark.parametrize("threads", [0, 128])
@pytest.mark.parametrize("algo_name", ["AC", "WM"])
def test_dense_matches_back_to_back(acwm, oracle, torch_cuda, threads, algo_name):
    """Match-dense text (1 position in 20 matches; then every position): the warps' staging reservations double
    and straddle blocks, every CTA places its own matches behind a look-back over the spans in front of it, and
    consecutive launches overlap -- count and positions still equal the oracle's."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    rng = np.random.default_rng(5)
    n = 24 << 20
    texts = [dg.text_host(n, 4, 31), dg.text_host(n - 777, 4, 32)]
    all4 = np.array([[(v >> (2 * k)) & 3 for k in range(4)] for v in range(256)], np.uint8)
    for pats in (all4[rng.choice(256, 13, replace=False)], all4):
        refs = [oracle.set_search(pats, t) for t in texts]
        d = [torch.from_numpy(t).cuda() for t in texts]
        opts = dict(force_threads=threads) if threads else {}
        mt = acwm.Matcher(acwm.AC if algo_name == "AC" else acwm.WM, pats, 4, **opts).upload(pos_capacity=n)
        mt.set_overlap(True)
        st = torch.cuda.current_stream().cuda_stream
        for last in (0, 1, 0):
            for k in range(3):  # no synchronisation between the launches
                mt.scan_tensor(d[(last + k + 1) % 2])
            mt.scan_tensor(d[last])
            count, pos, _ = mt.fetch(cap=n, stream=st)
            assert count == refs[last]["count"], (algo_name, threads, len(pats), last)
            assert np.array_equal(pos, refs[last]["positions"]), (algo_name, threads, len(pats), last)
        mt.close()


def test_host_text_packed_on_the_host(acwm, oracle, torch_cuda):
    """Host texts of 2-bit matchers (>= 4 Mi symbols) are packed 4 symbols per byte by the host cores before the H2D
    copy and scanned packed: same count and positions as the oracle, for sizes that end inside a byte / a 16-byte
    piece / a tile, for fronts with and without verification (incl. the packed-text compare for m > 16)."""
    import os
    dg = __import__("acwm_pkg").submodule("datagen")
    base = dg.text_host((17 << 20) + 1000, 4, 41)
    saved = {k: os.environ.get(k) for k in ("ACWM_HOST_PACK", "ACWM_HOST_RAW_PERCENT")}
    os.environ["ACWM_HOST_PACK"] = "2"  # also on a box with few cores
    os.environ["ACWM_HOST_RAW_PERCENT"] = "30"  # the default follows the core count (none on a many-core box)
    try:
        _host_packed_cases(acwm, oracle, dg, base)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_hybrid_transfer_raw_share_variants(acwm, oracle, torch_cuda):
    """The raw share of the hybrid host transfer follows the host's core count (csrc/api.cu host_raw_percent): every
    split of a pinned text into raw chunks + packed rest gives the oracle's count and positions."""
    import os
    import torch
    dg = __import__("acwm_pkg").submodule("datagen")
    base = dg.text_host((15 << 20) + 333, 4, 47)
    pinned = torch.from_numpy(np.concatenate([base] * 6)[: (88 << 20) + 1001]).pin_memory()  # 7 chunks of 14 Mi
    saved = {k: os.environ.get(k) for k in ("ACWM_HOST_PACK", "ACWM_HOST_RAW_PERCENT")}
    os.environ["ACWM_HOST_PACK"] = "2"
    try:
        for algo, p, m in ((acwm.WM, 1000, 16), (acwm.AC, 100, 8), (acwm.WM, 300, (8, 64))):
            pats = (dg.mixed_patterns_with_hits(base, p, m[0], m[1], 4, 11) if isinstance(m, tuple)
                    else dg.patterns_with_hits(base, p, m, 4, 11))
            mt = acwm.Matcher(algo, pats, 4)
            ref = oracle.set_search(pats, pinned.numpy())
            for pct in (5, 10, 35, 60, 90):  # 0 / 1 / 2 / 4 / 6 raw chunks
                os.environ["ACWM_HOST_RAW_PERCENT"] = str(pct)
                count, pos = mt.search_host(pinned, cap=max(1, ref["count"]))
                assert count == ref["count"] and np.array_equal(pos, ref["positions"]), (algo, p, m, pct)
                raw_chunks = min(6, (7 * pct + 50) // 100)
                assert mt.last_h2d_bytes > pinned.numel() // 4 + raw_chunks * (14 << 20) * 3 // 4 - 4096, (pct, mt.last_h2d_bytes)
            mt.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _host_packed_cases(acwm, oracle, dg, base):
    cases = [(acwm.AC, 100, 8, {}), (acwm.WM, 1000, 16, {}), (acwm.AC, 1000, 16, {}), (acwm.WM, 20000, 32, {}),
             (acwm.AC, 3, 70, {}), (acwm.WM, 300, (8, 64), {})]
    for algo, p, m, opts in cases:
        if isinstance(m, tuple):
            pats = dg.mixed_patterns_with_hits(base, p, m[0], m[1], 4, 7)
        else:
            pats = dg.patterns_with_hits(base, p, m, 4, 7)
        mt = acwm.Matcher(algo, pats, 4, **opts)
        sizes = ((4 << 20), (4 << 20) + 1, (5 << 20) + 63, (16 << 20) + 3584 * 3 + 17, base.size)
        if p == 1000 and algo == acwm.WM:  # once: more chunks than the pinned ring has slots
            big = np.concatenate([base] * 15)[: (250 << 20) + 5]
            ref = oracle.set_search(pats, big)
            count, pos = mt.search_host(big, cap=max(1, ref["count"]))
            assert count == ref["count"] and np.array_equal(pos, ref["positions"])
        # from PINNED memory a prefix of the text travels unpacked beside the packed rest (>= 5 chunks of 14 Mi symbols)
        import torch
        pinned = torch.from_numpy(np.concatenate([base] * 5)[: (75 << 20) + 1001]).pin_memory()
        ref = oracle.set_search(pats, pinned.numpy())
        for rep in range(2):
            count, pos = mt.search_host(pinned, cap=max(1, ref["count"]))
            assert count == ref["count"], (algo, p, m, "pinned", rep)
            assert np.array_equal(pos, ref["positions"]), (algo, p, m, "pinned", rep)
        for n in sizes:
            text = base[:n]
            ref = oracle.set_search(pats, text)
            count, pos = mt.search_host(text, cap=max(1, ref["count"]))
            assert count == ref["count"], (algo, p, m, n)
            assert np.array_equal(pos, ref["positions"]), (algo, p, m, n)
        bad = base[:5 << 20].copy()
        bad[(3 << 20) + 5] = 7
        with pytest.raises(acwm.AcwmError) as ei:
            mt.search_host(bad, cap=10)
        assert ei.value.code == acwm.ERR_BAD_TEXT
        mt.close()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_scans_sum_to_whole(acwm, oracle, torch_cuda, world):
    """The multi-GPU geometry run on one GPU: shard scans are exactly-once."""
    torch = torch_cuda
    sh = __import__("acwm_pkg").submodule("sharding")
    for cname in ("c2_wm_dna_p1000_m16", "c4_wm_ascii_mixed_8_64", "ac_dna_p1000_m16_trunc"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        m_max = max(q.size for q in pats) if isinstance(pats, list) else pats.shape[1]
        mt = acwm.Matcher(algo, pats, alphabet, **opts).upload(pos_capacity=text.size)
        d_all = torch.from_numpy(text).cuda()
        st = torch.cuda.current_stream().cuda_stream
        got = []
        for r in range(world):
            start, length, report_from = sh.shard_of(text.size, world, r, m_max)
            mt.scan_tensor(d_all[start:start + length], report_from=report_from)
            count, pos, _ = mt.fetch(cap=text.size, stream=st)
            assert count == pos.size
            got.append(pos + np.uint64(start))
        got = np.concatenate(got)
        ref = oracle.set_search(pats, text)
        assert got.size == ref["count"] and np.array_equal(got, ref["positions"])
        mt.close()


@pytest.mark.parametrize("world", [1, 3, 8])
def test_search_host_sharded_one_process(acwm, oracle, torch_cuda, world):
    """acwm_search_host_sharded = Scatterv + per-rank search + Reduce of main.c:464-656 in one call: one matcher and
    one host thread per shard, spread over the devices present (all on GPU 0 on a one-GPU box)."""
    n_dev = acwm.device_count()
    assert n_dev >= 1
    for cname in ("c2_wm_dna_p1000_m16", "c4_wm_ascii_mixed_8_64", "c1_ac_dna_p100_m8"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        ref = oracle.set_search(pats, text)
        mts = [acwm.Matcher(algo, pats, alphabet, **opts) for _ in range(world)]
        for r, mt in enumerate(mts):
            mt.upload(device=r % n_dev)
        for rep in range(2):  # the second call reuses every buffer of the first
            count, pos, per = acwm.search_host_sharded(mts, text, cap=max(1, ref["count"]))
            assert count == ref["count"] and int(per.sum()) == count, (cname, world, count, ref["count"])
            assert np.array_equal(pos, ref["positions"]), (cname, world)
        # per-shard counts are those of the reference's ranks: ends e with start + m_max-1 <= e (first shard: all)
        m_max = max(q.size for q in pats) if isinstance(pats, list) else pats.shape[1]
        for r in range(world):
            start, length = acwm.shard_bounds(text.size, world, r, m_max - 1)
            lo = start + (m_max - 1 if r else 0)
            inside = (ref["positions"] >= lo) & (ref["positions"] < start + length)
            assert int(per[r]) == int(inside.sum()), (cname, world, r)
        # count only, and too small a position buffer: the count stays exact
        count, pos, _ = acwm.search_host_sharded(mts, text, want_positions=False)
        assert count == ref["count"] and pos.size == 0
        if ref["count"] > 2:
            count, pos, _ = acwm.search_host_sharded(mts, text, cap=2, allow_overflow=True)
            assert count == ref["count"] and np.array_equal(pos, ref["positions"][:2])
        for mt in mts:
            mt.close()
    if world == 3:  # > 1 Mi positions: the gather runs one thread per shard
        text = np.zeros(1_500_000, np.uint8)
        text[700_000] = 1
        pats = np.zeros((1, 8), np.uint8)
        mts = [acwm.Matcher(acwm.AC, pats, 4).upload(device=r % n_dev) for r in range(world)]
        ref = oracle.set_search(pats, text)
        count, pos, per = acwm.search_host_sharded(mts, text, cap=text.size)
        assert count == ref["count"] > (1 << 20) and np.array_equal(pos, ref["positions"])
        for mt in mts:
            mt.close()
    # one matcher handed in twice / different pattern sets are refused
    a = acwm.Matcher(acwm.WM, np.zeros((1, 8), np.uint8), 4)
    b = acwm.Matcher(acwm.WM, np.ones((1, 8), np.uint8), 4)
    for bad in ([a, a], [a, b]):
        with pytest.raises(acwm.AcwmError):
            acwm.search_host_sharded(bad, np.zeros(1000, np.uint8), cap=10)
    a.close()
    b.close()


def test_search_host_sharded_packed_shards(acwm, oracle, torch_cuda):
    """Shards large enough for the host packer (and, from pinned memory, for the hybrid raw + packed transfer): the
    ends a shard leaves to its predecessor (mixed-length patterns: report_from = m_max - 1) are dropped on these
    paths too."""
    import os
    import torch
    dg = __import__("acwm_pkg").submodule("datagen")
    base = dg.text_host((17 << 20) + 1000, 4, 43)
    n_dev = acwm.device_count()
    saved = {k: os.environ.get(k) for k in ("ACWM_HOST_PACK", "ACWM_HOST_RAW_PERCENT")}
    os.environ["ACWM_HOST_PACK"] = "2"
    os.environ["ACWM_HOST_RAW_PERCENT"] = "30"
    try:
        for algo, p, m in ((acwm.WM, 300, (8, 64)), (acwm.AC, 100, 8)):
            pats = (dg.mixed_patterns_with_hits(base, p, m[0], m[1], 4, 9) if isinstance(m, tuple)
                    else dg.patterns_with_hits(base, p, m, 4, 9))
            mts = [acwm.Matcher(algo, pats, 4).upload(device=r % n_dev) for r in range(2)]
            ref = oracle.set_search(pats, base)
            count, pos, per = acwm.search_host_sharded(mts, base, cap=max(1, ref["count"]))
            assert count == ref["count"] and np.array_equal(pos, ref["positions"]), (algo, p, m)
            pinned = torch.from_numpy(np.concatenate([base] * 9)[: (150 << 20) + 77]).pin_memory()
            ref = oracle.set_search(pats, pinned.numpy())
            count, pos, per = acwm.search_host_sharded(mts, pinned, cap=max(1, ref["count"]))
            assert count == ref["count"] and np.array_equal(pos, ref["positions"]), (algo, p, m, "pinned")
            for mt in mts:
                mt.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_overflow_and_bad_text(acwm, torch_cuda):
    text = np.zeros(10_000, np.uint8)
    mt = acwm.Matcher(acwm.AC, np.zeros((1, 4), np.uint8), 4)
    count, pos = mt.search_host(text, cap=100, allow_overflow=True)
    assert count == 10_000 - 3 and mt.last_rc == acwm.ERR_OVERFLOW
    bad = text.copy()
    bad[5000] = 9
    with pytest.raises(acwm.AcwmError) as e:
        mt.search_host(bad, cap=20_000)
    assert e.value.code == acwm.ERR_BAD_TEXT
    # after an error the matcher still works
    count, pos = mt.search_host(text, cap=20_000)
    assert count == 10_000 - 3 and np.array_equal(pos, np.arange(3, 10_000, dtype=np.uint64))
    # the bytes path has no such restriction: out-of-alphabet bytes simply never match
    mt2 = acwm.Matcher(acwm.AC, np.zeros((1, 4), np.uint8), 4, force_bytes_path=1)
    count, pos = mt2.search_host(bad, cap=20_000)
    assert count == 10_000 - 3 - 4


def test_reference_shaped_shims(acwm, oracle, torch_cuda, capfd):
    """Reads like the reference's own driver (main.c:125-157, 268-298, 582-648)."""
    sm = __import__("acwm_pkg").submodule("smatcher")
    case = RANDOM_CASES[0]
    name, algo, alphabet, p, m, n, opts = case
    pattern, text = make_case(case)
    n = text.size
    want = oracle.set_search(pattern, text)["count"]
    # multiac
    state_transition, state_supply, state_final = sm.alloc_ac_tables(m, p, alphabet)
    table = sm.preproc_ac(pattern, m, p, alphabet, state_transition, state_supply, state_final)
    assert sm.search_ac(text, n, table) == want
    sm.free_ac(table, alphabet)
    # cuda_ac1..5 work from the flat tables alone
    for k in (1, 5):
        assert sm.cuda_ac(k, m, text, n, p, alphabet, state_transition, state_supply, state_final) == want
    out = capfd.readouterr().out
    assert f"Kernel 5 matches \t{want}\t time" in out
    # the caller refills the SAME arrays with another set of the same terminal-state layout (one pattern: state m is
    # final either way): the goto table is part of the matcher's cache key, so the new set is what gets searched
    one_a, one_b = pattern[:1].copy(), pattern[5:6].copy()
    text1 = text.copy()
    text1[100:100 + m], text1[500:500 + m], text1[900:900 + m] = one_a[0], one_b[0], one_b[0]
    tr_a, sup_a, fin_a = sm.alloc_ac_tables(m, 1, alphabet)
    sm.free_ac(sm.preproc_ac(one_a, m, 1, alphabet, tr_a, sup_a, fin_a), alphabet)
    assert sm.cuda_ac(5, m, text1, n, 1, alphabet, tr_a, sup_a, fin_a) == oracle.set_search(one_a, text1)["count"]
    tr_b, sup_b, fin_b = sm.alloc_ac_tables(m, 1, alphabet)
    sm.free_ac(sm.preproc_ac(one_b, m, 1, alphabet, tr_b, sup_b, fin_b), alphabet)
    assert np.array_equal(fin_a, fin_b) and not np.array_equal(tr_a, tr_b)
    np.copyto(tr_a, tr_b), np.copyto(sup_a, sup_b)
    assert sm.cuda_ac(5, m, text1, n, 1, alphabet, tr_a, sup_a, fin_a) == oracle.set_search(one_b, text1)["count"]
    capfd.readouterr()
    # multiwm / multiwm2 + cuda_wm
    case = RANDOM_CASES[1]
    name, algo, alphabet, p, m, n, opts = case
    pattern, text = make_case(case)
    n = text.size
    pattern2 = np.ascontiguousarray(pattern).reshape(-1)
    want = oracle.set_search(pattern, text)["count"]
    SHIFT, PV, PI, PS = sm.alloc_wu_tables(m, p, alphabet)
    sm.preproc_wu(pattern, m, p, alphabet, 3, SHIFT, PV, PI, PS)
    assert sm.search_wu(pattern, m, p, text, n, SHIFT, PV, PI, PS) == want
    SHIFT2, PV2, PI2, PS2 = sm.alloc_wu_tables(m, p, alphabet)
    sm.preproc_wu2(pattern2, m, p, alphabet, 3, SHIFT2, PV2, PI2, PS2)
    assert sm.search_wu2(pattern2, m, p, text, n, SHIFT2, PV2, PI2, PS2) == want
    cnt, secs = sm.cuda_wm(5, pattern2, m, text, n, p, alphabet, 3, SHIFT2, PV2, PI2, PS2)
    assert cnt == want and secs > 0


def test_baseline_configs_full_size(acwm, oracle, have_ref, torch_cuda):
    """BASELINE configs 1 and 2 at their real size (128 MiB DNA): bit-exact against the
    linear-time oracle, against the unmodified reference's count on a prefix, and through
    size-independent properties (sorted, distinct, AC == WM on the same set, shard sums)."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    n = 128 << 20
    text = dg.text_host(n, 4, 1)
    d_text = torch.from_numpy(text).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for algo, p, m in ((acwm.AC, 100, 8), (acwm.WM, 1000, 16)):
        pats = dg.patterns_with_hits(text, p, m, 4, 2)
        ref = oracle.set_search(pats, text)
        mt = acwm.Matcher(algo, pats, 4).upload(pos_capacity=1 << 22)
        mt.scan_tensor(d_text)
        count, pos, _ = mt.fetch(cap=1 << 22, stream=st)
        assert count == ref["count"]
        assert np.array_equal(pos, ref["positions"])
        assert np.all(pos[1:] > pos[:-1])
        other = acwm.Matcher(acwm.WM if algo == acwm.AC else acwm.AC, pats, 4).upload(pos_capacity=1 << 22)
        other.scan_tensor(d_text)
        c2, p2, _ = other.fetch(cap=1 << 22, stream=st)
        assert c2 == count and np.array_equal(p2, pos)
        if have_ref:
            pre = 8 << 20
            rc = (oracle.ref_ac if algo == acwm.AC else oracle.ref_wu)(pats, 4, text[:pre])["count"]
            assert rc == int(np.count_nonzero(pos < pre))
        # host path, end to end
        c3, p3 = mt.search_host(text, cap=1 << 22)
        assert c3 == count and np.array_equal(p3, pos)
        mt.close()
        other.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_device_sharded_count_exchange_c_abi(acwm, oracle, torch_cuda, world):
    """acwm_peers_create + acwm_scan_device_sharded + acwm_fetch_sharded: the multi-rank flow for device-resident shards
    in one process through the C ABI alone.  The per-shard counts cross between the matchers through the mailboxes
    the scan kernels write (peer-mapped over NVLink when the shards sit on several GPUs; on a one-GPU box every
    matcher lives on GPU 0, which still runs acwm_set_peers, the in-kernel publish and the collect): the exchanged sum,
    the per-shard counts and every position equal the oracle's, over several rounds of back-to-back scans."""
    torch = torch_cuda
    n_dev = acwm.device_count()
    for cname in ("c1_ac_dna_p100_m8", "c2_wm_dna_p1000_m16", "c4_wm_ascii_mixed_8_64"):
        case = next(c for c in RANDOM_CASES if c[0] == cname)
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        texts = [text, np.roll(text, 12345)]
        refs = [oracle.set_search(pats, t) for t in texts]
        m_max = max(q.size for q in pats) if isinstance(pats, list) else pats.shape[1]
        mts = [acwm.Matcher(algo, pats, alphabet, **opts).upload(device=r % n_dev, pos_capacity=text.size) for r in range(world)]
        acwm.peers_create(mts)
        bounds = [acwm.shard_bounds(text.size, world, r, m_max - 1) for r in range(world)]
        shards = [[torch.from_numpy(t[s:s + l]).to(f"cuda:{r % n_dev}") for r, (s, l) in enumerate(bounds)] for t in texts]
        for rounds in (1, 3):
            for last in (0, 1):
                for k in range(rounds):  # back-to-back collective scans, no host synchronisation in between
                    acwm.scan_device_sharded(mts, shards[(last + k + 1) % 2])
                acwm.scan_device_sharded(mts, shards[last])
                g, per = acwm.fetch_sharded(mts)
                ref = refs[last]
                assert g == ref["count"] == int(per.sum()), (cname, world, rounds, last)
                got = []
                for r, (start, length) in enumerate(bounds):
                    torch.cuda.set_device(r % n_dev)
                    c, pos, _ = mts[r].fetch(cap=text.size)
                    assert c == int(per[r]) == pos.size
                    got.append(pos + np.uint64(start))
                assert np.array_equal(np.concatenate(got), ref["positions"]), (cname, world, rounds, last)
        acwm.peers_destroy(mts)
        torch.cuda.set_device(0)
        for mt in mts:
            mt.close()


def _planted(torch, d_text, pats, ends):
    """Overwrite d_text so that pattern k ends at ends[k]; returns the sorted ends."""
    for k, e in enumerate(ends):
        q = np.ascontiguousarray(pats[k])
        d_text[e - q.size + 1:e + 1] = torch.from_numpy(q).cuda()
    return np.array(sorted(ends), np.uint64)


def test_baseline_configs_3_and_4_full_size_by_planting(acwm, torch_cuda):
    """BASELINE configs[2] (one GPU's shard: 1e9 bytes of DNA, 100 000 patterns of 32) and configs[3] (2e9 bytes over 256
    symbols, 10 000 patterns of 8..64 bytes) at their real sizes.  On texts this random no pattern occurs by chance
    (1e9 * 1e5 / 4^32 ~ 5e-6; 2e9 * 1e4 / 256^8 ~ 1e-9), so the result must be EXACTLY the planted occurrences: count,
    every position, ascending.  Plus: AC == WM on the same set, a second scan gives the same result, and two shards
    with an (m_max-1)-byte halo report every match exactly once (main.c:467-477)."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    sh = __import__("acwm_pkg").submodule("sharding")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(21)
    T = 3584  # warp tile: plant around its edges too

    # ---- configs[2]: equal-length DNA patterns, AC (automaton in L2) and WM
    n = 1_000_000_000
    d_text = dg.text_device(n, 4, 11)
    pats = rng.integers(0, 4, (100_000, 32), dtype=np.uint8)
    # windows must not overlap (a later plant would overwrite an earlier one): >= 32 apart
    fixed = [31, T - 1, T + 31, 5 * T + 17, n // 2 - 1, n // 2 + 40, n - 1 - 64, n - 1]
    rnd = [int(x) for x in rng.integers(10_000, n - 10_000, 40) // 100 * 100 + 99]
    ends = sorted(set(fixed + [e for e in rnd if all(abs(e - f) > 200 for f in fixed)]))
    want = _planted(torch, d_text, pats[:len(ends)], ends)
    results = []
    for algo in (acwm.AC, acwm.WM):
        mt = acwm.Matcher(algo, pats, 4).upload(pos_capacity=1 << 20)
        for rep in range(2):
            mt.scan_tensor(d_text)
            count, pos, _ = mt.fetch(cap=1 << 20, stream=st)
            assert count == want.size and np.array_equal(pos, want), (algo, rep)
        got = []
        for r in range(2):
            start, length, report_from = sh.shard_of(n, 2, r, 32)
            mt.scan_tensor(d_text[start:start + length], report_from=report_from)
            c, ppos, _ = mt.fetch(cap=1 << 20, stream=st)
            got.append(ppos + np.uint64(start))
        assert np.array_equal(np.concatenate(got), want), algo
        results.append(pos)
        mt.close()
    assert np.array_equal(results[0], results[1])
    del d_text

    # ---- configs[3]: mixed-length byte patterns (WM; the reference has no mixed-length mode: union over lengths)
    n = 2_000_000_000
    d_text = dg.text_device(n, 256, 12)
    lens = rng.integers(8, 65, 10_000)
    mixed = [rng.integers(0, 256, int(L), dtype=np.uint8) for L in lens]
    fixed = [63, T - 1, T + 70, n // 2 + 5, n - 1]
    rnd = [int(x) for x in rng.integers(10_000, n - 10_000, 55) // 200 * 200 + 150]
    ends = sorted(set(fixed + [e for e in rnd if all(abs(e - f) > 200 for f in fixed)]))
    want = _planted(torch, d_text, mixed[:len(ends)], ends)
    mt = acwm.Matcher(acwm.WM, mixed, 256).upload(pos_capacity=1 << 20)
    for rep in range(2):
        mt.scan_tensor(d_text)
        count, pos, _ = mt.fetch(cap=1 << 20, stream=st)
        assert count == want.size and np.array_equal(pos, want), rep
    got = []
    for r in range(2):
        start, length, report_from = sh.shard_of(n, 2, r, 64)
        mt.scan_tensor(d_text[start:start + length], report_from=report_from)
        c, ppos, _ = mt.fetch(cap=1 << 20, stream=st)
        got.append(ppos + np.uint64(start))
    assert np.array_equal(np.concatenate(got), want)
    mt.close()


def test_multi_gib_text_positions_beyond_32_bits(acwm, torch_cuda):
    """5 GiB DNA text resident in HBM: 64-bit positions, AC == WM, planted matches found."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    n = 5 << 30
    d_text = dg.text_device(n, 4, 9)
    rng = np.random.default_rng(3)
    pats = rng.integers(0, 4, (64, 24), dtype=np.uint8)
    # ends on both sides of a warp-tile edge and of the 2^32 boundary; >= 24 apart so the planted
    # windows do not overwrite each other
    plant = [23, 3583, 3584 + 30, 2 * 3584, (1 << 32) - 1, (1 << 32) + 29, n - 1]
    for k, e in enumerate(plant):
        d_text[e - 23:e + 1] = torch.from_numpy(pats[k]).cuda()
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for algo in (acwm.AC, acwm.WM):
        mt = acwm.Matcher(algo, pats, 4).upload(pos_capacity=1 << 20)
        mt.scan_tensor(d_text)
        count, pos, _ = mt.fetch(cap=1 << 20, stream=st)
        assert count == pos.size
        assert set(plant) <= set(pos.tolist())
        res.append(pos)
        mt.close()
    assert np.array_equal(res[0], res[1])


def test_lookback_timeout_is_reported_not_hung(acwm, torch_cuda, tmp_path):
    """The span look-back of the position ordering waits for the totals of the spans in front of it; a predecessor that
    never publishes (fault injected through ACWM_TUNE bit 1: CTA 0 hides its total, the wait is cut to 20 ms) must end
    in an error from acwm_fetch -- count still exact -- and not in a hung stream.  Own process: the library reads
    ACWM_TUNE once."""
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, torch, acwm_pkg
acwm = acwm_pkg.load(); dg = acwm_pkg.submodule("datagen")
n = 8 << 20
text = dg.text_host(n, 4, 1)
pats = dg.patterns_with_hits(text, 100, 8, 4, 2)
mt = acwm.Matcher(acwm.AC, pats, 4)
mt.upload(pos_capacity=n)
d = torch.from_numpy(text).cuda()
mt.scan_tensor(d)
try:
    mt.fetch(cap=n, stream=torch.cuda.current_stream().cuda_stream)
    print("NO ERROR")
except acwm.AcwmError as e:
    print("ERROR", e.code, str(e))
torch.cuda.synchronize()
print("DONE")
""" % (str(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))),)
    env = dict(__import__("os").environ, ACWM_TUNE="0x0303")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env).stdout
    assert "DONE" in out and "ERROR %d" % acwm.ERR_CUDA in out and "position ordering gave up" in out, out


def test_host_search_stays_exact_while_it_adapts(acwm, oracle, torch_cuda):
    """acwm_search_host measures its hybrid and its plain transfer on its first calls, keeps the faster, and lets the raw
    share of the hybrid one climb on the measured call times: every one of those calls returns the oracle's matches
    (pinned text, so the hybrid path is open; the text changes from call to call)."""
    torch = torch_cuda
    dg = __import__("acwm_pkg").submodule("datagen")
    n = 96 << 20  # 7 chunks: enough for the hybrid split
    texts = [dg.text_host(n, 4, 40 + k) for k in range(2)]
    pats = dg.patterns_with_hits(texts[0], 200, 12, 4, 9)
    refs = [oracle.set_search(pats, t) for t in texts]
    pinned = [torch.from_numpy(t).pin_memory() for t in texts]
    mt = acwm.Matcher(acwm.WM, pats, 4)
    h2d = set()
    for i in range(14):
        count, pos = mt.search_host(pinned[i % 2], cap=max(1024, refs[i % 2]["count"]))
        assert count == refs[i % 2]["count"], i
        assert np.array_equal(pos, refs[i % 2]["positions"]), i
        h2d.add(int(mt.last_h2d_bytes))
    assert len(h2d) >= 2  # both transfers (and more than one split) were exercised
    mt.close()
