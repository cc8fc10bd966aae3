"""Sibling algorithms behind the same matcher: Set-Horspool, Set Backward Oracle
Matching and Shift-Or with q-grams, reached through the reference's own entry points (smatcher.h:93-99,109-110).

CPU tests: the flat tables the shims leave in the caller's arrays equal what the UNMODIFIED reference sh/sh.c and
sbom/sbom.c leave there (oracle/_ref/libref_siblings.so), and the reference's own counts equal the result definition
the GPU path is checked against.  GPU tests: the counts the shims return equal the compiled reference's."""
import numpy as np
import pytest

import acwm_pkg
import oracle

sm = None


def _sm():
    global sm
    if sm is None:
        sm = acwm_pkg.submodule("smatcher")
    return sm


CASES = [  # alphabet, m, p, n, seed
    (4, 8, 100, 1 << 18, 1), (4, 16, 1000, 1 << 18, 2), (4, 8, 20, 50_000, 3), (256, 8, 50, 1 << 18, 4),
    (20, 10, 200, 1 << 18, 5), (2, 12, 30, 100_000, 6), (4, 32, 300, 1 << 17, 7), (128, 5, 64, 1 << 17, 8),
]


def make_case(alphabet, m, p, n, seed):
    rng = np.random.default_rng(seed)
    text = rng.integers(0, alphabet, n, dtype=np.uint8)
    pats = rng.integers(0, alphabet, (p, m), dtype=np.uint8)
    for i in range(0, p, 2):  # half of them occur
        o = int(rng.integers(0, n - m))
        pats[i] = text[o:o + m]
    if p > 4:
        pats[3] = pats[2]  # a duplicate pattern (the tries collapse it, ac/ac.c:183)
    return text, np.ascontiguousarray(pats)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


needs_ref = pytest.mark.skipif(not oracle.siblings_available(), reason="oracle/_ref/libref_siblings.so not built")


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_reference_siblings_count_the_same_set(case):
    """search_sh / search_sbom of the reference return |M|: the definition every GPU result is held to."""
    alphabet, m, p, n, seed = case
    text, pats = make_case(*case)
    want = oracle.set_search(pats, text)["count"]
    assert oracle.ref_sh(pats, alphabet, text)["count"] == want
    assert oracle.ref_sbom(pats, alphabet, text)["count"] == want


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_sh_tables_equal_reference(case):
    alphabet, m, p, n, seed = case
    _, pats = make_case(*case)
    ref = oracle.ref_sh(pats, alphabet, want_tables=True)
    s = _sm()
    ns = m * p + 1
    tr, fin = np.full(ns * alphabet, -1, np.int32), np.zeros(ns, np.uint32)
    t = s.preproc_sh(pats, m, p, alphabet, tr, fin)
    try:
        assert t.contents.idcounter == ref["n_states"] and t.contents.patterncounter == ref["n_distinct"]
        assert np.array_equal(tr, ref["state_transition"])
        assert np.array_equal(fin, ref["state_final"])
    finally:
        s.free_sh(t, alphabet)


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_sbom_tables_equal_reference(case):
    alphabet, m, p, n, seed = case
    _, pats = make_case(*case)
    ref = oracle.ref_sbom(pats, alphabet, want_tables=True)
    s = _sm()
    ns = m * p + 1
    tr, fm = np.full(ns * alphabet, -1, np.int32), np.zeros(ns * 200, np.uint32)
    t = s.preproc_sbom(pats, m, p, alphabet, tr, fm)
    try:
        assert t.contents.idcounter == ref["n_states"] and t.contents.patterncounter == p
        assert np.array_equal(tr, ref["state_transition"])
        assert np.array_equal(fm, ref["state_final_multi"])
    finally:
        s.free_sbom(t, m)


def test_sog8_tables():
    """3-gram masks, sorted hashes and the two-level bitmap as sog/sog8.c:113-170 defines them."""
    text, pats = make_case(256, 8, 40, 50_000, 11)
    s = _sm()
    T8, hs, idx, hs2 = s.alloc_sog8_tables(40)
    s.preproc_sog8(T8, hs, idx, hs2, pats, 8, text, len(text), 40)
    try:
        want = np.full(1 << 24, 0xff, np.uint8)
        for q in pats:
            for k in range(6):
                want[int(q[k]) + (int(q[k + 1]) << 8) + (int(q[k + 2]) << 16)] &= 0xff - (1 << k)
        assert np.array_equal(T8, want)
        h = np.array([(int.from_bytes(bytes(q[:4]), "big") ^ int.from_bytes(bytes(q[4:]), "big")) for q in pats], np.uint32)
        assert np.array_equal(hs, np.sort(h)) and np.array_equal(h[idx], hs)
        bits = np.zeros(8192, np.uint8)
        for v in h:
            f = ((int(v) >> 16) ^ int(v)) & 0xffff
            bits[f >> 3] |= 1 << (f & 7)
        assert np.array_equal(hs2, bits)
    finally:
        s.shim_forget(T8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_sibling_searches_equal_reference(case, torch_cuda):
    alphabet, m, p, n, seed = case
    text, pats = make_case(*case)
    want = oracle.set_search(pats, text)["count"]
    if oracle.siblings_available():
        assert oracle.ref_sh(pats, alphabet, text)["count"] == want
        assert oracle.ref_sbom(pats, alphabet, text)["count"] == want
    s = _sm()
    ns = m * p + 1
    tr, fin = np.full(ns * alphabet, -1, np.int32), np.zeros(ns, np.uint32)
    t = s.preproc_sh(pats, m, p, alphabet, tr, fin)
    try:
        assert s.search_sh(m, text, len(text), t, np.zeros(alphabet, np.int32)) == want
    finally:
        s.free_sh(t, alphabet)
    tr, fm = np.full(ns * alphabet, -1, np.int32), np.zeros(ns * 200, np.uint32)
    t = s.preproc_sbom(pats, m, p, alphabet, tr, fm)
    try:
        assert s.search_sbom(pats, m, text, len(text), t) == want
    finally:
        s.free_sbom(t, m)
    if m == 8:
        T8, hs, idx, hs2 = s.alloc_sog8_tables(p)
        s.preproc_sog8(T8, hs, idx, hs2, pats, 8, text, len(text), p)
        try:
            assert s.search_sog8(T8, hs, idx, hs2, pats, 8, text, len(text), p) == want
        finally:
            s.shim_forget(T8)
