"""The oracle itself: restatement (oracle_port.c) vs the committed golden vectors (whose
counts came from the unmodified reference) and, where the compiled reference is present
(oracle/_ref), vs the reference directly -- counts and reference-layout tables."""
import numpy as np
import pytest

from cases import RANDOM_CASES, edge_cases, make_case
from golden_util import load_golden

GOLD = load_golden()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_port_matches_golden(oracle, name):
    g = GOLD[name]
    pats, text, alphabet = g["patterns"], g["text"], int(g["alphabet"])
    m = pats.shape[1]
    exp = g["positions"]
    assert int(g["ref_ac_count"]) == int(g["ref_wu_count"]) == exp.size
    a = oracle.port_ac(pats, alphabet, text)
    assert a["count"] == exp.size and np.array_equal(a["positions"], exp)
    if m >= 3 and oracle.port().oracle_wu_shiftsize(alphabet):
        w = oracle.port_wu(pats, alphabet, text)
        assert w["count"] == exp.size and np.array_equal(w["positions"], exp)
    s = oracle.set_search(pats, text)
    assert s["count"] == exp.size and np.array_equal(s["positions"], exp)
    if text.size * pats.shape[0] * m <= 5e7:
        nv = oracle.naive(pats, text)
        assert np.array_equal(nv["positions"], exp)
    if g.get("ref_states"):
        assert a["n_states"] == int(g["ref_states"]) and a["n_distinct"] == int(g["ref_distinct"])


@pytest.mark.parametrize("case", [c for c in RANDOM_CASES if not isinstance(c[4], tuple)][::3],
                         ids=lambda c: c[0])
def test_port_matches_reference(oracle, have_ref, case):
    if not have_ref:
        pytest.skip("compiled reference (oracle/_ref) not present")
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    text = text[:60_000]
    ra = oracle.ref_ac(pats, alphabet, text, want_tables=True)
    pa = oracle.port_ac(pats, alphabet, text, want_tables=True)
    assert ra["count"] == pa["count"]
    assert ra["n_states"] == pa["n_states"] and ra["n_distinct"] == pa["n_distinct"]
    for k in ("state_transition", "state_supply", "state_final"):
        assert np.array_equal(ra[k], pa[k]), k
    if m >= 3 and alphabet in (2, 4, 8, 20, 128, 256):
        rw = oracle.ref_wu(pats, alphabet, text, want_tables=True)
        rw2 = oracle.ref_wu(pats, alphabet, text, flat=True)
        pw = oracle.port_wu(pats, alphabet, text, want_tables=True)
        assert rw["count"] == rw2["count"] == pw["count"] == ra["count"]
        for k in ("SHIFT", "PREFIX_size", "PREFIX_value", "PREFIX_index"):
            assert np.array_equal(rw[k], pw[k]), k
        assert np.array_equal(pw["positions"], pa["positions"])


def test_reference_sharding_sums(oracle, have_ref):
    """main.c:467-477 shards (threads = MPI ranks) lose and duplicate nothing."""
    if not have_ref:
        pytest.skip("compiled reference (oracle/_ref) not present")
    pats, text = make_case(RANDOM_CASES[0])
    text = text[:200_000]
    whole = oracle.ref_ac(pats, 4, text)["count"]
    for threads in (2, 3, 4, 8):
        assert oracle.ref_ac(pats, 4, text, threads=threads)["count"] == whole
        assert oracle.ref_wu(pats, 4, text, threads=threads)["count"] == whole


def test_mixed_length_definition(oracle):
    """Mixed lengths = union over length classes of the equal-length result."""
    case = next(c for c in RANDOM_CASES if c[0] == "wm_dna_mixed_8_64")
    pats, text = make_case(case)
    text = text[:50_000]
    s = oracle.set_search(pats, text)
    per_len = {}
    for q in pats:
        per_len.setdefault(q.size, []).append(q)
    allpos = []
    for L, group in per_len.items():
        r = oracle.port_ac(np.stack(group), 4, text)
        allpos.append(r["positions"])
    allpos = np.sort(np.concatenate(allpos))
    assert s["count"] == allpos.size and np.array_equal(s["positions"], allpos)


def test_edge_cases_hand_checked(oracle):
    for name, alphabet, pats, text, exp in edge_cases():
        assert oracle.naive(pats, text)["positions"].tolist() == exp, name
        assert oracle.port_ac(pats, alphabet, text)["positions"].tolist() == exp, name
