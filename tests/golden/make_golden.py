"""Generates the golden vectors under tests/golden/ (run HERE, where /root/reference
exists; the GPU box only reads the committed .npz files).

For every case the COUNT is produced by the UNMODIFIED reference (oracle/_ref:
preproc_ac+search_ac and preproc_wu+search_wu compiled from /root/reference) and the
POSITIONS by the oracle restatement (oracle/oracle_port.c), accepted only after
port count == reference AC count == reference WM count == naive/set count.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from cases import edge_cases  # noqa: E402

GOLDEN = [
    # name, alphabet, p, m, n, seed
    ("dna_p100_m8", 4, 100, 8, 65536, 101),
    ("dna_p1000_m16", 4, 1000, 16, 65536, 102),
    ("dna_p16_m4_dense", 4, 16, 4, 16384, 103),
    ("bin_p10_m6", 2, 10, 6, 16384, 104),
    ("oct_p50_m6", 8, 50, 6, 32768, 105),
    ("protein_p200_m5", 20, 200, 5, 32768, 106),
    ("english_p100_m3", 128, 100, 3, 32768, 107),
    ("ascii_p500_m8", 256, 500, 8, 65536, 108),
    ("dna_p200_m32", 4, 200, 32, 65536, 109),
]


def gen(alphabet, p, m, n, seed):
    rng = np.random.default_rng(seed)
    text = rng.integers(0, alphabet, n, dtype=np.uint8)
    pats = rng.integers(0, alphabet, (p, m), dtype=np.uint8)
    for j in range(0, p, 2):
        o = int(rng.integers(0, n - m + 1))
        pats[j] = text[o:o + m]
    if p > 3:
        pats[3] = pats[2]            # duplicate
    text[:m] = pats[0]               # match ending at column m-1
    text[n - m:] = pats[1]           # match ending at column n-1
    return pats, text


def main():
    assert oracle.ref_available(), "needs /root/reference (or a prebuilt oracle/_ref)"
    out = {}
    for name, alphabet, p, m, n, seed in GOLDEN:
        pats, text = gen(alphabet, p, m, n, seed)
        ra = oracle.ref_ac(pats, alphabet, text)
        rw = oracle.ref_wu(pats, alphabet, text)
        pa = oracle.port_ac(pats, alphabet, text)
        pw = oracle.port_wu(pats, alphabet, text)
        ss = oracle.set_search(pats, text)
        assert ra["count"] == rw["count"] == pa["count"] == pw["count"] == ss["count"], name
        assert np.array_equal(pa["positions"], pw["positions"]) and np.array_equal(pa["positions"], ss["positions"])
        out[name] = dict(alphabet=alphabet, patterns=pats, text=text, ref_ac_count=ra["count"],
                         ref_wu_count=rw["count"], ref_states=ra["n_states"], ref_distinct=ra["n_distinct"],
                         positions=pa["positions"])
        print(f"{name}: count {ra['count']} states {ra['n_states']}")
    for name, alphabet, pats, text, exp in edge_cases():
        m = pats.shape[1]
        cnt_ac = oracle.ref_ac(pats, alphabet, text)["count"] if text.size else 0
        cnt_wu = oracle.ref_wu(pats, alphabet, text)["count"] if m >= 3 else cnt_ac
        nv = oracle.naive(pats, text)
        assert cnt_ac == cnt_wu == nv["count"] == len(exp), (name, cnt_ac, cnt_wu, nv["count"], exp)
        assert nv["positions"].tolist() == exp, name
        out["edge_" + name] = dict(alphabet=alphabet, patterns=pats, text=text, ref_ac_count=cnt_ac,
                                   ref_wu_count=cnt_wu, ref_states=0, ref_distinct=0,
                                   positions=np.array(exp, np.uint64))
        print(f"edge_{name}: count {cnt_ac}")
    flat = {}
    for name, d in out.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **flat)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
