"""Small all-kernel-family sanity run (used under compute-sanitizer and as a first hang check)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import acwm_pkg, oracle
from cases import RANDOM_CASES, make_case
acwm = acwm_pkg.load()
names = sys.argv[1:] or ["c1_ac_dna_p100_m8", "c2_wm_dna_p1000_m16", "ac_dna_depth5", "wm_dna_mixed_8_64",
                         "ac_ascii_p100_m8", "ac_protein_p100_m6", "wm_ascii_p1000_m8", "c4_wm_ascii_mixed_8_64"]
bad = 0
for case in RANDOM_CASES:
    if case[0] not in names:
        continue
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    text = text[:40_000]
    t0 = time.time()
    mt = acwm.Matcher(algo, pats, alphabet, **opts)
    count, pos = mt.search_host(text, cap=text.size * 2)
    ref = oracle.set_search(pats, text)
    ok = count == ref["count"] and np.array_equal(pos, ref["positions"])
    bad += not ok
    print(("OK  " if ok else "BAD ") + name, count, ref["count"], f"{time.time()-t0:.2f}s", flush=True)
    mt.close()
sys.exit(1 if bad else 0)
