"""Small acwm_search_host_sharded run for compute-sanitizer: 3 shards over the devices present, AC and WM (mixed
lengths), checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import acwm_pkg  # noqa: E402
import oracle  # noqa: E402
from cases import RANDOM_CASES, make_case  # noqa: E402

acwm = acwm_pkg.load()
n_dev = acwm.device_count()
for cname in ("c1_ac_dna_p100_m8", "c4_wm_ascii_mixed_8_64"):
    case = next(c for c in RANDOM_CASES if c[0] == cname)
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    ref = oracle.set_search(pats, text)
    mts = [acwm.Matcher(algo, pats, alphabet, **opts).upload(device=r % n_dev) for r in range(3)]
    for rep in range(2):
        count, pos, per = acwm.search_host_sharded(mts, text, cap=max(1, ref["count"]))
        assert count == ref["count"] and np.array_equal(pos, ref["positions"]), cname
    for mt in mts:
        mt.close()
    print(cname, "ok", count, per.tolist())
