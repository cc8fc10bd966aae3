import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")


def load_golden():
    """{name: dict(alphabet, patterns, text, ref_ac_count, ref_wu_count, positions, ...)}"""
    z = np.load(_PATH)
    out = {}
    for key in z.files:
        name, field = key.split("/")
        v = z[key]
        out.setdefault(name, {})[field] = v if v.ndim else v.item()
    return out
