"""Synthetic texts and pattern sets for tests and the bench harness.

Plays the role of the reference's missing ``load_files`` /
``create_multiple_pattern_with_hits`` helpers (main.c:49,453): uniform i.i.d. symbol
codes in [0, alphabet), one byte per symbol; pattern sets half sampled from the text
("with hits") and half uniform random.  Everything is seeded (numpy PCG64 on the host,
torch Philox on the device) so that every rank / every run sees the same bytes.
"""
from __future__ import annotations

import numpy as np


def text_host(n: int, alphabet: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, alphabet, size=n, dtype=np.uint8)


def text_device(n: int, alphabet: int, seed: int, device="cuda"):
    """Uniform symbols generated directly in HBM (for multi-GB texts)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 28
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = torch.randint(0, alphabet, (hi - lo,), dtype=torch.uint8, device=device, generator=g)
    return out


def patterns_with_hits(text, p: int, m: int, alphabet: int, seed: int, hit_fraction: float = 0.5) -> np.ndarray:
    """(p, m) uint8: the first ceil(p*hit_fraction) rows are windows of `text`, the rest uniform."""
    rng = np.random.default_rng(seed)
    pats = rng.integers(0, alphabet, size=(p, m), dtype=np.uint8)
    n = int(text.shape[0]) if hasattr(text, "shape") else len(text)
    n_hits = int(np.ceil(p * hit_fraction)) if n >= m else 0
    if n_hits:
        offs = rng.integers(0, n - m + 1, size=n_hits)
        for j, o in enumerate(offs.tolist()):
            w = text[o:o + m]
            pats[j] = w.cpu().numpy() if hasattr(w, "cpu") else np.asarray(w)
    return pats


def mixed_patterns_with_hits(text, p: int, m_lo: int, m_hi: int, alphabet: int, seed: int,
                             hit_fraction: float = 0.5):
    """List of p uint8 arrays with lengths uniform in [m_lo, m_hi] (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(m_lo, m_hi + 1, size=p)
    n = int(text.shape[0]) if hasattr(text, "shape") else len(text)
    out = []
    for j, L in enumerate(lens.tolist()):
        if j < int(np.ceil(p * hit_fraction)) and n >= L:
            o = int(rng.integers(0, n - L + 1))
            w = text[o:o + L]
            out.append((w.cpu().numpy() if hasattr(w, "cpu") else np.asarray(w)).astype(np.uint8).copy())
        else:
            out.append(rng.integers(0, alphabet, size=L, dtype=np.uint8))
    return out
