"""Seeded parity cases shared by the CPU (emulator) tests and the GPU tests.

Each case: (name, algo, alphabet, p, m, n, opts).  m may be a (lo, hi) tuple for
mixed-length WM sets (BASELINE config 4).  The edge cases follow SURVEY.md section 4:
overlapping occurrences, duplicates in the pattern set, matches at column m-1 and n-1,
no match, every position matches, alphabets 2/4/8/20/128/256."""
import numpy as np

import acwm_pkg

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")
AC, WM = acwm.AC, acwm.WM

RANDOM_CASES = [
    ("c1_ac_dna_p100_m8", AC, 4, 100, 8, 300_000, {}),
    ("c2_wm_dna_p1000_m16", WM, 4, 1000, 16, 300_000, {}),
    ("ac_dna_p1000_m16_trunc", AC, 4, 1000, 16, 200_000, {}),
    ("ac_dna_k2", AC, 4, 100, 8, 100_000, dict(force_stride=2)),
    ("ac_dna_k1", AC, 4, 100, 8, 100_000, dict(force_stride=1)),
    ("ac_dna_depth5", AC, 4, 100, 8, 100_000, dict(force_depth=5)),
    ("ac_dna_p5000_m32", AC, 4, 5000, 32, 150_000, {}),
    ("ac_dna_m70", AC, 4, 3, 70, 100_000, {}),
    ("ac_bin_p10_m6", AC, 2, 10, 6, 50_000, {}),
    ("wm_dna_p100_m8", WM, 4, 100, 8, 200_000, {}),
    ("wm_dna_p10_m3", WM, 4, 10, 3, 50_000, {}),
    ("wm_dna_p20000_m32", WM, 4, 20000, 32, 200_000, {}),
    ("wm_dna_p50_m64", WM, 4, 50, 64, 100_000, {}),
    ("wm_bin_p10_m12", WM, 2, 10, 12, 50_000, {}),
    ("wm_dna_s1", WM, 4, 100, 8, 100_000, dict(force_stride=1)),
    ("wm_dna_s4", WM, 4, 1000, 16, 100_000, dict(force_stride=4)),
    ("wm_dna_s16", WM, 4, 100, 20, 100_000, dict(force_stride=16)),
    ("wm_dna_mixed_8_64", WM, 4, 300, (8, 64), 150_000, {}),
    ("ac_ascii_p100_m8", AC, 256, 100, 8, 150_000, {}),
    ("ac_protein_p100_m6", AC, 20, 100, 6, 100_000, {}),
    ("ac_oct_p50_m6", AC, 8, 50, 6, 100_000, {}),
    ("ac_english_p300_m4", AC, 128, 300, 4, 100_000, {}),
    ("ac_ascii_p3000_m8", AC, 256, 3000, 8, 100_000, {}),
    ("ac_dna_bytes_path", AC, 4, 100, 8, 100_000, dict(force_bytes_path=1)),
    ("wm_ascii_p1000_m8", WM, 256, 1000, 8, 200_000, {}),
    ("wm_protein_p200_m6", WM, 20, 200, 6, 100_000, {}),
    ("wm_oct_p100_m5", WM, 8, 100, 5, 100_000, {}),
    ("wm_english_p100_m3", WM, 128, 100, 3, 100_000, {}),
    ("wm_dna_bytes_path", WM, 4, 100, 8, 100_000, dict(force_bytes_path=1)),
    ("c4_wm_ascii_mixed_8_64", WM, 256, 2000, (8, 64), 200_000, {}),
    # large sets: stage-1 bitmap / offset masks / stage-2 bitmap in global memory (L2-resident)
    ("c3_wm_dna_p100000_m32_l2", WM, 4, 100000, 32, 200_000, {}),
    ("wm_dna_p10000_m16_l2", WM, 4, 10000, 16, 200_000, {}),
    ("wm_dna_p10000_m16_smem_only", WM, 4, 10000, 16, 100_000, dict(force_smem_tables=1)),
    ("c4_wm_ascii_p10000_mixed_l2", WM, 256, 10000, (8, 64), 200_000, {}),
    ("wm_ascii_p10000_m12_smem_only", WM, 256, 10000, 12, 100_000, dict(force_smem_tables=1)),
    ("c3_ac_dna_p100000_m32", AC, 4, 100000, 32, 60_000, {}),
    ("ac_dna_p10000_m16_f2_l2", AC, 4, 10000, 16, 100_000, {}),
    # AC front ends pinned down: the automaton walks every symbol (force_front=1: shared memory, truncated, or the
    # L2-resident one) / a sampled block filter in front and the automaton decides candidate windows (force_front=2)
    ("ac_dna_p1000_m16_walk", AC, 4, 1000, 16, 200_000, dict(force_front=1)),
    ("c3_ac_dna_p100000_m32_walk_l2", AC, 4, 100000, 32, 60_000, dict(force_front=1)),
    ("ac_dna_p10000_m16_walk_l2", AC, 4, 10000, 16, 100_000, dict(force_front=1)),
    ("c1_ac_dna_p100_m8_filtered", AC, 4, 100, 8, 300_000, dict(force_front=2)),
    ("ac_dna_p1000_m16_filtered", AC, 4, 1000, 16, 200_000, dict(force_front=2)),
    ("ac_dna_p300_m3_filtered", AC, 4, 40, 3, 100_000, dict(force_front=2)),
    ("ac_bin_p10_m12_filtered", AC, 2, 10, 12, 50_000, dict(force_front=2)),
    # one CTA per SM / two half-size CTAs per SM
    ("c1_ac_one_cta", AC, 4, 100, 8, 300_000, dict(force_ctas=1)),
    ("c2_wm_one_cta", WM, 4, 1000, 16, 300_000, dict(force_ctas=1)),
    ("c1_ac_two_ctas", AC, 4, 100, 8, 300_000, dict(force_ctas=2)),
    ("c2_wm_two_ctas", WM, 4, 1000, 16, 300_000, dict(force_ctas=2)),
]


def make_case(case, seed=11):
    """-> (patterns, text).  Half the patterns are windows of the text; a duplicate, a match
    ending at column m-1 and one ending at n-1 are planted."""
    name, algo, alphabet, p, m, n, opts = case
    text = dg.text_host(n, alphabet, seed)
    if isinstance(m, tuple):
        pats = dg.mixed_patterns_with_hits(text, p, m[0], m[1], alphabet, seed + 1)
        pats.append(pats[0].copy())                   # duplicate
        pats.append(pats[1][-m[0]:].copy())           # a pattern that is a suffix of another
        text[:pats[2].size] = pats[2]                 # ends at column len-1
        text[n - pats[3].size:] = pats[3]             # ends at column n-1
    else:
        pats = dg.patterns_with_hits(text, p, m, alphabet, seed + 1)
        if p > 1:
            pats[1] = pats[0]
        text[:m] = pats[2 % p]
        text[n - m:] = pats[3 % p]
    return pats, text


def edge_cases():
    """Small hand-checkable inputs: (name, alphabet, patterns (p,m), text, expected positions)."""
    A = np.uint8
    out = []
    # overlapping occurrences: AAAA in AAAAAAAA -> ends 3..7
    out.append(("overlap", 4, np.zeros((1, 4), A), np.zeros(8, A), [3, 4, 5, 6, 7]))
    # every position matches: all 3-symbol binary patterns
    out.append(("all_match", 2, np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1],
                                          [1, 1, 0], [1, 1, 1]], A),
                np.array([0, 1, 1, 0, 1, 0, 0, 1, 1, 1], A), list(range(2, 10))))
    # no match
    out.append(("no_match", 4, np.array([[3, 3, 3, 3]], A), np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], A), []))
    # duplicates in the pattern set count once
    out.append(("duplicates", 4, np.array([[0, 1, 2], [0, 1, 2], [0, 1, 2]], A),
                np.array([0, 1, 2, 0, 1, 2, 3], A), [2, 5]))
    # match at column m-1 and at n-1 only
    out.append(("borders", 4, np.array([[1, 2, 3, 0]], A), np.array([1, 2, 3, 0, 2, 2, 1, 2, 3, 0], A), [3, 9]))
    # text shorter than the pattern
    out.append(("short_text", 4, np.array([[0, 1, 2, 3, 0]], A), np.array([0, 1, 2], A), []))
    # text exactly one pattern long
    out.append(("exact_len", 4, np.array([[0, 1, 2, 3, 0]], A), np.array([0, 1, 2, 3, 0], A), [4]))
    # bytes alphabet, pattern sharing prefixes/suffixes
    out.append(("ascii_share", 256, np.array([list(b"abcab"), list(b"bcabc"), list(b"cabca")], A),
                np.frombuffer(b"abcabcabcabxabcab", A).copy(), [4, 5, 6, 7, 8, 9, 10, 16]))
    return out
