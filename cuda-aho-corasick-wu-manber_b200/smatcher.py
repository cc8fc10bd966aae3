"""Python mirror of the reference's smatcher.h interface for the AC / WM path.

Same function names, argument order and meaning as the reference
(smatcher.h:89-91,101-106; cuda/cuda_ac.cu:594; cuda/cuda_wm.cu:183), bound with ctypes
to the reference-shaped shims exported by libacwm_b200.so -- i.e. this is exactly the
binding a maintainer of the reference would write.  The caller owns and
pre-initialises the flat tables just like main.c:410-449 does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint)


class ac_table(C.Structure):  # smatcher.h:49-53
    _fields_ = [("idcounter", C.c_uint), ("patterncounter", C.c_uint), ("zerostate", C.c_void_p)]


class sbom_table(C.Structure):  # smatcher.h:65-69
    _fields_ = [("idcounter", C.c_uint), ("patterncounter", C.c_uint), ("zerostate", C.c_void_p)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    pp = C.POINTER(_u8p)
    L.preproc_ac.restype = C.POINTER(ac_table)
    L.preproc_ac.argtypes = [pp, C.c_int, C.c_int, C.c_int, _i32p, _u32p, _u32p]
    L.search_ac.restype = C.c_uint
    L.search_ac.argtypes = [_u8p, C.c_int, C.POINTER(ac_table)]
    L.free_ac.restype = None
    L.free_ac.argtypes = [C.POINTER(ac_table), C.c_int]
    L.wu_determine_shiftsize.restype = None
    L.wu_determine_shiftsize.argtypes = [C.c_int]
    wu_tabs = [_i32p, _i32p, _i32p, _i32p]
    L.preproc_wu.restype = None
    L.preproc_wu.argtypes = [pp, C.c_int, C.c_int, C.c_int, C.c_int] + wu_tabs
    L.preproc_wu2.restype = None
    L.preproc_wu2.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int] + wu_tabs
    L.search_wu.restype = C.c_uint
    L.search_wu.argtypes = [pp, C.c_int, C.c_int, _u8p, C.c_int] + wu_tabs
    L.search_wu2.restype = C.c_uint
    L.search_wu2.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int] + wu_tabs
    for k in range(1, 6):
        f = getattr(L, f"cuda_ac{k}")
        f.restype = None
        f.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, _i32p, _u32p, _u32p]
        g = getattr(L, f"cuda_wm{k}")
        g.restype = C.c_int
        g.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int] + wu_tabs + [C.POINTER(C.c_double)]
    L.acwm_shim_last_count.restype = C.c_ulonglong
    # sibling algorithms (smatcher.h:93-99,109-110)
    L.preproc_sh.restype = C.POINTER(ac_table)
    L.preproc_sh.argtypes = [pp, C.c_int, C.c_int, C.c_int, _i32p, _u32p]
    L.search_sh.restype = C.c_uint
    L.search_sh.argtypes = [C.c_int, _u8p, C.c_int, C.POINTER(ac_table), _i32p]
    L.free_sh.restype = None
    L.free_sh.argtypes = [C.POINTER(ac_table), C.c_int]
    L.preproc_sbom.restype = C.POINTER(sbom_table)
    L.preproc_sbom.argtypes = [pp, C.c_int, C.c_int, C.c_int, _i32p, _u32p]
    L.search_sbom.restype = C.c_uint
    L.search_sbom.argtypes = [pp, C.c_int, _u8p, C.c_int, C.POINTER(sbom_table)]
    L.free_sbom.restype = None
    L.free_sbom.argtypes = [C.POINTER(sbom_table), C.c_int]
    sog_tabs = [_u8p, _u32p, _i32p, _u8p]
    L.preproc_sog8.restype = None
    L.preproc_sog8.argtypes = sog_tabs + [pp, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
    L.search_sog8.restype = C.c_uint
    L.search_sog8.argtypes = sog_tabs + [pp, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
    L.acwm_shim_forget.restype = None
    L.acwm_shim_forget.argtypes = [C.c_void_p]
    _bound = True
    return L


def _rows(pattern: np.ndarray):
    """unsigned char *pattern[p_size] over a (p, m) uint8 array (kept alive by the caller)."""
    p = pattern.shape[0]
    arr = (_u8p * p)()
    for j in range(p):
        arr[j] = pattern[j].ctypes.data_as(_u8p)
    return arr


def _p(a, t):
    return a.ctypes.data_as(t)


# ---- Aho-Corasick (smatcher.h:89-91)
def preproc_ac(pattern, m, p_size, alphabet, state_transition, state_supply, state_final):
    return _bind().preproc_ac(_rows(pattern), m, p_size, alphabet, _p(state_transition, _i32p),
                              _p(state_supply, _u32p), _p(state_final, _u32p))


def search_ac(text, n, table):
    return int(_bind().search_ac(_p(text, _u8p), n, table))


def free_ac(table, alphabet):
    _bind().free_ac(table, alphabet)


# ---- Wu-Manber (smatcher.h:101-106)
def wu_determine_shiftsize(alphabet):
    L = _bind()
    L.wu_determine_shiftsize(alphabet)
    return int(C.c_uint.in_dll(L, "shiftsize").value)


def set_m_nBitsInShift(v: int):
    C.c_ushort.in_dll(_bind(), "m_nBitsInShift").value = v


def preproc_wu(pattern, m, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size):
    _bind().preproc_wu(_rows(pattern), m, p_size, alphabet, B, _p(SHIFT, _i32p), _p(PREFIX_value, _i32p),
                       _p(PREFIX_index, _i32p), _p(PREFIX_size, _i32p))


def preproc_wu2(pattern2, m, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size):
    _bind().preproc_wu2(_p(pattern2, _u8p), m, p_size, alphabet, B, _p(SHIFT, _i32p), _p(PREFIX_value, _i32p),
                        _p(PREFIX_index, _i32p), _p(PREFIX_size, _i32p))


def search_wu(pattern, m, p_size, text, n, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size):
    return int(_bind().search_wu(_rows(pattern), m, p_size, _p(text, _u8p), n, _p(SHIFT, _i32p),
                                 _p(PREFIX_value, _i32p), _p(PREFIX_index, _i32p), _p(PREFIX_size, _i32p)))


def search_wu2(pattern2, m, p_size, text, n, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size):
    return int(_bind().search_wu2(_p(pattern2, _u8p), m, p_size, _p(text, _u8p), n, _p(SHIFT, _i32p),
                                  _p(PREFIX_value, _i32p), _p(PREFIX_index, _i32p), _p(PREFIX_size, _i32p)))


# ---- GPU wrappers (cuda/cuda_ac.cu:594.., cuda/cuda_wm.cu:183..)
def cuda_ac(variant, m, text, n, p_size, alphabet, state_transition, state_supply, state_final):
    """cuda_ac<variant>: prints the reference's "Kernel N matches" line; returns the count it printed."""
    L = _bind()
    getattr(L, f"cuda_ac{variant}")(m, _p(text, _u8p), n, p_size, alphabet, _p(state_transition, _i32p),
                                    _p(state_supply, _u32p), _p(state_final, _u32p))
    return int(L.acwm_shim_last_count())


def cuda_wm(variant, pattern2, m, text, n, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size):
    """cuda_wm<variant>: returns (count, gpuTime seconds) like the reference's out-parameter."""
    t = C.c_double(0)
    c = getattr(_bind(), f"cuda_wm{variant}")(_p(pattern2, _u8p), m, _p(text, _u8p), n, p_size, alphabet, B,
                                              _p(SHIFT, _i32p), _p(PREFIX_value, _i32p), _p(PREFIX_index, _i32p),
                                              _p(PREFIX_size, _i32p), C.byref(t))
    return int(c), float(t.value)


# ---- sibling algorithms: Set-Horspool, SBOM, Shift-Or with q-grams (smatcher.h:93-99,109-110)
def preproc_sh(pattern, m, p_size, alphabet, state_transition, state_final):
    return _bind().preproc_sh(_rows(pattern), m, p_size, alphabet, _p(state_transition, _i32p), _p(state_final, _u32p))


def search_sh(m, text, n, table, bmBc):
    return int(_bind().search_sh(m, _p(text, _u8p), n, table, _p(bmBc, _i32p)))


def free_sh(table, alphabet):
    _bind().free_sh(table, alphabet)


def preproc_sbom(pattern, m, p_size, alphabet, state_transition, state_final_multi):
    return _bind().preproc_sbom(_rows(pattern), m, p_size, alphabet, _p(state_transition, _i32p),
                                _p(state_final_multi, _u32p))


def search_sbom(pattern, m, text, n, table):
    return int(_bind().search_sbom(_rows(pattern), m, _p(text, _u8p), n, table))


def free_sbom(table, m):
    _bind().free_sbom(table, m)


def alloc_sog8_tables(p_size):
    """T8 (2^24 bytes), scanner_hs, scanner_index, scanner_hs2 (8192 bytes): the allocation main.c keeps in comments."""
    return (np.empty(1 << 24, np.uint8), np.zeros(p_size, np.uint32), np.zeros(p_size, np.int32), np.zeros(8192, np.uint8))


def preproc_sog8(T8, scanner_hs, scanner_index, scanner_hs2, pattern, m, text, n, p_size, B=3):
    _bind().preproc_sog8(_p(T8, _u8p), _p(scanner_hs, _u32p), _p(scanner_index, _i32p), _p(scanner_hs2, _u8p),
                         _rows(pattern), m, _p(text, _u8p), n, p_size, B)


def search_sog8(T8, scanner_hs, scanner_index, scanner_hs2, pattern, m, text, n, p_size, B=3):
    return int(_bind().search_sog8(_p(T8, _u8p), _p(scanner_hs, _u32p), _p(scanner_index, _i32p), _p(scanner_hs2, _u8p),
                                   _rows(pattern), m, _p(text, _u8p), n, p_size, B))


def shim_forget(table):
    _bind().acwm_shim_forget(table.ctypes.data)


def alloc_ac_tables(m, p_size, alphabet):
    """Caller-side allocation + initialisation of main.c:410-420."""
    ns = m * p_size + 1
    return (np.full(ns * alphabet, -1, np.int32), np.zeros(ns, np.uint32), np.zeros(ns, np.uint32))


def alloc_wu_tables(m, p_size, alphabet, B=3):
    """Caller-side allocation + initialisation of main.c:429-449."""
    ss = wu_determine_shiftsize(alphabet)
    set_m_nBitsInShift(2)
    SHIFT = np.full(ss, m - B + 1, np.int32)
    return (SHIFT, np.full(ss * p_size, -7, np.int32), np.full(ss * p_size, -7, np.int32), np.zeros(ss, np.int32))
