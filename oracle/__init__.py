"""TEST INFRASTRUCTURE ONLY -- ctypes bindings to the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package; the product (``cuda-aho-corasick-wu-manber_b200``) never does.

Two shared objects, both built by ``oracle/Makefile``:

* ``oracle/_ref/libref_oracle.so`` -- the UNMODIFIED reference ``ac/ac.c`` + ``wu/wu.c``
  (compiled where they lie under /root/reference) behind ``ref_driver.c``.  Returns
  match COUNTS only, like the reference (ac/ac.c:198-222, wu/wu.c:49-107).
* ``oracle/liboracle_port.so`` -- our restatement (``oracle_port.c``): same counts, plus
  match POSITIONS (``column`` convention of ac/ac.c:217 / wu/wu.c:93) and two
  window-membership scanners.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle_port.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
_SIB_SO = os.path.join(_HERE, "_ref", "libref_siblings.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def build(quiet: bool = True) -> None:
    """(Re)build both shared objects; the reference one only if /root/reference exists."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _as_u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            build()
        lib = C.CDLL(_PORT_SO)
        lib.oracle_ac_build.restype = C.c_void_p
        lib.oracle_ac_build.argtypes = [_u8p, C.c_int, C.c_int, C.c_int]
        lib.oracle_ac_free.argtypes = [C.c_void_p]
        lib.oracle_ac_states.restype = C.c_uint32
        lib.oracle_ac_states.argtypes = [C.c_void_p]
        lib.oracle_ac_distinct.restype = C.c_uint32
        lib.oracle_ac_distinct.argtypes = [C.c_void_p]
        lib.oracle_ac_export_ref_tables.argtypes = [C.c_void_p, C.c_int, _i32p, _u32p, _u32p]
        lib.oracle_ac_scan.restype = C.c_uint64
        lib.oracle_ac_scan.argtypes = [C.c_void_p, _u8p, C.c_uint64, _u64p, C.c_uint64]
        lib.oracle_wu_shiftsize.restype = C.c_uint32
        lib.oracle_wu_shiftsize.argtypes = [C.c_int]
        lib.oracle_wu_build.restype = C.c_void_p
        lib.oracle_wu_build.argtypes = [_u8p, C.c_int, C.c_int, C.c_int]
        lib.oracle_wu_free.argtypes = [C.c_void_p]
        lib.oracle_wu_export_ref_tables.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p]
        lib.oracle_wu_scan.restype = C.c_uint64
        lib.oracle_wu_scan.argtypes = [C.c_void_p, _u8p, C.c_uint64, _u64p, C.c_uint64]
        for name in ("oracle_naive_search", "oracle_set_search"):
            fn = getattr(lib, name)
            fn.restype = C.c_uint64
        lib.oracle_naive_search.argtypes = [_u8p, _u64p, _u32p, C.c_int, _u8p, C.c_uint64, _u64p, C.c_uint64]
        lib.oracle_set_search.argtypes = [_u8p, _u64p, _u32p, C.c_int, _u8p, C.c_uint64, _u64p, C.c_uint64, C.c_int]
        _port = lib
    return _port


def ref_available() -> bool:
    if os.path.exists(_REF_SO):
        return True
    if os.path.isdir("/root/reference/ac"):
        build()
        return os.path.exists(_REF_SO)
    return False


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libref_oracle.so missing (build it where /root/reference exists)")
        lib = C.CDLL(_REF_SO)
        lib.ref_ac_search.restype = C.c_ulonglong
        lib.ref_ac_search.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_uint64, C.c_int,
                                      _i32p, _u32p, _u32p, _u32p, _f64p]
        lib.ref_wu_shiftsize.restype = C.c_uint
        lib.ref_wu_shiftsize.argtypes = [C.c_int]
        lib.ref_wu_search.restype = C.c_ulonglong
        lib.ref_wu_search.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_uint64, C.c_int,
                                      _i32p, _i32p, _i32p, _i32p, _f64p]
        _ref = lib
    return _ref


_sib = None


def siblings_available() -> bool:
    if os.path.exists(_SIB_SO):
        return True
    if os.path.isdir("/root/reference/sh"):
        build()
        return os.path.exists(_SIB_SO)
    return False


def siblings():
    """oracle/_ref/libref_siblings.so: the UNMODIFIED reference sh/sh.c + sbom/sbom.c behind ref_siblings.c."""
    global _sib
    if _sib is None:
        if not siblings_available():
            raise RuntimeError("oracle/_ref/libref_siblings.so missing (build it where /root/reference exists)")
        lib = C.CDLL(_SIB_SO)
        for name in ("ref_sh_search", "ref_sbom_search"):
            f = getattr(lib, name)
            f.restype = C.c_ulonglong
            f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        lib.ref_sh_tables.restype = C.c_uint
        lib.ref_sh_tables.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, _u32p, _u32p]
        lib.ref_sbom_tables.restype = C.c_uint
        lib.ref_sbom_tables.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, _u32p]
        _sib = lib
    return _sib


def ref_sh(patterns, alphabet: int, text=None, want_tables: bool = False):
    """Reference preproc_sh (+ search_sh with the classic bad-character table).  Returns dict."""
    pats = _as_u8(patterns)
    p, m = pats.shape
    out = {}
    if text is not None:
        txt = _as_u8(text)
        out["count"] = int(siblings().ref_sh_search(_ptr(pats, _u8p), m, p, alphabet, _ptr(txt, _u8p), len(txt)))
    if want_tables:
        ns = m * p + 1
        tr, fin, nd = np.empty(ns * alphabet, np.int32), np.empty(ns, np.uint32), np.zeros(1, np.uint32)
        out["n_states"] = int(siblings().ref_sh_tables(_ptr(pats, _u8p), m, p, alphabet, _ptr(tr, _i32p), _ptr(fin, _u32p),
                                                       _ptr(nd, _u32p)))
        out.update(state_transition=tr, state_final=fin, n_distinct=int(nd[0]))
    return out


def ref_sbom(patterns, alphabet: int, text=None, want_tables: bool = False):
    """Reference preproc_sbom (+ search_sbom).  Returns dict."""
    pats = _as_u8(patterns)
    p, m = pats.shape
    out = {}
    if text is not None:
        txt = _as_u8(text)
        out["count"] = int(siblings().ref_sbom_search(_ptr(pats, _u8p), m, p, alphabet, _ptr(txt, _u8p), len(txt)))
    if want_tables:
        ns = m * p + 1
        tr, fm = np.empty(ns * alphabet, np.int32), np.empty(ns * 200, np.uint32)
        out["n_states"] = int(siblings().ref_sbom_tables(_ptr(pats, _u8p), m, p, alphabet, _ptr(tr, _i32p), _ptr(fm, _u32p)))
        out.update(state_transition=tr, state_final_multi=fm)
    return out


# ------------------------------------------------------------------ reference
def ref_ac(patterns, alphabet: int, text, threads: int = 1, want_tables: bool = False):
    """Reference preproc_ac + search_ac.  patterns: (p, m) uint8.  Returns dict."""
    pats = _as_u8(patterns)
    p, m = pats.shape
    txt = _as_u8(text)
    times = np.zeros(2, np.float64)
    meta = np.zeros(2, np.uint32)
    tr = sup = fin = None
    if want_tables:
        ns = m * p + 1
        tr = np.empty(ns * alphabet, np.int32)
        sup = np.empty(ns, np.uint32)
        fin = np.empty(ns, np.uint32)
    cnt = ref().ref_ac_search(_ptr(pats, _u8p), m, p, alphabet, _ptr(txt, _u8p), txt.size, threads,
                              _ptr(tr, _i32p), _ptr(sup, _u32p), _ptr(fin, _u32p), _ptr(meta, _u32p),
                              _ptr(times, _f64p))
    return dict(count=int(cnt), n_states=int(meta[0]), n_distinct=int(meta[1]), preproc_s=float(times[0]),
                search_s=float(times[1]), state_transition=tr, state_supply=sup, state_final=fin)


def ref_wu(patterns, alphabet: int, text, threads: int = 1, flat: bool = False, want_tables: bool = False):
    """Reference preproc_wu(2) + search_wu(2) with B = 3 (main.c:335).  Returns dict."""
    pats = _as_u8(patterns)
    p, m = pats.shape
    txt = _as_u8(text)
    times = np.zeros(2, np.float64)
    ss = int(ref().ref_wu_shiftsize(alphabet))
    if ss == 0:
        raise ValueError("alphabet not supported by the reference Wu-Manber (wu/wu.c:18-47)")
    SH = PV = PI = PS = None
    if want_tables:
        SH = np.empty(ss, np.int32)
        PS = np.empty(ss, np.int32)
        PV = np.full(ss * p, -7, np.int32)
        PI = np.full(ss * p, -7, np.int32)
    cnt = ref().ref_wu_search(_ptr(pats, _u8p), m, p, alphabet, int(flat), _ptr(txt, _u8p), txt.size, threads,
                              _ptr(SH, _i32p), _ptr(PV, _i32p), _ptr(PI, _i32p), _ptr(PS, _i32p),
                              _ptr(times, _f64p))
    return dict(count=int(cnt), shiftsize=ss, preproc_s=float(times[0]), search_s=float(times[1]),
                SHIFT=SH, PREFIX_value=PV, PREFIX_index=PI, PREFIX_size=PS)


# ----------------------------------------------------------------------- port
def port_ac(patterns, alphabet: int, text, want_tables: bool = False, cap: int | None = None):
    pats = _as_u8(patterns)
    p, m = pats.shape
    txt = _as_u8(text)
    lib = port()
    h = lib.oracle_ac_build(_ptr(pats, _u8p), m, p, alphabet)
    try:
        cap = int(cap if cap is not None else max(1, txt.size))
        pos = np.empty(cap, np.uint64)
        cnt = int(lib.oracle_ac_scan(h, _ptr(txt, _u8p), txt.size, _ptr(pos, _u64p), cap))
        out = dict(count=cnt, positions=pos[:min(cnt, cap)].copy(), n_states=int(lib.oracle_ac_states(h)),
                   n_distinct=int(lib.oracle_ac_distinct(h)))
        if want_tables:
            ns = m * p + 1
            tr = np.empty(ns * alphabet, np.int32)
            sup = np.empty(ns, np.uint32)
            fin = np.empty(ns, np.uint32)
            lib.oracle_ac_export_ref_tables(h, p, _ptr(tr, _i32p), _ptr(sup, _u32p), _ptr(fin, _u32p))
            out.update(state_transition=tr, state_supply=sup, state_final=fin)
        return out
    finally:
        lib.oracle_ac_free(h)


def port_wu(patterns, alphabet: int, text, want_tables: bool = False, cap: int | None = None):
    pats = _as_u8(patterns)
    p, m = pats.shape
    txt = _as_u8(text)
    lib = port()
    h = lib.oracle_wu_build(_ptr(pats, _u8p), m, p, alphabet)
    if not h:
        raise ValueError("unsupported alphabet (wu/wu.c:18-47) or m < 3")
    try:
        cap = int(cap if cap is not None else max(1, txt.size))
        pos = np.empty(cap, np.uint64)
        cnt = int(lib.oracle_wu_scan(h, _ptr(txt, _u8p), txt.size, _ptr(pos, _u64p), cap))
        out = dict(count=cnt, positions=pos[:min(cnt, cap)].copy())
        if want_tables:
            ss = int(lib.oracle_wu_shiftsize(alphabet))
            SH = np.empty(ss, np.int32)
            PS = np.empty(ss, np.int32)
            PV = np.full(ss * p, -7, np.int32)
            PI = np.full(ss * p, -7, np.int32)
            lib.oracle_wu_export_ref_tables(h, _ptr(SH, _i32p), _ptr(PV, _i32p), _ptr(PI, _i32p), _ptr(PS, _i32p))
            out.update(shiftsize=ss, SHIFT=SH, PREFIX_value=PV, PREFIX_index=PI, PREFIX_size=PS)
        return out
    finally:
        lib.oracle_wu_free(h)


def _flatten(patterns):
    """list of uint8 arrays (any lengths) or (p, m) array -> flat, offsets, lens."""
    if isinstance(patterns, np.ndarray) and patterns.ndim == 2:
        p, m = patterns.shape
        flat = _as_u8(patterns).reshape(-1)
        lens = np.full(p, m, np.uint32)
    else:
        arrs = [_as_u8(x).reshape(-1) for x in patterns]
        lens = np.array([a.size for a in arrs], np.uint32)
        flat = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    offsets = np.zeros(len(lens), np.uint64)
    if len(lens):
        offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    if flat.size == 0:
        flat = np.zeros(1, np.uint8)
    return flat, offsets, lens


def naive(patterns, text, cap: int | None = None):
    """Brute-force window membership; positions ascending (ties: ascending length)."""
    flat, offsets, lens = _flatten(patterns)
    txt = _as_u8(text)
    cap = int(cap if cap is not None else max(1, txt.size * max(1, len(set(lens.tolist())))))
    pos = np.empty(cap, np.uint64)
    cnt = int(port().oracle_naive_search(_ptr(flat, _u8p), _ptr(offsets, _u64p), _ptr(lens, _u32p), len(lens),
                                         _ptr(txt, _u8p), txt.size, _ptr(pos, _u64p), cap))
    return dict(count=cnt, positions=pos[:min(cnt, cap)].copy())


def set_search(patterns, text, cap: int | None = None, want_positions: bool = True):
    """Rolling-hash window membership (linear time); positions ascending."""
    flat, offsets, lens = _flatten(patterns)
    txt = _as_u8(text)
    if want_positions:
        cap = int(cap if cap is not None else max(1, txt.size * max(1, len(set(lens.tolist())))))
        pos = np.empty(cap, np.uint64)
    else:
        cap, pos = 0, None
    cnt = int(port().oracle_set_search(_ptr(flat, _u8p), _ptr(offsets, _u64p), _ptr(lens, _u32p), len(lens),
                                       _ptr(txt, _u8p), txt.size, _ptr(pos, _u64p), cap, 1))
    return dict(count=cnt, positions=(pos[:min(cnt, cap)].copy() if pos is not None else None))
