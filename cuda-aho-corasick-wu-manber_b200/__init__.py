"""B200-native Aho-Corasick / Wu-Manber multi-pattern scan -- host-side Python layer.

Everything here is a thin ctypes binding over the C ABI of ``libacwm_b200.so``
(``include/acwm.h``); the scan itself runs only in the hand-written sm_100a kernels of
``csrc/``.  There is no CPU search path: if the shared library is missing, importing
this package raises.

    from acwm_pkg import load; acwm = load()
    mt = acwm.Matcher(acwm.AC, patterns, alphabet=4)          # preproc_ac equivalent
    count, positions = mt.search_host(text)                   # search_ac equivalent (+ positions)

Reference-shaped functions (same names / arguments as smatcher.h) are in
``.smatcher``; the multi-GPU sharding layer in ``.sharding``; synthetic data in
``.datagen``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacwm_b200.so")

AC, WM = 0, 1
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_OVERFLOW, ERR_BAD_TEXT = range(7)

BLOB_FRONT, BLOB_FILTER2, BLOB_BUCKET_START, BLOB_ENTRIES, BLOB_PATTERNS, BLOB_PARAMS, BLOB_SYMCLASS, BLOB_RMASK, BLOB_VDFA = range(9)


class AcwmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"acwm error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("smem_table_budget", C.c_uint32), ("force_stride", C.c_uint32), ("force_depth", C.c_uint32),
                ("force_bytes_path", C.c_uint32), ("force_threads", C.c_uint32), ("force_stages", C.c_uint32),
                ("force_f2_bits", C.c_uint32), ("force_r_bits", C.c_uint32), ("force_smem_tables", C.c_uint32),
                ("force_front", C.c_uint32), ("force_ctas", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("algo", "alphabet", "n_patterns", "n_distinct", "m_min", "m_max", "packed2bit", "stride", "depth",
                 "exact_front", "n_states", "n_rows", "table_in_smem", "smem_bytes")] + \
               [("table_bytes", C.c_uint64), ("threads", C.c_uint32), ("stages", C.c_uint32), ("ctas_per_sm", C.c_uint32),
                ("front_kind", C.c_uint32)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


class ScanParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("algo", "packed2bit", "alphabet", "m_min", "m_max", "stride", "depth", "exact_front", "n_rows",
                 "f1_sh1", "f1_mult", "f1_sh2", "f1_words", "b2", "f2_mult", "f2_sh", "f2_words", "hb_mult",
                 "hb_sh", "n_buckets", "n_entries", "n_classes", "r_mult", "r_sh", "r_entries", "r_entry_bytes", "r_in_smem", "f2_in_smem", "front_kind", "verify_kind", "v_rows", "ilp", "f1_k")]


VENTRY_DTYPE = np.dtype([("key", "<u4"), ("len", "<u4"), ("offset", "<u8")])

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.acwm_last_error.restype = C.c_char_p
        L.acwm_build.argtypes = [C.c_int, _u8p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Options),
                                 C.POINTER(C.c_void_p)]
        L.acwm_upload.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.acwm_scan_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.acwm_fetch.argtypes = [C.c_void_p, _u64p, C.c_void_p, C.c_uint64, _u64p, C.c_void_p]
        L.acwm_result_device_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.acwm_search_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, _u64p, C.c_void_p, C.c_uint64, _u64p]
        L.acwm_last_kernel_seconds.restype = C.c_double
        L.acwm_last_kernel_seconds.argtypes = [C.c_void_p]
        L.acwm_last_h2d_bytes.restype = C.c_uint64
        L.acwm_last_h2d_bytes.argtypes = [C.c_void_p]
        L.acwm_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.acwm_set_overlap.argtypes = [C.c_void_p, C.c_int]
        L.acwm_set_peers.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _u64p]
        L.acwm_fetch_global_count.argtypes = [C.c_void_p, _u64p, C.c_void_p]
        L.acwm_profiled_seconds.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.acwm_launch_count.restype = C.c_ulonglong
        L.acwm_launch_count.argtypes = [C.c_void_p]
        L.acwm_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.acwm_free.argtypes = [C.c_void_p]
        L.acwm_free.restype = None
        L.acwm_shard_bounds.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u64p]
        L.acwm_shard_bounds.restype = None
        L.acwm_table_blob.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _u64p]
        L.acwm_device_count.restype = C.c_int
        L.acwm_search_host_sharded.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p, C.c_uint64, _u64p,
                                               C.c_void_p, C.c_uint64, _u64p, _u64p]
        L.acwm_peers_create.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
        L.acwm_peers_destroy.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
        L.acwm_scan_device_sharded.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(C.c_void_p), _u64p, C.c_int]
        L.acwm_fetch_sharded.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, _u64p, _u64p]
        L.acwm_text_to_device.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.acwm_device_free.argtypes = [C.c_int, C.c_void_p]
        L.acwm_device_free.restype = None
        L.acwm_set_trace.argtypes = [C.c_void_p, C.c_void_p]
        L.acwm_pack_text_2bit.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int)]
        L.acwm_trace_words_per_cta.restype = C.c_uint32
        L.acwm_symbol_map.argtypes = [C.c_uint32, _u8p]
        L.acwm_encode_symbols.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, _u64p]
        L.acwm_load_text.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), _u64p]
        L.acwm_free_text.argtypes = [C.c_void_p]
        L.acwm_free_text.restype = None
        L.acwm_patterns_with_hits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                              C.c_uint32, C.c_void_p]
        L.acwm_select_data_file.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_char_p, C.c_char_p, C.c_char_p,
                                            C.c_size_t]
        _lib = L
    return _lib


def _check(rc: int, allow=()):
    if rc != OK and rc not in allow:
        raise AcwmError(rc, lib().acwm_last_error().decode())
    return rc


def flatten_patterns(patterns):
    """(p, m) uint8 array, or a list of uint8 arrays of any lengths -> (flat, lens|None, m, p)."""
    if isinstance(patterns, np.ndarray) and patterns.ndim == 2:
        pats = np.ascontiguousarray(patterns, dtype=np.uint8)
        return pats.reshape(-1), None, pats.shape[1], pats.shape[0]
    arrs = [np.ascontiguousarray(x, dtype=np.uint8).reshape(-1) for x in patterns]
    lens = np.array([a.size for a in arrs], dtype=np.uint32)
    flat = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    return flat, lens, 0, len(arrs)


def shard_bounds(n: int, world: int, rank: int, halo: int):
    """Shard geometry of main.c:467-477 (see acwm_shard_bounds)."""
    s, l = C.c_uint64(), C.c_uint64()
    lib().acwm_shard_bounds(n, world, rank, halo, C.byref(s), C.byref(l))
    return int(s.value), int(l.value)


def device_count() -> int:
    return int(lib().acwm_device_count())


def search_host_sharded(matchers, text, cap: int | None = None, want_positions: bool = True,
                        allow_overflow: bool = False):
    """The multi-rank flow of main.c:464-656 in one process (``acwm_search_host_sharded``): matchers[r] scans
    shard r of the host text on its own device and host thread; returns (count, global sorted positions,
    per-shard counts)."""
    if hasattr(text, "data_ptr"):
        ptr, n = text.data_ptr(), text.numel()
    else:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        ptr, n = text.ctypes.data, text.size
    world = len(matchers)
    hs = (C.c_void_p * world)(*[m._h for m in matchers])
    count, nw = C.c_uint64(), C.c_uint64()
    per = np.zeros(world, np.uint64)
    if want_positions:
        cap = int(cap if cap is not None else max(1, n))
        pos = np.empty(cap, np.uint64)
        pptr = pos.ctypes.data_as(C.c_void_p)
    else:
        cap, pos, pptr = 0, np.zeros(0, np.uint64), None
    _check(lib().acwm_search_host_sharded(hs, world, C.c_void_p(ptr), n, C.byref(count), pptr, cap, C.byref(nw),
                                          per.ctypes.data_as(_u64p)),
           allow=(ERR_OVERFLOW,) if allow_overflow else ())
    return int(count.value), pos[:int(nw.value)], per


def _handles(matchers):
    return (C.c_void_p * len(matchers))(*[m._h for m in matchers])


def peers_create(matchers):
    """Wire the in-kernel count exchange between uploaded matchers of ONE process (``acwm_peers_create``): mailboxes
    by cudaMalloc, peer access between their devices -- no torch, NCCL or MPI involved."""
    _check(lib().acwm_peers_create(_handles(matchers), len(matchers)))


def peers_destroy(matchers):
    _check(lib().acwm_peers_destroy(_handles(matchers), len(matchers)))


def scan_device_sharded(matchers, shards, want_positions: bool = True):
    """One scan of every device-resident shard (``acwm_scan_device_sharded``); shards[r] = CUDA uint8 tensor on
    matchers[r]'s device holding shard r with its halo, or a (device pointer, length) pair."""
    ptrs, lens = [], []
    for sh in shards:
        if hasattr(sh, "data_ptr"):
            ptrs.append(sh.data_ptr())
            lens.append(sh.numel())
        else:
            ptrs.append(int(sh[0]))
            lens.append(int(sh[1]))
    pa = (C.c_void_p * len(ptrs))(*ptrs)
    la = (C.c_uint64 * len(lens))(*lens)
    _check(lib().acwm_scan_device_sharded(_handles(matchers), len(matchers), pa, la, int(want_positions)))


def fetch_sharded(matchers):
    """-> (global count as exchanged inside the kernels, per-shard counts)."""
    g = C.c_uint64()
    per = np.zeros(len(matchers), np.uint64)
    _check(lib().acwm_fetch_sharded(_handles(matchers), len(matchers), C.byref(g), per.ctypes.data_as(_u64p)))
    return int(g.value), per


class Matcher:
    """A compiled pattern set (``acwm_matcher``).  algo = AC or WM."""

    def __init__(self, algo: int, patterns, alphabet: int, **opts):
        flat, lens, m, p = flatten_patterns(patterns)
        o = Options()
        for k, v in opts.items():
            setattr(o, k, int(v))
        h = C.c_void_p()
        keep = flat if flat.size else np.zeros(1, np.uint8)
        _check(lib().acwm_build(algo, keep.ctypes.data_as(_u8p),
                                None if lens is None else lens.ctypes.data_as(_u32p), m, p, alphabet,
                                C.byref(o), C.byref(h)))
        self._h = h
        self.algo = algo

    # ---- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            lib().acwm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> dict:
        i = Info()
        _check(lib().acwm_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def blob(self, which: int) -> np.ndarray:
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        _check(lib().acwm_table_blob(self._h, which, C.byref(ptr), C.byref(nbytes)))
        if nbytes.value == 0:
            return np.zeros(0, np.uint8)
        buf = (C.c_uint8 * nbytes.value).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def params(self) -> ScanParams:
        return ScanParams.from_buffer_copy(self.blob(BLOB_PARAMS).tobytes())

    # ---- device
    def upload(self, device: int = -1, pos_capacity: int = 0):
        _check(lib().acwm_upload(self._h, device, pos_capacity))
        return self

    def scan_device(self, d_text_ptr: int, n: int, want_positions: bool = True, stream: int = 0,
                    report_from: int = 0):
        """Asynchronous scan of device-resident text (raw pointer + length)."""
        _check(lib().acwm_scan_device(self._h, C.c_void_p(d_text_ptr), n, report_from, int(want_positions),
                                      C.c_void_p(stream)))

    def scan_tensor(self, text, want_positions: bool = True, report_from: int = 0):
        """Scan a CUDA uint8 torch tensor on torch's current stream."""
        import torch
        assert text.is_cuda and text.dtype == torch.uint8 and text.is_contiguous()
        st = torch.cuda.current_stream(text.device).cuda_stream
        self.scan_device(text.data_ptr(), text.numel(), want_positions, st, report_from)

    def fetch(self, cap: int = 0, stream: int = 0, allow_overflow: bool = False):
        count, nw = C.c_uint64(), C.c_uint64()
        pos = np.empty(max(cap, 1), np.uint64)
        rc = _check(lib().acwm_fetch(self._h, C.byref(count), pos.ctypes.data_as(C.c_void_p) if cap else None, cap,
                                     C.byref(nw), C.c_void_p(stream)),
                    allow=(ERR_OVERFLOW,) if allow_overflow else ())
        return int(count.value), pos[:int(nw.value)], rc

    def result_device_ptrs(self):
        c, p = C.c_void_p(), C.c_void_p()
        _check(lib().acwm_result_device_ptrs(self._h, C.byref(c), C.byref(p)))
        return c.value, p.value

    # ---- host text, end to end
    def search_host(self, text, cap: int | None = None, want_positions: bool = True, allow_overflow: bool = False,
                    out: np.ndarray | None = None):
        """text: numpy uint8 array (or anything exposing a host pointer via ``data_ptr``/``ctypes``).
        out: optional uint64 array the positions are written to (its size is the capacity)."""
        if hasattr(text, "data_ptr"):  # pinned torch tensor
            ptr, n = text.data_ptr(), text.numel()
        else:
            text = np.ascontiguousarray(text, dtype=np.uint8)
            ptr, n = text.ctypes.data, text.size
        count, nw = C.c_uint64(), C.c_uint64()
        if want_positions:
            if out is not None:
                assert out.dtype == np.uint64 and out.flags.c_contiguous
                pos, cap = out, int(out.size)
            else:
                cap = int(cap if cap is not None else max(1, n))
                pos = np.empty(cap, np.uint64)
            pptr = pos.ctypes.data_as(C.c_void_p)
        else:
            cap, pos, pptr = 0, np.zeros(0, np.uint64), None
        rc = _check(lib().acwm_search_host(self._h, C.c_void_p(ptr), n, C.byref(count), pptr, cap, C.byref(nw)),
                    allow=(ERR_OVERFLOW,) if allow_overflow else ())
        self.last_rc = rc
        return int(count.value), pos[:int(nw.value)]

    def set_overlap(self, on: bool):
        """Back-to-back device scans may overlap (programmatic dependent launch); see acwm_set_overlap."""
        _check(lib().acwm_set_overlap(self._h, int(on)))

    def set_peers(self, rank: int, world: int, mailbox_ptrs):
        """In-kernel count exchange over peer memory (see acwm_set_peers); None / world<=1 = off."""
        if not mailbox_ptrs or world <= 1:
            _check(lib().acwm_set_peers(self._h, 0, 0, None))
            return
        arr = (C.c_uint64 * world)(*[int(p) for p in mailbox_ptrs])
        _check(lib().acwm_set_peers(self._h, rank, world, arr))

    def fetch_global_count(self, stream: int = 0) -> int:
        g = C.c_uint64()
        _check(lib().acwm_fetch_global_count(self._h, C.byref(g), C.c_void_p(stream)))
        return int(g.value)

    def set_trace(self, d_trace_ptr: int | None):
        """Kernel timeline instrumentation (see acwm_set_trace); None = off."""
        _check(lib().acwm_set_trace(self._h, C.c_void_p(d_trace_ptr or 0)))

    def set_profiling(self, on: bool):
        _check(lib().acwm_set_profiling(self._h, int(on)))

    def profiled_seconds(self):
        """(scan kernel seconds, finalize kernels seconds) of the last profiled scan_device."""
        a, b = C.c_double(), C.c_double()
        _check(lib().acwm_profiled_seconds(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    @property
    def last_kernel_seconds(self) -> float:
        return float(lib().acwm_last_kernel_seconds(self._h))

    @property
    def last_h2d_bytes(self) -> int:
        """Bytes of text the last search_host sent over the link (n, or n/4 when the host packer ran)."""
        return int(lib().acwm_last_h2d_bytes(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().acwm_launch_count(self._h))


# ---------------------------------------------------------------- data files (main.c:31-123,453)
def symbol_map(alphabet: int) -> np.ndarray:
    m = np.zeros(256, np.uint8)
    _check(lib().acwm_symbol_map(alphabet, m.ctypes.data_as(_u8p)))
    return m


def encode_symbols(raw, alphabet: int) -> np.ndarray:
    """Corpus bytes -> symbol codes in [0, alphabet) (FASTA headers / non-symbols dropped)."""
    raw = np.ascontiguousarray(np.frombuffer(raw, np.uint8) if isinstance(raw, (bytes, bytearray)) else raw, np.uint8)
    out = np.empty(max(raw.size, 1), np.uint8)
    n = C.c_uint64()
    _check(lib().acwm_encode_symbols(C.c_void_p(raw.ctypes.data), raw.size, alphabet, C.c_void_p(out.ctypes.data),
                                     C.byref(n)))
    return out[:int(n.value)].copy()


def load_text(path: str, alphabet: int, max_symbols: int = 0) -> np.ndarray:
    ptr, n = C.c_void_p(), C.c_uint64()
    _check(lib().acwm_load_text(os.fsencode(path), alphabet, max_symbols, C.byref(ptr), C.byref(n)))
    try:
        buf = (C.c_uint8 * max(int(n.value), 1)).from_address(ptr.value)
        return np.frombuffer(buf, np.uint8)[:int(n.value)].copy()
    finally:
        lib().acwm_free_text(ptr)


def patterns_with_hits(text: np.ndarray, m: int, p: int, alphabet: int, seed: int, hit_percent: int = 50) -> np.ndarray:
    text = np.ascontiguousarray(text, np.uint8)
    out = np.empty((p, m), np.uint8)
    _check(lib().acwm_patterns_with_hits(C.c_void_p(text.ctypes.data), text.size, m, p, alphabet, seed, hit_percent,
                                         C.c_void_p(out.ctypes.data)))
    return out


def select_data_file(m: int, n: int, alphabet: int, data_root: str | None = None):
    """(pattern_path, text_path) of main.c:31-123; raises AcwmError for the reference's fail() cases."""
    pp, tp = C.create_string_buffer(4096), C.create_string_buffer(4096)
    _check(lib().acwm_select_data_file(m, n, alphabet, os.fsencode(data_root) if data_root else None, pp, tp, 4096))
    return os.fsdecode(pp.value), os.fsdecode(tp.value)


def pack_text_2bit(text: np.ndarray):
    """(packed uint8[ceil(n/4)], bad_text) -- the host-side packer of acwm_search_host (see acwm_pack_text_2bit)."""
    text = np.ascontiguousarray(text, np.uint8)
    out = np.empty((text.size + 3) // 4, np.uint8)
    bad = C.c_int()
    _check(lib().acwm_pack_text_2bit(C.c_void_p(text.ctypes.data), text.size, C.c_void_p(out.ctypes.data), C.byref(bad)))
    return out, bool(bad.value)
