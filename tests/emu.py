"""TEST INFRASTRUCTURE ONLY: a numpy emulation of what the scan kernels do with the
compiled tables (csrc/scan_packed.cu, csrc/scan_bytes.cu), used on the GPU-less CI box
to check the table compiler (csrc/tables.cpp) and the chunk / warm-up / candidate logic
against the oracle.  It reads the tables through ``acwm_table_blob`` and is never
imported by the product.
"""
from __future__ import annotations

import numpy as np

import acwm_pkg

acwm = acwm_pkg.load()

M32 = np.uint64(0xFFFFFFFF)


def _mul32(a, b):
    return (a.astype(np.uint64) * np.uint64(b)) & M32


def _mix64(v):
    lo = v & M32
    hi = v >> np.uint64(32)
    return (lo * np.uint64(0x9E3779B1) + hi * np.uint64(0x85EBCA77)) & M32


class Emulator:
    def __init__(self, mt):
        self.mt = mt
        self.p = mt.params()
        self.front = mt.blob(acwm.BLOB_FRONT)
        self.f2 = mt.blob(acwm.BLOB_FILTER2).view(np.uint32)
        self.bstart = mt.blob(acwm.BLOB_BUCKET_START).view(np.uint32)
        self.entries = mt.blob(acwm.BLOB_ENTRIES).view(acwm.VENTRY_DTYPE)
        self.pbytes = mt.blob(acwm.BLOB_PATTERNS)
        rm = mt.blob(acwm.BLOB_RMASK)
        self.rmask = rm.view(np.uint16 if self.p.r_entry_bytes == 2 else np.uint8) if rm.size else None
        self.info = mt.info
        self.vdfa = mt.blob(acwm.BLOB_VDFA).view(np.uint32) if self.p.verify_kind else None

    # ------------------------------------------------------------ windows
    def _win16(self, sym):
        """win[e] = 16 symbols ending at e, older symbol at lower bits (zeros before the text)."""
        n = sym.size
        pad = np.concatenate([np.zeros(15, np.uint64), sym.astype(np.uint64)])
        w = np.zeros(n, np.uint64)
        for i in range(16):
            w |= pad[i:i + n] << np.uint64(2 * i)
        return w

    def _win8(self, text):
        n = text.size
        pad = np.concatenate([np.zeros(7, np.uint64), text.astype(np.uint64)])
        w = np.zeros(n, np.uint64)
        for i in range(8):
            w |= pad[i:i + n] << np.uint64(8 * i)
        return w

    # ------------------------------------------------------------ verification
    def _verify(self, text, ends, keys):
        """ends/keys: candidate end positions and their stage-2 keys -> list of (e, mult)."""
        p = self.p
        n = text.size
        i2 = _mul32(keys, p.f2_mult) >> np.uint64(p.f2_sh)
        bits = (self.f2[(i2 >> np.uint64(5)).astype(np.int64)] >> (i2 & np.uint64(31)).astype(np.uint32)) & 1
        out = []
        if p.verify_kind:  # filtered AC: the window is walked through the full-depth automaton from its root (verify_dfa)
            m = p.m_min
            for e in ends[bits == 1].tolist():
                if e + 1 < m or e >= n:
                    continue
                ent = 0
                for c in text[e + 1 - m:e + 1].tolist():
                    ent = int(self.vdfa[(ent >> 1) * 4 + (c & 3)])
                if ent & 1:
                    out.append((e, 1))
            return out
        for e, key in zip(ends[bits == 1].tolist(), keys[bits == 1].tolist()):
            b = ((key * p.hb_mult) & 0xFFFFFFFF) >> p.hb_sh
            mult = 0
            for i in range(int(self.bstart[b]), int(self.bstart[b + 1])):
                en = self.entries[i]
                if int(en["key"]) != key:
                    continue
                ln = int(en["len"]) & 0x7FFFFFFF
                if e + 1 < ln or e >= n:
                    continue
                if int(en["len"]) >> 31:
                    mult += 1
                else:
                    off = int(en["offset"])
                    mult += bool(np.array_equal(text[e + 1 - ln:e + 1], self.pbytes[off:off + ln]))
            if mult:
                out.append((e, mult))
        return out

    # ------------------------------------------------------------ front ends
    def _ac_hits(self, sym, chunk, K_bits, cols_of, row_shift, hit_shift=0):
        """Generic chunked DFA walk; returns sorted hit positions (may include e >= n).
        Mirrors FrontAC::scan / FrontACB::scan: the in-chunk strides start `off` symbols in
        front of the chunk so that they end exactly at the chunk end, preceded by the
        warm-up strides that bring the state up to date (depth-1 symbols of history)."""
        p = self.p
        K = p.stride
        D = p.depth
        n = sym.size
        nchunks = (n + chunk - 1) // chunk
        off = (K - chunk % K) % K
        need = D - 1
        nwu = (need - off + K - 1) // K if need > off else 0
        hist = off + K * nwu
        ent_bytes = 2 if self.info["table_in_smem"] else 4
        tab = self.front.view(np.uint16 if ent_bytes == 2 else np.uint32).astype(np.int64)
        cols = cols_of
        total = nchunks * chunk
        ext = np.zeros(total + hist, np.int64)
        ext[hist:hist + n] = sym
        starts = np.arange(nchunks, dtype=np.int64) * chunk
        state = np.zeros(nchunks, np.int64)
        hits = []
        for t in range(-nwu, (chunk + off) // K):
            first = -off + K * t  # chunk-relative symbol of this stride's first symbol
            idx = np.zeros(nchunks, np.int64)
            for i in range(K):
                idx |= ext[starts + hist + first + i] << (K_bits * i)
            ent = tab[state * cols + idx]
            state = ent >> row_shift
            # packed DFA entries keep the K hit bits above log2(entry bytes) clear low bits; bytes path: bit 0
            h = (ent >> hit_shift) & ((1 << K) - 1)
            if t >= 0 and h.any():
                for i in range(K):
                    if first + i < 0:
                        continue  # belongs to the previous lane
                    sel = np.nonzero((h >> i) & 1)[0]
                    if sel.size:
                        hits.append(starts[sel] + first + i)
        if not hits:
            return np.zeros(0, np.int64)
        return np.sort(np.concatenate(hits))

    def _expand(self, cand, blocks, s, n):
        """Candidate sample positions -> end positions to probe: the offsets r < s whose bit is set in
        the offset mask of the candidate's block (probe_mask of FrontWM / FrontWMB)."""
        if s == 1 or self.rmask is None:
            ends = cand
        else:
            ri = (_mul32(blocks, self.p.r_mult) >> np.uint64(self.p.r_sh)).astype(np.int64)
            masks = self.rmask[ri].astype(np.int64)
            sel = (masks[:, None] >> np.arange(s)[None, :]) & 1
            ends = (cand[:, None] + np.arange(s)[None, :])[sel == 1]
        return ends[ends < n]

    def search(self, text):
        """-> (count, positions) exactly as the kernels would produce them."""
        p = self.p
        text = np.ascontiguousarray(text, np.uint8)
        n = text.size
        if n == 0:
            return 0, np.zeros(0, np.uint64)
        m_min = p.m_min
        if p.packed2bit:
            assert int(text.max()) < 4, "bad text for the 2-bit path"
            sym = text.astype(np.int64)
            win = self._win16(text)
            if p.algo == acwm.AC and not p.front_kind:
                # entry = byte offset of the next row | hits (uint16 in shared memory, uint32 in global memory)
                ss = 1 if self.info["table_in_smem"] else 2
                hits = self._ac_hits(sym, 112, 2, 1 << (2 * p.stride), 2 * p.stride + ss, ss)
                hits = hits[(hits >= m_min - 1) & (hits < n)]
                if p.exact_front:
                    return int(hits.size), hits.astype(np.uint64)
                ends = hits
            else:
                s = p.stride
                cpos = np.arange(0, ((n + 111) // 112) * 112, s, dtype=np.int64)
                cin = cpos[cpos < n]
                v = np.zeros(cpos.size, np.uint64)
                v[:cin.size] = win[cin]
                h = _mul32(v >> np.uint64(p.f1_sh1), p.f1_mult)
                idx = h >> np.uint64(p.f1_sh2)
                bm = self.front.view(np.uint32)
                word = bm[(idx >> np.uint64(5)).astype(np.int64)]
                bit = (word >> (idx & np.uint64(31)).astype(np.uint32)) & 1
                if p.f1_k == 2:  # blocked Bloom filter: the entry's second bit in the same word
                    bit &= (word >> ((h >> np.uint64(p.f1_sh2 - 5)) & np.uint64(31)).astype(np.uint32)) & 1
                ends = self._expand(cpos[bit == 1], (v >> np.uint64(p.f1_sh1))[bit == 1], s, n)
            keys = win[ends] >> np.uint64(32 - 2 * p.b2)
        else:
            win = self._win8(text)
            if p.algo == acwm.AC:
                alpha = min(p.alphabet, 255)
                cls = np.minimum(text.astype(np.int64), alpha)
                lognc = int(p.n_classes).bit_length() - 1
                hits = self._ac_hits(cls, 112, lognc, 1 << lognc, 1)
                hits = hits[(hits >= m_min - 1) & (hits < n)]
                if p.exact_front:
                    return int(hits.size), hits.astype(np.uint64)
                ends = hits
            else:
                s = p.stride
                cpos = np.arange(0, n, s, dtype=np.int64)
                blk = win[cpos] >> np.uint64(p.f1_sh1)
                h = _mul32(_mix64(blk), p.f1_mult)
                idx = h >> np.uint64(p.f1_sh2)
                bm = self.front.view(np.uint32)
                word = bm[(idx >> np.uint64(5)).astype(np.int64)]
                bit = (word >> (idx & np.uint64(31)).astype(np.uint32)) & 1
                if p.f1_k == 2:
                    bit &= (word >> ((h >> np.uint64(p.f1_sh2 - 5)) & np.uint64(31)).astype(np.uint32)) & 1
                ends = self._expand(cpos[bit == 1], _mix64(blk)[bit == 1], s, n)
            keys = _mix64(win[ends] >> np.uint64(64 - 8 * p.b2))
        res = self._verify(text, ends, keys)
        pos = []
        for e, mult in res:
            pos.extend([e] * mult)
        return len(pos), np.array(pos, np.uint64)
