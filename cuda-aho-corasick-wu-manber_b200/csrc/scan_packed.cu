// Front ends of the 2-bit path (alphabet <= 4: DNA), hand-written for sm_100a.
//
// The tile sits RAW in shared memory (scan_kernel.cuh).  A lane reads its 112-symbol
// chunk plus 16 symbols of history with 8 conflict-free LDS.128 and packs them in
// registers, 16 symbols -> one 32-bit word (4 IMAD + 3 PRMT), so the walk below runs on
// 8 registers per lane:
//   AC<K>  dense DFA with the failure function folded in, K symbols per shared-memory
//          lookup (uint16 entry = byte offset of the next row | K hit bits << 1, bit 0 clear: the
//          hit bits sit where the next lookup's symbols go, so the address of the next lookup is ONE
//          bitwise select of entry and text).  Supersedes the one-symbol goto/failure walk of
//          cuda/cuda_ac.cu:88-95.
//   WM<S>  Wu-Manber block filter sampled every S symbols (SHIFT[block] < S as a
//          bitmap).  Supersedes the divergent skip loop of cuda/cuda_wm.cu:136-176.
// Once the chunk is in registers the warp refills its raw slot (TMA) and walks; one slot
// per warp therefore overlaps load and scan.  Hits of an exact AC automaton are matches;
// hits of a depth-truncated automaton and WM candidates are checked by the lane that found
// them (offset mask -> stage-2 suffix bitmap -> buckets) on 16-symbol windows read from a
// 2-bit copy of the tile the lanes leave in shared memory (7 STS per lane).
#include <cstring>

#include "scan_kernel.cuh"

namespace acwm {

// ------------------------------------------------------------ pack: 16 symbols -> 32 bits
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return w * 0x01041040u; } // result in the top byte

__device__ __forceinline__ uint32_t pack16(const uint4 r, uint32_t &badacc) {
	badacc |= (r.x | r.y) | (r.z | r.w);
	const uint32_t a = __byte_perm(pack4(r.x), pack4(r.y), 0x7373);
	const uint32_t b = __byte_perm(pack4(r.z), pack4(r.w), 0x7373);
	return __byte_perm(a, b, 0x5410);
}

// W[0] = the 16 symbols in front of the chunk, W[1..7] = the chunk; optionally mirrored
// into the warp's 2-bit copy pk (pk[0] = history of the tile, pk[1 + q] = tile symbols 16q..16q+15).
// `buf` = the warp's slot (shared address): raw [64 B history][3584 B tile], or -- text packed by the host
// (packed_in) -- [16 B history][896 B tile] in the very layout of W / pk (nothing to pack: 8 LDS.32 per lane).
template <bool STORE>
__device__ __forceinline__ void load_pack(const ScanArgs &a, uint32_t buf, uint32_t (&W)[8], uint32_t pk, uint32_t &badacc) {
	const uint32_t lane = lane_id();
	if (a.packed_in) {
		const uint32_t w = buf + 12 + 28 * lane; // word stride 7: conflict-free
#pragma unroll
		for (int k = 0; k < 8; k++)
			W[k] = lds32(w + 4 * k);
	} else {
		const uint32_t c = buf + kHalo - 16 + lane * kLane;
#pragma unroll
		for (int k = 0; k < 8; k++)
			W[k] = pack16(lds128(c + 16 * k), badacc);
	}
	if (STORE) {
		if (lane == 0)
			sts32(pk, W[0]);
		const uint32_t q = pk + 4 + 28 * lane; // word stride 7 (odd): conflict-free
#pragma unroll
		for (int k = 0; k < 7; k++)
			sts32(q + 4 * k, W[1 + k]);
	}
}

// 32-bit window of the 16 symbols ending at tile symbol `pos` (>= 0), from the 2-bit copy.
__device__ __forceinline__ uint32_t window16(uint32_t pk, uint32_t pos) {
	const uint32_t bit = 2 * pos + 2;
	const uint32_t wa = pk + ((bit >> 3) & ~3u);
	return __funnelshift_r(lds32(wa), lds32(wa + 4), bit);
}

struct PackedKey {
	static constexpr bool kPacked = true;
	static __device__ __forceinline__ uint32_t key_at(const ScanArgs &a, uint32_t, uint32_t pk, uint32_t pos) {
		return window16(pk, pos) >> (32 - 2 * a.prm.b2);
	}
};

// ------------------------------------------------------------ front end: AC, K symbols per lookup
// The lane's strides are laid out so that they END at the chunk end: in-chunk strides
// start kOff symbols in front of the chunk (kOff = 2 for K = 3, since 112 = 3*37 + 1),
// preceded by the warm-up strides that bring the state up to date.
// GLOBAL: the automaton lives in global memory (L2-resident), uint32 entries, K <= 2.
// ILP = 2: the lane walks its chunk as TWO independent chains (first and second half of the strides, the second
// with its own warm-up over the symbols in front of it).  A walk is a chain of dependent shared-memory lookups,
// ~35 cycles each; two chains per lane keep twice as many lookups in flight per warp.
//
// The K = 3 automaton in shared memory (the table of every small pattern set) runs a stride in FOUR instructions
// around its lookup: SHF (the stride's symbols to the address bits), LOP3 (address = bitwise select of entry and
// text), LDS.U16, SHF (a funnel shift drops the entry's low four bits -- a clear bit and its three hit bits --
// into the chain's hit word from the top): eight strides per hit word, in text order.
template <int K, bool EXACT, bool GLOBAL = false, int ILP = 1>
struct FrontAC : PackedKey {
	static constexpr int kSS = GLOBAL ? 2 : 1;                  // log2(entry bytes): symbols sit above it in the address
	static constexpr int kOff = (K - (int) kLane % K) % K;      // 2 / 0 / 0
	static constexpr int kStrides = ((int) kLane + kOff) / K;   // 38 / 56 / 112
	static constexpr bool kFunnel = K == 3 && !GLOBAL;          // hit words filled by funnel shifts (see above)
	static constexpr int kBits = kSS + K;                       // funnel: bits a stride adds to its hit word (4)
	static constexpr int kGroup = kFunnel ? 32 / kBits : (32 - kSS) / K; // strides per hit word: 8 | 10 / 15 / 31 (global: 15 / 30)
	static constexpr int kGroupSyms = kGroup * K;
	static constexpr int kLen0 = ILP == 2 ? kStrides / 2 : kStrides;    // strides of chain A (chain B: the rest)
	static constexpr int kWords0 = (kLen0 + kGroup - 1) / kGroup;
	static constexpr int kWords = kFunnel ? kWords0 + (kStrides - kLen0 + kGroup - 1) / kGroup : (kStrides + kGroup - 1) / kGroup;
	static constexpr uint32_t kSymMask2 = ((1u << (2 * K)) - 1) << kSS;
	static constexpr uint32_t kHitMask = ((1u << K) - 1) << kSS; // in the entry, above its kSS clear bits
	static_assert(kFunnel || K * kGroup + kSS <= 32, "the hit word keeps the entry's kSS low bits clear");
	static_assert(ILP == 1 || kFunnel, "two chains: K = 3 in shared memory only");
	static constexpr int kExpand = 1; // probes per candidate

	uint32_t tab_s;     // shared-memory DFA (shared address)
	const uint8_t *tab; // global-memory DFA
	uint32_t ent;       // current entry (byte offset of the row | hits)
	uint32_t nwu, hist; // warm-up strides / symbols of history they (and the first in-chunk stride) cover
	uint32_t nwb;       // warm-up strides of the second chain
	uint32_t W[8];      // W[0] = the 16 symbols in front of the chunk, W[1..7] = the chunk
	uint32_t H[3];      // long warm-up only: symbols -64..-17
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(uint32_t table_s, const TabRef &, const ScanArgs &a) {
		tab_s = table_s;
		tab = a.front;
		// the state must have seen depth-1 symbols of history when the chunk starts; the first
		// in-chunk stride covers kOff of them
		const uint32_t need = a.prm.depth - 1;
		nwu = need > (uint32_t) kOff ? (need - kOff + K - 1) / K : 0u;
		nwb = (need + K - 1) / K; // chain B starts on a stride boundary: whole strides of history (the host picks two chains for depth <= 13: 4 strides)
		hist = kOff + K * nwu; // <= 16: W[0] is enough; else up to 64 symbols of raw history
	}
	// bit b of hit word g = a hit at this chunk symbol (bits of symbols in front of the chunk are cleared by walk())
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) {
		if constexpr (kFunnel) {
			// word g holds strides I0 .. I0 + 7 of its chain, four bits each: [clear | hit at symbol 0, 1, 2 of the stride]
			const int i0 = g < kWords0 ? g * kGroup : kLen0 + (g - kWords0) * kGroup;
			return (uint32_t) (K * i0 - kOff - 1 + b - (b >> 2));
		} else
			return (uint32_t) (g * kGroupSyms + b - kSS - kOff);
	}
	__device__ __forceinline__ uint32_t probe_mask(const ScanArgs &, uint32_t, uint32_t, uint32_t) const {
		return 1u; // a hit of the truncated automaton is probed where it ends
	}

	// x = text bits with the stride's K symbols at bits [kSS, kSS + 2K) (anything elsewhere).  The entry keeps its row
	// offset above those bits, its hit bits inside them and zeros below: the next address is a bitwise select
	// (one LOP3 between two dependent lookups).
	__device__ __forceinline__ void step_of(uint32_t &e, uint32_t x) const {
		uint32_t addr;
		asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(addr) : "r"(e), "r"(x), "n"(kSymMask2)); // mask ? x : e
		if (GLOBAL)
			e = __ldg(reinterpret_cast<const uint32_t *>(tab + addr));
		else
			e = lds_u16(tab_s + addr);
	}
	// text bits of in-chunk stride i with its symbols at [kSS, kSS + 2K) (not masked)
	template <int I>
	__device__ __forceinline__ uint32_t sym_at() const {
		constexpr int bit = 32 - 2 * kOff + 2 * K * I; // W[0] holds stream bits 0..31, chunk symbol c sits at bit 32 + 2c
		constexpr int wi = bit >> 5, sh = bit & 31;
		if constexpr (sh + 2 * K <= 32)
			return sh >= kSS ? (W[wi] >> (sh >= kSS ? sh - kSS : 0)) : (W[wi] << (sh < kSS ? kSS - sh : 0));
		else // only K = 3 straddles words, and K = 3 tables are never global
			return __funnelshift_r(W[wi], W[wi + 1], sh - 1);
	}
	template <int I>
	__device__ __forceinline__ void stride(uint32_t &e) {
		step_of(e, sym_at<I>());
		if constexpr (kFunnel) {
			constexpr int g = I < kLen0 ? I / kGroup : kWords0 + (I - kLen0) / kGroup;
			hw[g] = __funnelshift_r(hw[g], e, kBits); // the entry's low bits [clear | 3 hits] enter from the top
		} else {
			uint32_t h = e & kHitMask;
			if (I == 0 && kOff)
				h &= ~(((1u << kOff) - 1) << kSS); // symbols in front of the chunk belong to the previous lane
			hw[I / kGroup] += h << (K * (I % kGroup));
		}
	}
	template <int J, int HALF>
	__device__ __forceinline__ void two_chains(uint32_t &ea, uint32_t &eb) {
		if constexpr (J < kStrides - HALF) {
			if constexpr (J < HALF)
				stride<J>(ea);
			stride<HALF + J>(eb);
			two_chains<J + 1, HALF>(ea, eb);
		}
	}
	template <int I>
	__device__ __forceinline__ void one_chain() {
		if constexpr (I < kStrides) {
			stride<I>(ent);
			one_chain<I + 1>();
		}
	}

	__device__ __forceinline__ void load(const ScanArgs &a, uint32_t buf, uint32_t pk, uint32_t &badacc) {
		load_pack<!EXACT>(a, buf, W, pk, badacc);
		if (hist > 16) {
			if (a.packed_in) {
				const uint32_t w = buf + 28 * lane_id();
				H[0] = lds32(w);
				H[1] = lds32(w + 4);
				H[2] = lds32(w + 8);
			} else {
				const uint32_t c = buf + kHalo + lane_id() * kLane;
				uint32_t dummy = 0;
				H[0] = pack16(lds128(c - 64), dummy);
				H[1] = pack16(lds128(c - 48), dummy);
				H[2] = pack16(lds128(c - 32), dummy);
			}
		}
	}

	__device__ __forceinline__ void walk(const ScanArgs &) {
		ent = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
		if constexpr (ILP == 2) {
			// both chains warm up over the same number of strides (the symbols in front of their first stride:
			// chain A's in W[0], chain B's inside the chunk), one loop, two lookups in flight
			constexpr int bitB = 32 - 2 * kOff + 2 * K * kLen0, wb = bitB >> 5, sb = bitB & 31;
			static_assert(wb >= 1, "chain B's history lies in the chunk");
			uint32_t ha = W[0] << (2 * kOff);                                           // stream bits [-2 kOff, 32 - 2 kOff)
			uint32_t hb = sb ? __funnelshift_r(W[wb - 1], W[wb], sb) : W[wb - 1];       // stream bits [bitB - 32, bitB)
			// the warm-up symbols, oldest first, at bit kSS; chain A's first in-chunk stride already covers kOff symbols of
			// history, so it may need one warm-up stride less than chain B and sits out the first round(s) of the loop
			ha >>= 32 - 2 * K * nwu - kSS;
			hb >>= 32 - 2 * K * nwb - kSS;
			uint32_t eb = 0;
#pragma unroll 1
			for (uint32_t i = 0; i < nwb; i++) {
				if (i + nwu >= nwb) {
					step_of(ent, ha);
					ha >>= 2 * K;
				}
				step_of(eb, hb);
				hb >>= 2 * K;
			}
			two_chains<0, kLen0>(ent, eb);
		} else {
			if (nwu) {
				if (hist <= 16) {
					uint32_t h = W[0] >> (32 - 2 * hist);
#pragma unroll 1
					for (uint32_t i = 0; i < nwu; i++) {
						step_of(ent, h << kSS);
						h >>= 2 * K;
					}
				} else {
					uint64_t lo = ((uint64_t) H[1] << 32) | H[0], hi = ((uint64_t) W[0] << 32) | H[2];
					const uint32_t sh = 128 - 2 * hist; // bits to drop from the front
					if (sh >= 64) {
						lo = hi >> (sh - 64);
						hi = 0;
					} else if (sh) {
						lo = (lo >> sh) | (hi << (64 - sh));
						hi >>= sh;
					}
#pragma unroll 1
					for (uint32_t i = 0; i < nwu; i++) {
						step_of(ent, (uint32_t) lo << kSS);
						lo = (lo >> (2 * K)) | (hi << (64 - 2 * K));
						hi >>= 2 * K;
					}
				}
			}
			one_chain<0>();
		}
		if constexpr (kFunnel) {
			// a chain's last word holds fewer than eight strides: bring them down to bit 0; then drop the hits at the
			// kOff symbols in front of the chunk (they belong to the previous lane)
			constexpr int r0 = kLen0 % kGroup, r1 = (kStrides - kLen0) % kGroup;
			if constexpr (r0 != 0)
				hw[kWords0 - 1] >>= kBits * (kGroup - r0);
			if constexpr (ILP == 2 && r1 != 0)
				hw[kWords - 1] >>= kBits * (kGroup - r1);
			if constexpr (kOff != 0)
				hw[0] &= ~(((1u << kOff) - 1) << kSS);
		}
	}

	__device__ __forceinline__ uint32_t count() const {
		uint32_t c = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			c += __popc(hw[g]);
		return c;
	}
	// keep only chunk-relative symbols in [lo_sym, hi_sym) (first / last tiles of the text only)
	__device__ __forceinline__ void mask_range(uint32_t lo_sym, uint32_t hi_sym) {
#pragma unroll
		for (int g = 0; g < kWords; g++) {
			if constexpr (kFunnel) {
				uint32_t keep = 0;
#pragma unroll 1
				for (uint32_t w = hw[g]; w; w &= w - 1) {
					const int b = __ffs(w) - 1;
					const uint32_t s = sym_of(g, b);
					if (s >= lo_sym && s < hi_sym)
						keep |= 1u << b;
				}
				hw[g] = keep;
			} else {
				const int base = g * kGroupSyms - kOff - kSS; // bit b of word g is symbol g*kGroupSyms + b - kSS - kOff
				const int lo = max((int) lo_sym - base, 0), hi = min((int) hi_sym - base, 32);
				uint32_t keep = 0;
				if (hi > lo)
					keep = (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1)) << lo;
				hw[g] &= keep;
			}
		}
	}
};

// ------------------------------------------------------------ front end: WM, sampled every S symbols
// MODE 0: direct-indexed bitmap in shared memory, 1: hashed bitmap in shared memory,
//      2: bitmap in global memory (L2-resident; direct index = multiplier 1, shift 0)
// A sample costs seven instructions around its lookup: the 16 symbols ending at it (funnel shift), the block's
// index, the word's address (shift, mask), LDS, the bit (shift by the index, which wraps at 32), and a funnel
// shift that drops the bit into the hit word from the top.
template <int S, int MODE>
struct FrontWM : PackedKey {
	static constexpr int kSamples = (int) kLane / S; // 112 / 56 / 28 / 14 / 7
	static constexpr int kWords = (kSamples + 31) / 32;
	static constexpr int kExpand = S;

	uint32_t bm_s;      // shared-memory block bitmap (shared address)
	const uint32_t *bm; // global-memory block bitmap
	TabRef rmk;         // offset masks
	uint32_t sh1, mult, sh2, sh2b;
	uint32_t W[8];
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(uint32_t table_s, const TabRef &rmask, const ScanArgs &a) {
		bm_s = table_s;
		bm = reinterpret_cast<const uint32_t *>(a.front);
		rmk = rmask;
		sh1 = a.prm.f1_sh1;
		mult = a.prm.f1_mult;
		sh2 = a.prm.f1_sh2;
		sh2b = a.prm.f1_k == 2 ? sh2 - 5 : sh2; // hashed bitmaps: where the entry's second bit comes from (one bit: the same again)
	}
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) ((g * 32 + b) * S); }

	// offsets r < S at which some pattern holds the block ending at tile symbol `pos`
	__device__ __forceinline__ uint32_t probe_mask(const ScanArgs &a, uint32_t, uint32_t pk, uint32_t pos) const {
		if (S == 1)
			return 1u;
		const uint32_t blk = window16(pk, pos) >> sh1;
		const uint32_t ri = (uint32_t) (blk * a.prm.r_mult) >> a.prm.r_sh;
		return S > 8 ? rmk.u16(ri) : rmk.u8(ri);
	}

	__device__ __forceinline__ void load(const ScanArgs &a, uint32_t buf, uint32_t pk, uint32_t &badacc) {
		load_pack<true>(a, buf, W, pk, badacc);
	}

	__device__ __forceinline__ void walk(const ScanArgs &) {
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
#pragma unroll
		for (int j = 0; j < kSamples; j++) {
			// 16 symbols ending at chunk symbol c = j*S; W[0] holds chunk symbols -16..-1
			const int bit = 2 * (j * S) + 2;
			const int wi = bit >> 5, sh = bit & 31;
			const uint32_t v = sh ? __funnelshift_r(W[wi], W[wi + 1], sh) : W[wi];
			uint32_t idx = v >> sh1, h = 0;
			if (MODE != 0) {
				h = idx * mult;
				idx = h >> sh2;
			}
			const uint32_t word = MODE == 2 ? __ldg(bm + (idx >> 5)) : lds32(bm_s + ((idx >> 3) & ~3u));
			uint32_t t = word >> (idx & 31);
			if (MODE != 0) // blocked Bloom filter: the entry's second bit sits in the same word
				t &= word >> ((h >> sh2b) & 31);
			hw[j / 32] = __funnelshift_r(hw[j / 32], t, 1); // the sample's bit enters from the top
		}
		constexpr int r = kSamples % 32; // the last word holds fewer than 32 samples: bring them down to bit 0
		if constexpr (r != 0)
			hw[kWords - 1] >>= 32 - r;
	}

	__device__ __forceinline__ uint32_t count() const {
		uint32_t c = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			c += __popc(hw[g]);
		return c;
	}
	__device__ __forceinline__ void mask_range(uint32_t, uint32_t) {} // probes are range-checked in verify
};

// ------------------------------------------------------------ dispatch
cudaError_t launch_scan_packed(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
#ifdef ACWM_DEV_ONLY // scripts/sass_dev.sh: only the kernels of BASELINE configs[0] / [1], one launch shape, for a quick look at their SASS
	if (a.prm.algo == ACWM_ALGO_AC)
		return launch_shape<FrontAC<3, true, false, 2>, true, 384, 2>(a, smem, grid, st);
	return launch_shape<FrontWM<8, 0>, false, 384, 2>(a, smem, grid, st);
#else
	const acwm_scan_params &p = a.prm;
	if (p.algo == ACWM_ALGO_AC && !p.front_kind) {
		const bool ex = p.exact_front != 0;
		if (!a.front_in_smem) {
			if (p.stride == 2)
				return ex ? launch_front<FrontAC<2, true, true>, true>(a, threads, smem, grid, st)
						  : launch_front<FrontAC<2, false, true>, false>(a, threads, smem, grid, st);
			if (p.stride == 1)
				return ex ? launch_front<FrontAC<1, true, true>, true>(a, threads, smem, grid, st)
						  : launch_front<FrontAC<1, false, true>, false>(a, threads, smem, grid, st);
			return cudaErrorInvalidValue;
		}
		if (p.stride == 3 && ex && p.ilp == 2) // the automaton of a small set (BASELINE configs[0]): two chains per lane
			return launch_front<FrontAC<3, true, false, 2>, true>(a, threads, smem, grid, st);
		switch (p.stride) {
		case 3: return ex ? launch_front<FrontAC<3, true>, true>(a, threads, smem, grid, st)
						  : launch_front<FrontAC<3, false>, false>(a, threads, smem, grid, st);
		case 2: return ex ? launch_front<FrontAC<2, true>, true>(a, threads, smem, grid, st)
						  : launch_front<FrontAC<2, false>, false>(a, threads, smem, grid, st);
		case 1: return ex ? launch_front<FrontAC<1, true>, true>(a, threads, smem, grid, st)
						  : launch_front<FrontAC<1, false>, false>(a, threads, smem, grid, st);
		default: return cudaErrorInvalidValue;
		}
	}
	const int mode = !a.front_in_smem ? 2 : (p.f1_mult != 1 ? 1 : 0);
	switch (p.stride) {
#define ACWM_WM(S)                                                                        \
	case S: return mode == 2 ? launch_front<FrontWM<S, 2>, false>(a, threads, smem, grid, st) \
		 : mode == 1 ? launch_front<FrontWM<S, 1>, false>(a, threads, smem, grid, st)         \
					 : launch_front<FrontWM<S, 0>, false>(a, threads, smem, grid, st);
	ACWM_WM(16)
	ACWM_WM(8)
	ACWM_WM(4)
	ACWM_WM(2)
	ACWM_WM(1)
#undef ACWM_WM
	default: return cudaErrorInvalidValue;
	}
#endif
}

} // namespace acwm
