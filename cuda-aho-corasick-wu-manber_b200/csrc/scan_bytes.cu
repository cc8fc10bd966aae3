// Front ends of the bytes path (alphabet > 4: one byte per symbol), sm_100a.
//
// Same skeleton as the 2-bit path (scan_kernel.cuh): every warp keeps a ring of RAW
// tiles of kTile = 3584 bytes (+64 bytes of history) in shared memory, each filled by ONE
// TMA bulk copy (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) issued by
// lane 0 one or two tiles ahead; lanes read their 112-byte chunk with 7 conflict-free LDS.128.
//
// Front ends:
//   ACB   dense DFA, failure function folded in, one class-compressed symbol per
//         lookup; table in shared memory (uint16) or, when it does not fit, in global
//         memory served from L2 (uint32; the host pins it with an access-policy window).
//   WMB<S> Wu-Manber block filter (up to 8 bytes, mixed to 32 bits) sampled every S bytes.
#include <cstring>

#include "scan_kernel.cuh"

namespace acwm {

// 64-bit window of the 8 bytes ending at buffer byte index bidx (>= 7); buf = shared address of the slot.
__device__ __forceinline__ uint64_t window8(uint32_t buf, uint32_t bidx) {
	const uint32_t b0 = bidx - 7;
	const uint32_t w = buf + (b0 & ~3u);
	const uint32_t sh = (b0 & 3) * 8;
	const uint32_t w0 = lds32(w), w1 = lds32(w + 4), w2 = lds32(w + 8);
	const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
	return ((uint64_t) hi << 32) | lo;
}

__device__ __forceinline__ uint32_t mix64(uint64_t v) {
	return (uint32_t) v * 0x9E3779B1u + (uint32_t) (v >> 32) * 0x85EBCA77u;
}

struct BytesKey {
	static constexpr bool kPacked = false;
	static __device__ __forceinline__ uint32_t key_at(const ScanArgs &a, uint32_t buf, uint32_t, uint32_t pos) {
		return mix64(window8(buf, kHalo + pos) >> (64 - 8 * a.prm.b2));
	}
};

// one lane's 112-byte chunk: 7 groups of 16 bytes, each handed to the front end together
// with the 16 bytes in front of it
#define ACWM_SCAN_GROUPS()                                         \
	uint32_t chunk;                                                \
	__device__ __forceinline__ void load(const ScanArgs &, uint32_t buf, uint32_t, uint32_t &) { chunk = buf + kHalo + lane_id() * kLane; } \
	__device__ __forceinline__ void walk(const ScanArgs &a) {      \
		begin(a, chunk);                                           \
		uint4 prev = lds128(chunk - 16);                           \
		uint4 c;                                                   \
		c = lds128(chunk + 0);  group<0>(prev, c); prev = c;       \
		c = lds128(chunk + 16); group<1>(prev, c); prev = c;       \
		c = lds128(chunk + 32); group<2>(prev, c); prev = c;       \
		c = lds128(chunk + 48); group<3>(prev, c); prev = c;       \
		c = lds128(chunk + 64); group<4>(prev, c); prev = c;       \
		c = lds128(chunk + 80); group<5>(prev, c); prev = c;       \
		c = lds128(chunk + 96); group<6>(prev, c);                 \
		finish_words();                                            \
	}

// ------------------------------------------------------------ front end: AC over bytes
template <bool IN_SMEM>
struct FrontACB : BytesKey {
	static constexpr int kWords = 4; // 112 hit bits
	static constexpr int kExpand = 1;
	uint32_t tab_s;     // shared-memory automaton (shared address)
	const uint8_t *tab; // global-memory automaton
	uint32_t ent, lognc, alpha;
	uint32_t hw[kWords];

	__device__ __forceinline__ uint32_t probe_mask(const ScanArgs &, uint32_t, uint32_t, uint32_t) const {
		return 1u;
	}
	__device__ __forceinline__ void init(uint32_t smem_tab, const TabRef &, const ScanArgs &a) {
		tab_s = smem_tab;
		tab = a.front;
		lognc = 31 - __clz(a.prm.n_classes);
		alpha = min(a.prm.alphabet, 255u);
	}
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) (g * 32 + b); }

	__device__ __forceinline__ uint32_t step(uint32_t byte) {
		const uint32_t cls = min(byte, alpha);
		const uint32_t idx = ((ent >> 1) << lognc) + cls;
		if (IN_SMEM)
			ent = lds_u16(tab_s + 2 * idx);
		else
			ent = __ldg(reinterpret_cast<const uint32_t *>(tab) + idx);
		return ent & 1u;
	}
	__device__ __forceinline__ void begin(const ScanArgs &a, uint32_t chunk) {
		ent = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
		const uint32_t nwu = a.prm.depth - 1;
#pragma unroll 1
		for (uint32_t i = 0; i < nwu; i++)
			(void) step(lds_u8(chunk + i - nwu));
	}
	// one 16-byte group: G = group index 0..6
	template <int G>
	__device__ __forceinline__ void group(const uint4 &, const uint4 &c) {
		const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const uint32_t h = step((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
			hw[(G * 16 + k) / 32] += h << ((G * 16 + k) % 32);
		}
	}
	__device__ __forceinline__ void finish_words() {}
	ACWM_SCAN_GROUPS()
	__device__ __forceinline__ uint32_t count() const {
		return __popc(hw[0]) + __popc(hw[1]) + __popc(hw[2]) + __popc(hw[3]);
	}
	__device__ __forceinline__ void mask_range(uint32_t lo_sym, uint32_t hi_sym) {
#pragma unroll
		for (int g = 0; g < kWords; g++) {
			const int lo = max((int) lo_sym - g * 32, 0), hi = min((int) hi_sym - g * 32, 32);
			uint32_t keep = 0;
			if (hi > lo)
				keep = (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1)) << lo;
			hw[g] &= keep;
		}
	}
};

// ------------------------------------------------------------ front end: WM over bytes
template <int S, bool GLOBAL>
struct FrontWMB : BytesKey {
	static constexpr int kSamples = (int) kLane / S; // 112 / 56 / 28 / 14 / 7
	static constexpr int kWords = (kSamples + 31) / 32;
	static constexpr int kExpand = S;
	uint32_t bm_s;      // shared-memory block bitmap (shared address)
	const uint32_t *bm; // global-memory block bitmap
	uint32_t sh1, mult, sh2, sh2b;
	uint32_t hw[kWords];

	TabRef rmk;
	// offsets r < S at which some pattern holds the block ending at tile byte `pos`
	__device__ __forceinline__ uint32_t probe_mask(const ScanArgs &a, uint32_t buf, uint32_t, uint32_t pos) const {
		if (S == 1)
			return 1u;
		const uint32_t blk = mix64(window8(buf, kHalo + pos) >> sh1);
		const uint32_t ri = (uint32_t) (blk * a.prm.r_mult) >> a.prm.r_sh;
		return S > 8 ? rmk.u16(ri) : rmk.u8(ri);
	}
	__device__ __forceinline__ void init(uint32_t smem_tab, const TabRef &rmask, const ScanArgs &a) {
		rmk = rmask;
		bm_s = smem_tab;
		bm = reinterpret_cast<const uint32_t *>(a.front);
		sh1 = a.prm.f1_sh1;
		mult = a.prm.f1_mult;
		sh2 = a.prm.f1_sh2;
		sh2b = a.prm.f1_k == 2 ? sh2 - 5 : sh2; // where the entry's second bit comes from (one bit per entry: the same again)
	}
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) ((g * 32 + b) * S); }
	__device__ __forceinline__ void begin(const ScanArgs &, uint32_t) {
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
	}
	__device__ __forceinline__ void finish_words() {
		constexpr int r = kSamples % 32; // the last word holds fewer than 32 samples: bring them down to bit 0
		if constexpr (r != 0)
			hw[kWords - 1] >>= 32 - r;
	}
	template <int G>
	__device__ __forceinline__ void group(const uint4 &p, const uint4 &c) {
		const uint32_t X[8] = {p.x, p.y, p.z, p.w, c.x, c.y, c.z, c.w};
#pragma unroll
		for (int o = 0; o < 16; o++) {
			if ((G * 16 + o) % S)
				continue;
			const int j = (G * 16 + o) / S;
			// window = X bytes [9+o, 16+o]
			const int bl = 9 + o, bh = 13 + o;
			const uint32_t lo = (bl & 3) ? __byte_perm(X[bl >> 2], X[(bl >> 2) + 1], 0x3210 + 0x1111 * (bl & 3)) : X[bl >> 2];
			const uint32_t hi = (bh & 3) ? __byte_perm(X[bh >> 2], X[(bh >> 2) + 1], 0x3210 + 0x1111 * (bh & 3)) : X[bh >> 2];
			const uint64_t blk = (((uint64_t) hi << 32) | lo) >> sh1;
			const uint32_t h = mix64(blk) * mult, idx = h >> sh2;
			const uint32_t word = GLOBAL ? __ldg(bm + (idx >> 5)) : lds32(bm_s + ((idx >> 3) & ~3u));
			// blocked Bloom filter: both bits of an entry sit in one word
			hw[j / 32] = __funnelshift_r(hw[j / 32], (word >> (idx & 31)) & (word >> ((h >> sh2b) & 31)), 1); // the sample's bit enters from the top
		}
	}
	ACWM_SCAN_GROUPS()
	__device__ __forceinline__ uint32_t count() const {
		uint32_t c = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			c += __popc(hw[g]);
		return c;
	}
	__device__ __forceinline__ void mask_range(uint32_t, uint32_t) {}
};

// ------------------------------------------------------------ dispatch
cudaError_t launch_scan_bytes(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	const acwm_scan_params &p = a.prm;
	if (p.algo == ACWM_ALGO_AC) {
		const bool ex = p.exact_front != 0;
		if (a.front_in_smem)
			return ex ? launch_front<FrontACB<true>, true>(a, threads, smem, grid, st)
					  : launch_front<FrontACB<true>, false>(a, threads, smem, grid, st);
		return ex ? launch_front<FrontACB<false>, true>(a, threads, smem, grid, st)
				  : launch_front<FrontACB<false>, false>(a, threads, smem, grid, st);
	}
	switch (p.stride) {
#define ACWM_WMB(S)                                                                                   \
	case S: return a.front_in_smem ? launch_front<FrontWMB<S, false>, false>(a, threads, smem, grid, st) \
								   : launch_front<FrontWMB<S, true>, false>(a, threads, smem, grid, st);
	ACWM_WMB(16)
	ACWM_WMB(8)
	ACWM_WMB(4)
	ACWM_WMB(2)
	ACWM_WMB(1)
#undef ACWM_WMB
	default: return cudaErrorInvalidValue;
	}
}

} // namespace acwm
