// Scan kernels for the 2-bit path (alphabet <= 4: DNA), hand-written for sm_100a.
//
// One persistent CTA per SM, every WARP owns whole tiles of kWarpTile = 6144 symbols
// and runs its own software pipeline -- no block-wide barrier after start-up:
//
//   global text (1 B/symbol) --LDG.128, coalesced 512 B per warp instruction-->
//   registers --16 symbols -> one 32-bit word (IMAD + PRMT)--> packed tile in shared
//   memory (double-buffered per warp, 64-symbol history in front of the tile)
//   --3 x LDS.128 per lane (lane stride 48 B = 16 B * odd: conflict-free)--> scan.
//
// The loads of tile i+1 are issued in four batches interleaved with the four quarters
// of the scan of tile i, so every warp keeps 2 KB of HBM reads in flight.
//
// Front ends (template parameter):
//   AC<K>  dense DFA with the failure function folded in, K symbols per shared-memory
//          lookup (uint16 entry = next_row << K | K hit bits).  Supersedes the
//          one-symbol goto/failure walk of cuda/cuda_ac.cu:88-95.
//   WM<S>  Wu-Manber block filter sampled every S symbols (SHIFT[block] < S as a
//          bitmap).  Supersedes the divergent skip loop of cuda/cuda_wm.cu:136-176.
// Hits of an exact AC automaton are matches; hits of a depth-truncated automaton and
// WM candidates go through the stage-2 suffix bitmap and the bucket verification.
#include "scan_common.cuh"

namespace acwm {

// ------------------------------------------------------------ pack: 16 symbols -> 32 bits
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return w * 0x01041040u; } // result in the top byte

__device__ __forceinline__ uint32_t pack16(const uint4 r, uint32_t &badacc) {
	badacc |= (r.x | r.y) | (r.z | r.w);
	const uint32_t a = __byte_perm(pack4(r.x), pack4(r.y), 0x7373);
	const uint32_t b = __byte_perm(pack4(r.z), pack4(r.w), 0x7373);
	return __byte_perm(a, b, 0x5410);
}

// Load range of a tile in 16-byte chunks: chunk c (0..387) holds virtual bytes
// [tile*6144 - 64 + 16c, +16) and becomes word c of the warp buffer.
constexpr int kChunks = (int) (kHaloWords + kTileWords); // 388

__device__ __forceinline__ bool tile_is_interior(const ScanArgs &a, uint64_t tile) {
	return tile >= 1 && (tile + 1) * (uint64_t) kWarpTile <= (a.data_hi & ~(uint64_t) 15);
}

// Careful loader for the first / last tiles: chunks outside the text become zero
// words, bytes outside [data_lo, data_hi) are masked before packing.
__device__ __noinline__ void load_tile_edge(const ScanArgs &a, uint64_t tile, uint32_t *buf, uint32_t &badacc) {
	const long long base = (long long) (tile * (uint64_t) kWarpTile) - (long long) kHaloSyms;
	for (int c = (int) lane_id(); c < kChunks; c += 32) {
		const long long off = base + 16ll * c;
		uint32_t word = 0;
		if (off + 16 > (long long) a.data_lo && off < (long long) a.data_hi) {
			uint4 r = ldg_stream16(a.text16 + off);
			uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
			for (int k = 0; k < 16; k++) {
				const long long pos = off + k;
				if (pos < (long long) a.data_lo || pos >= (long long) a.data_hi)
					w[k >> 2] &= ~(0xffu << (8 * (k & 3)));
			}
			word = pack16(make_uint4(w[0], w[1], w[2], w[3]), badacc);
		}
		buf[c] = word;
	}
}

// Fast loader pieces for interior tiles: rounds of 32 chunks (one per lane).
template <int R0, int NR>
struct LoadBatch {
	uint4 v[NR];
	__device__ __forceinline__ void issue(const uint8_t *tile_base /* text16 + tile*6144 - 64 */) {
#pragma unroll
		for (int j = 0; j < NR; j++) {
			const int c = (R0 + j) * 32 + (int) lane_id();
			if (c < kChunks)
				v[j] = ldg_stream16(tile_base + 16 * c);
		}
	}
	__device__ __forceinline__ void store(uint32_t *buf, uint32_t &badacc) {
#pragma unroll
		for (int j = 0; j < NR; j++) {
			const int c = (R0 + j) * 32 + (int) lane_id();
			if (c < kChunks)
				buf[c] = pack16(v[j], badacc);
		}
	}
};

__device__ __forceinline__ void load_tile_sync(const ScanArgs &a, uint64_t tile, uint32_t *buf, uint32_t &badacc) {
	if (tile_is_interior(a, tile)) {
		const uint8_t *tb = a.text16 + tile * (uint64_t) kWarpTile - kHaloSyms;
		LoadBatch<0, 7> b0;
		b0.issue(tb);
		LoadBatch<7, 6> b1;
		b1.issue(tb);
		b0.store(buf, badacc);
		b1.store(buf, badacc);
	} else
		load_tile_edge(a, tile, buf, badacc);
}

// 32-bit window of the 16 symbols ending at buffer symbol index `sidx` (>= 15).
__device__ __forceinline__ uint32_t window16(const uint32_t *buf, uint32_t sidx) {
	const uint32_t bit = 2 * sidx - 30;
	const uint32_t wi = bit >> 5;
	return __funnelshift_r(buf[wi], buf[wi + 1], bit & 31);
}

// ------------------------------------------------------------ front end: AC, K symbols per lookup
template <int K>
struct FrontAC {
	static constexpr int kStrides = (int) kLaneSyms / K;       // 64 / 96 / 192
	static constexpr int kPerQuarter = kStrides / 4;           // 16 / 24 / 48
	static constexpr int kGroup = (K == 3) ? 10 : 32 / K;      // strides per hit word
	static constexpr int kGroupSyms = kGroup * K;              // 30 / 32 / 32
	static constexpr int kWords = ((int) kLaneSyms + kGroupSyms - 1) / kGroupSyms; // 7 / 6 / 6
	static constexpr uint32_t kSymMask2 = ((1u << (2 * K)) - 1) << 1;
	static constexpr uint32_t kRowMask = ~((1u << (2 * K + 1)) - 1);
	static constexpr uint32_t kHitMask = (1u << K) - 1;
	static constexpr int kExpand = 1; // probes per candidate

	const uint8_t *tab; // shared-memory DFA
	uint32_t ent;       // current entry (row << K | hits)
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(const uint8_t *table) { tab = table; }
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) (g * kGroupSyms + b); }

	__device__ __forceinline__ uint32_t step(uint32_t sym2) {
		const uint32_t addr = ((ent << (K + 1)) & kRowMask) | sym2;
		ent = *reinterpret_cast<const uint16_t *>(tab + addr);
		return ent & kHitMask;
	}

	// history: the 64 symbols in front of the lane's chunk (H = 4 words); walk the last
	// K*ceil((D-1)/K) of them from the root so that the state is right at the chunk start.
	__device__ __forceinline__ void begin(const ScanArgs &a, const uint4 H) {
		ent = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
		const uint32_t nwu = (a.prm.depth - 1 + K - 1) / K; // strides
		if (nwu == 0)
			return;
		uint64_t lo = ((uint64_t) H.y << 32) | H.x, hi = ((uint64_t) H.w << 32) | H.z;
		const uint32_t sh = 128 - 2 * K * nwu; // bits to drop from the front
		if (sh >= 64) {
			lo = hi >> (sh - 64);
			hi = 0;
		} else if (sh) {
			lo = (lo >> sh) | (hi << (64 - sh));
			hi >>= sh;
		}
		for (uint32_t i = 0; i < nwu; i++) {
			(void) step(((uint32_t) lo << 1) & kSymMask2);
			lo = (lo >> (2 * K)) | (hi << (64 - 2 * K));
			hi >>= 2 * K;
		}
	}

	template <int Q>
	__device__ __forceinline__ void quarter(const uint32_t (&W)[13]) {
#pragma unroll
		for (int i = Q * kPerQuarter; i < (Q + 1) * kPerQuarter; i++) {
			const int bit = 2 * K * i;
			const int wi = 1 + (bit >> 5), sh = bit & 31;
			uint32_t sym2;
			if (sh + 2 * K <= 32)
				sym2 = (sh >= 1 ? (W[wi] >> (sh - 1)) : (W[wi] << 1)) & kSymMask2;
			else
				sym2 = __funnelshift_r(W[wi], W[wi + 1], sh - 1) & kSymMask2;
			const uint32_t h = step(sym2);
			hw[i / kGroup] += h << (K * (i % kGroup));
		}
	}

	// candidate / hit positions of this lane, tile-relative, ascending
	__device__ __forceinline__ uint32_t count() const {
		uint32_t c = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			c += __popc(hw[g]);
		return c;
	}
	__device__ __forceinline__ void mask_range(uint32_t lo_sym, uint32_t hi_sym) {
		// keep only chunk-relative symbols in [lo_sym, hi_sym)
#pragma unroll
		for (int g = 0; g < kWords; g++) {
			const int base = g * kGroupSyms;
			const int lo = max((int) lo_sym - base, 0), hi = min((int) hi_sym - base, kGroupSyms);
			uint32_t keep = 0;
			if (hi > lo)
				keep = (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1)) << lo;
			hw[g] &= keep;
		}
	}
	template <typename F>
	__device__ __forceinline__ void for_each(F &&f) const {
#pragma unroll
		for (int g = 0; g < kWords; g++) {
			uint32_t w = hw[g];
			while (w) {
				const int b = __ffs(w) - 1;
				w &= w - 1;
				f((uint32_t) (g * kGroupSyms + b));
			}
		}
	}
};

// ------------------------------------------------------------ front end: WM, sampled every S symbols
template <int S, bool HASHED>
struct FrontWM {
	static constexpr int kSamples = (int) kLaneSyms / S;       // 192 / 96 / 48 / 24 / 12
	static constexpr int kPerQuarter = kSamples / 4;
	static constexpr int kWords = (kSamples + 31) / 32;
	static constexpr int kExpand = S;

	const uint32_t *bm; // shared-memory block bitmap
	uint32_t sh1, mult, sh2;
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(const uint8_t *table) { bm = reinterpret_cast<const uint32_t *>(table); }
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) ((g * 32 + b) * S); }

	__device__ __forceinline__ void begin(const ScanArgs &a, const uint4) {
		sh1 = a.prm.f1_sh1;
		mult = a.prm.f1_mult;
		sh2 = a.prm.f1_sh2;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
	}

	template <int Q>
	__device__ __forceinline__ void quarter(const uint32_t (&W)[13]) {
#pragma unroll
		for (int j = Q * kPerQuarter; j < (Q + 1) * kPerQuarter; j++) {
			// 16 symbols ending at chunk symbol c = j*S; W[0] holds chunk symbols -16..-1
			const int bit = 2 * (j * S) + 2;
			const int wi = bit >> 5, sh = bit & 31;
			const uint32_t v = sh ? __funnelshift_r(W[wi], W[wi + 1], sh) : W[wi];
			uint32_t idx = v >> sh1;
			if (HASHED)
				idx = (idx * mult) >> sh2;
			const uint32_t word = bm[idx >> 5];
			hw[j / 32] += ((word >> (idx & 31)) & 1u) << (j % 32);
		}
	}

	__device__ __forceinline__ uint32_t count() const {
		uint32_t c = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			c += __popc(hw[g]);
		return c;
	}
	__device__ __forceinline__ void mask_range(uint32_t, uint32_t) {} // probes are range-checked in verify
	template <typename F>
	__device__ __forceinline__ void for_each(F &&f) const {
#pragma unroll
		for (int g = 0; g < kWords; g++) {
			uint32_t w = hw[g];
			while (w) {
				const int b = __ffs(w) - 1;
				w &= w - 1;
				f((uint32_t) ((g * 32 + b) * S));
			}
		}
	}
};

// ------------------------------------------------------------ the kernel
template <class Front, bool EXACT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) scan_packed_kernel(const __grid_constant__ ScanArgs a) {
	extern __shared__ __align__(16) uint8_t smem[];
	const uint32_t front_bytes16 = (a.front_bytes + 15u) & ~15u;
	uint8_t *s_front = smem;
	uint32_t *s_f2 = reinterpret_cast<uint32_t *>(smem + front_bytes16);
	const uint32_t f2_words = EXACT ? 0 : a.prm.f2_words;
	uint8_t *s_warps = smem + front_bytes16 + ((f2_words * 4 + 15u) & ~15u);

	// tables: global -> shared, once per CTA
	for (uint32_t i = threadIdx.x; i < front_bytes16 / 16; i += THREADS)
		reinterpret_cast<uint4 *>(s_front)[i] = reinterpret_cast<const uint4 *>(a.front)[i];
	for (uint32_t i = threadIdx.x; i < f2_words; i += THREADS)
		s_f2[i] = a.filter2[i];
	__syncthreads();

	const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
	uint32_t *bufs = reinterpret_cast<uint32_t *>(s_warps + warp * kWarpSmemPacked);
	uint16_t *queue = reinterpret_cast<uint16_t *>(bufs + 2 * kBufWords);

	const uint64_t warps_total = (uint64_t) gridDim.x * (THREADS / 32);
	// warps of all CTAs interleave over the text so that concurrently running warps read
	// neighbouring tiles: first tile = warp-in-CTA * gridDim + CTA, then += warps_total
	uint64_t tile = a.tile_lo + (uint64_t) warp * gridDim.x + blockIdx.x;

	uint32_t badacc = 0;
	Emitter em{&a, 0, 0, 0};
	Front fr;
	fr.init(s_front);
	if (lane < 4) { // pad words behind each buffer: read (never used) by the funnel shift of the last window
		bufs[kHaloWords + kTileWords + lane] = 0;
		bufs[kBufWords + kHaloWords + kTileWords + lane] = 0;
	}

	int cur = 0;
	if (tile < a.tile_hi)
		load_tile_sync(a, tile, bufs, badacc);
	__syncwarp();

	for (; tile < a.tile_hi; tile += warps_total) {
		uint32_t *buf = bufs + cur * kBufWords;
		uint32_t *nbuf = bufs + (cur ^ 1) * kBufWords;
		const uint64_t next = tile + warps_total;
		const bool has_next = next < a.tile_hi;
		const bool fast_next = has_next && tile_is_interior(a, next);
		const uint8_t *ntb = a.text16 + next * (uint64_t) kWarpTile - kHaloSyms;

		// lane's words: W[0] = 16 symbols of history, W[1..12] = the 192-symbol chunk
		const uint4 H = *reinterpret_cast<const uint4 *>(buf + 12 * lane);
		uint32_t W[13];
		W[0] = H.w;
		{
			const uint4 c0 = *reinterpret_cast<const uint4 *>(buf + 12 * lane + 4);
			const uint4 c1 = *reinterpret_cast<const uint4 *>(buf + 12 * lane + 8);
			const uint4 c2 = *reinterpret_cast<const uint4 *>(buf + 12 * lane + 12);
			W[1] = c0.x; W[2] = c0.y; W[3] = c0.z; W[4] = c0.w;
			W[5] = c1.x; W[6] = c1.y; W[7] = c1.z; W[8] = c1.w;
			W[9] = c2.x; W[10] = c2.y; W[11] = c2.z; W[12] = c2.w;
		}
		fr.begin(a, H);

		LoadBatch<0, 4> b0;
		LoadBatch<4, 4> b1;
		LoadBatch<8, 4> b2;
		LoadBatch<12, 1> b3;
		if (fast_next)
			b0.issue(ntb);
		fr.template quarter<0>(W);
		if (fast_next) {
			b0.store(nbuf, badacc);
			b1.issue(ntb);
		}
		fr.template quarter<1>(W);
		if (fast_next) {
			b1.store(nbuf, badacc);
			b2.issue(ntb);
		}
		fr.template quarter<2>(W);
		if (fast_next) {
			b2.store(nbuf, badacc);
			b3.issue(ntb);
		}
		fr.template quarter<3>(W);
		if (fast_next)
			b3.store(nbuf, badacc);

		// ---- hits / candidates of this tile
		em.tile = tile;
		const uint64_t tile_start = tile * (uint64_t) kWarpTile;
		const uint64_t end_lo = a.report_lo; // first end position this scan reports (>= data_lo + m_min - 1)
		if constexpr (EXACT) {
			const bool inner = tile_start >= end_lo && tile_start + kWarpTile <= a.data_hi;
			if (!inner) { // first / last tiles: drop ends outside [end_lo, data_hi)
				const uint64_t cs = tile_start + (uint64_t) lane * kLaneSyms;
				const uint32_t lo_s = cs >= end_lo ? 0u : (uint32_t) min((uint64_t) kLaneSyms, end_lo - cs);
				const uint32_t hi_s = cs >= a.data_hi ? 0u : (uint32_t) min((uint64_t) kLaneSyms, a.data_hi - cs);
				fr.mask_range(lo_s, hi_s);
			}
			const uint32_t cnt = fr.count();
			if (__any_sync(kFull, cnt != 0)) {
				const uint32_t incl = warp_incl_scan(cnt);
				const uint32_t total = __shfl_sync(kFull, incl, 31);
				if (a.want_positions) {
					unsigned long long slot = 0;
					if (lane == 0)
						slot = atomicAdd(&a.ctl->cursor, (unsigned long long) total);
					slot = __shfl_sync(kFull, slot, 0) + (incl - cnt);
					uint32_t rank = incl - cnt;
					fr.for_each([&](uint32_t s) {
						if (slot < a.cap)
							a.staging[slot] = encode_stage(tile, rank, lane * kLaneSyms + s);
						slot++;
						rank++;
					});
				}
				em.tile_rank += total;
				em.warp_count += total;
			}
		} else {
			uint32_t cnt = fr.count();
			while (__any_sync(kFull, cnt != 0)) {
				// round: queue up to kQueueCap candidates in lane order, then verify them densely
				const uint32_t incl = warp_incl_scan(cnt);
				const uint32_t excl = incl - cnt;
				const uint32_t total = min(__shfl_sync(kFull, incl, 31), kQueueCap);
				{
					uint32_t k = excl, taken = 0;
#pragma unroll
					for (int g = 0; g < Front::kWords; g++) {
						uint32_t w = fr.hw[g];
						while (w && k < kQueueCap) {
							const int b = __ffs(w) - 1;
							w &= w - 1;
							queue[k++] = (uint16_t) (lane * kLaneSyms + Front::sym_of(g, b));
							taken++;
						}
						fr.hw[g] = w;
					}
					cnt -= taken;
				}
				__syncwarp();
				const uint32_t probes = total * Front::kExpand;
				for (uint32_t base = 0; base < probes; base += 32) {
					const uint32_t i = base + lane;
					uint32_t mult = 0, pos = 0;
					if (i < probes) {
						pos = (uint32_t) queue[i / Front::kExpand] + (i % Front::kExpand);
						const uint32_t v = window16(buf, kHaloSyms + pos);
						const uint32_t key = v >> (32 - 2 * a.prm.b2);
						const uint32_t i2 = (uint32_t) (key * a.prm.f2_mult) >> a.prm.f2_sh;
						if ((s_f2[i2 >> 5] >> (i2 & 31)) & 1u)
							mult = verify_window(a, key, tile_start + pos);
					}
					em.emit(mult, pos);
				}
				__syncwarp();
			}
		}
		em.end_tile();

		if (has_next && !fast_next)
			load_tile_edge(a, next, nbuf, badacc);
		__syncwarp();
		cur ^= 1;
	}

	// ---- per-warp totals
	if (lane == 0 && em.warp_count)
		atomicAdd(&a.ctl->count, em.warp_count);
	badacc &= 0xFCFCFCFCu;
	if (__any_sync(kFull, badacc != 0) && lane == 0)
		atomicOr(&a.ctl->bad_text, 1u);
}


// ------------------------------------------------------------ dispatch
template <class Front, bool EXACT>
static cudaError_t launch_one(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	cudaError_t e;
#define ACWM_LAUNCH(T)                                                                                          \
	e = cudaFuncSetAttribute(scan_packed_kernel<Front, EXACT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
			(int) smem);                                                                                        \
	if (e != cudaSuccess)                                                                                      \
		return e;                                                                                              \
	scan_packed_kernel<Front, EXACT, T><<<grid, T, smem, st>>>(a);                                             \
	return cudaGetLastError();
	switch (threads) {
	case 1024: ACWM_LAUNCH(1024)
	case 768: ACWM_LAUNCH(768)
	case 512: ACWM_LAUNCH(512)
	default: ACWM_LAUNCH(256)
	}
#undef ACWM_LAUNCH
}

cudaError_t launch_scan_packed(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	const acwm_scan_params &p = a.prm;
	if (p.algo == ACWM_ALGO_AC) {
		const bool ex = p.exact_front != 0;
		switch (p.stride) {
		case 3: return ex ? launch_one<FrontAC<3>, true>(a, threads, smem, grid, st)
						  : launch_one<FrontAC<3>, false>(a, threads, smem, grid, st);
		case 2: return ex ? launch_one<FrontAC<2>, true>(a, threads, smem, grid, st)
						  : launch_one<FrontAC<2>, false>(a, threads, smem, grid, st);
		case 1: return ex ? launch_one<FrontAC<1>, true>(a, threads, smem, grid, st)
						  : launch_one<FrontAC<1>, false>(a, threads, smem, grid, st);
		default: return cudaErrorInvalidValue;
		}
	}
	const bool hashed = p.f1_mult != 1;
	switch (p.stride) {
#define ACWM_WM(S)                                                                      \
	case S: return hashed ? launch_one<FrontWM<S, true>, false>(a, threads, smem, grid, st) \
						  : launch_one<FrontWM<S, false>, false>(a, threads, smem, grid, st);
	ACWM_WM(16)
	ACWM_WM(8)
	ACWM_WM(4)
	ACWM_WM(2)
	ACWM_WM(1)
#undef ACWM_WM
	default: return cudaErrorInvalidValue;
	}
}

} // namespace acwm
