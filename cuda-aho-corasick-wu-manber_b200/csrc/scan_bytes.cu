// Scan kernels for the bytes path (alphabet > 4: one byte per symbol), sm_100a.
//
// Same warp-autonomous organisation as the 2-bit path, but the text tile stays raw:
// every warp double-buffers tiles of kWarpTileB = 3584 bytes (+64 bytes of history)
// in shared memory and fills them with ONE TMA bulk copy per tile
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) issued by lane 0 a
// whole tile ahead; lanes read their 112-byte chunk with 7 conflict-free LDS.128.
//
// Front ends:
//   ACB   dense DFA, failure function folded in, one class-compressed symbol per
//         lookup; table in shared memory (uint16) or, when it does not fit, in global
//         memory served from L2 (uint32; the host pins it with an access-policy window).
//   WMB<S> Wu-Manber block filter (up to 8 bytes, mixed to 32 bits) sampled every S bytes.
#include "scan_common.cuh"

namespace acwm {

// ------------------------------------------------------------ mbarrier / TMA bulk helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
			"{\n"
			".reg .pred p;\n"
			"WAIT_%=:\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
			"@p bra DONE_%=;\n"
			"bra WAIT_%=;\n"
			"DONE_%=:\n"
			"}\n" ::"r"(smem_u32(bar)),
			"r"(parity)
			: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
						 smem_u32(dst)),
				 "l"(src), "r"(bytes), "r"(smem_u32(bar))
				 : "memory");
}

constexpr uint32_t kLoadBytesB = kHaloBytes + kWarpTileB; // 3648

__device__ __forceinline__ bool tile_is_interior_b(const ScanArgs &a, uint64_t tile) {
	return tile >= 1 && (tile + 1) * (uint64_t) kWarpTileB <= (a.data_hi & ~(uint64_t) 15);
}

// Careful loader for first / last tiles: never touches a byte outside [data_lo, data_hi).
__device__ __noinline__ void load_tile_edge_b(const ScanArgs &a, uint64_t tile, uint8_t *buf) {
	const long long base = (long long) (tile * (uint64_t) kWarpTileB) - (long long) kHaloBytes;
	for (int c = (int) lane_id(); c < (int) (kLoadBytesB / 16); c += 32) {
		const long long off = base + 16ll * c;
		uint4 r = make_uint4(0, 0, 0, 0);
		if (off >= (long long) a.data_lo && off + 16 <= (long long) a.data_hi)
			r = ldg_stream16(a.text16 + off);
		else if (off + 16 > (long long) a.data_lo && off < (long long) a.data_hi) {
			uint32_t w[4] = {0, 0, 0, 0};
			for (int k = 0; k < 16; k++) {
				const long long pos = off + k;
				if (pos >= (long long) a.data_lo && pos < (long long) a.data_hi)
					w[k >> 2] |= (uint32_t) a.text16[pos] << (8 * (k & 3));
			}
			r = make_uint4(w[0], w[1], w[2], w[3]);
		}
		*reinterpret_cast<uint4 *>(buf + 16 * c) = r;
	}
}

// 64-bit window of the 8 bytes ending at buffer byte index bidx (>= 7).
__device__ __forceinline__ uint64_t window8(const uint8_t *buf, uint32_t bidx) {
	const uint32_t b0 = bidx - 7;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(buf) + (b0 >> 2);
	const uint32_t sh = (b0 & 3) * 8;
	const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
	const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
	return ((uint64_t) hi << 32) | lo;
}

__device__ __forceinline__ uint32_t mix64(uint64_t v) {
	return (uint32_t) v * 0x9E3779B1u + (uint32_t) (v >> 32) * 0x85EBCA77u;
}

// ------------------------------------------------------------ front end: AC over bytes
template <bool IN_SMEM>
struct FrontACB {
	static constexpr int kWords = 4; // 112 hit bits
	static constexpr int kExpand = 1;
	const uint8_t *tab;
	uint32_t ent, lognc, alpha;
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(const uint8_t *smem_tab, const ScanArgs &a) {
		tab = IN_SMEM ? smem_tab : a.front;
		lognc = 31 - __clz(a.prm.n_classes);
		alpha = min(a.prm.alphabet, 255u);
	}
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) (g * 32 + b); }

	__device__ __forceinline__ uint32_t step(uint32_t byte) {
		const uint32_t cls = min(byte, alpha);
		const uint32_t idx = ((ent >> 1) << lognc) + cls;
		if (IN_SMEM)
			ent = reinterpret_cast<const uint16_t *>(tab)[idx];
		else
			ent = __ldg(reinterpret_cast<const uint32_t *>(tab) + idx);
		return ent & 1u;
	}
	__device__ __forceinline__ void begin(const ScanArgs &a, const uint8_t *chunk) {
		ent = 0;
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
		const uint32_t nwu = a.prm.depth - 1;
		for (uint32_t i = 0; i < nwu; i++)
			(void) step(chunk[(int) i - (int) nwu]);
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
template <int S>
struct FrontWMB {
	static constexpr int kSamples = (int) kLaneBytes / S; // 112 / 56 / 28 / 14 / 7
	static constexpr int kWords = (kSamples + 31) / 32;
	static constexpr int kExpand = S;
	const uint32_t *bm;
	uint32_t sh1, mult, sh2;
	uint32_t hw[kWords];

	__device__ __forceinline__ void init(const uint8_t *smem_tab, const ScanArgs &a) {
		bm = reinterpret_cast<const uint32_t *>(smem_tab);
		sh1 = a.prm.f1_sh1;
		mult = a.prm.f1_mult;
		sh2 = a.prm.f1_sh2;
	}
	static __device__ __forceinline__ uint32_t sym_of(int g, int b) { return (uint32_t) ((g * 32 + b) * S); }
	__device__ __forceinline__ void begin(const ScanArgs &, const uint8_t *) {
#pragma unroll
		for (int g = 0; g < kWords; g++)
			hw[g] = 0;
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
			const uint32_t idx = (uint32_t) (mix64(blk) * mult) >> sh2;
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
	__device__ __forceinline__ void mask_range(uint32_t, uint32_t) {}
};

// ------------------------------------------------------------ the kernel
template <class Front, bool EXACT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) scan_bytes_kernel(const __grid_constant__ ScanArgs a) {
	extern __shared__ __align__(16) uint8_t smem[];
	const uint32_t front_smem = a.front_in_smem ? ((a.front_bytes + 15u) & ~15u) : 0u;
	uint8_t *s_front = smem;
	uint32_t *s_f2 = reinterpret_cast<uint32_t *>(smem + front_smem);
	const uint32_t f2_words = EXACT ? 0 : a.prm.f2_words;
	uint8_t *s_warps = smem + front_smem + ((f2_words * 4 + 15u) & ~15u);

	for (uint32_t i = threadIdx.x; i < front_smem / 16; i += THREADS)
		reinterpret_cast<uint4 *>(s_front)[i] = reinterpret_cast<const uint4 *>(a.front)[i];
	for (uint32_t i = threadIdx.x; i < f2_words; i += THREADS)
		s_f2[i] = a.filter2[i];

	const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
	uint8_t *wbase = s_warps + warp * kWarpSmemBytes;
	uint8_t *bufs = wbase; // 2 x kBufBytesB
	uint16_t *queue = reinterpret_cast<uint16_t *>(wbase + 2 * kBufBytesB);
	uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + 2 * kBufBytesB + kQueueCap * 2);
	if (lane == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
	}
	if (lane < 4) { // pad behind each buffer (read, never used, by window8 of the last bytes)
		reinterpret_cast<uint32_t *>(bufs + kLoadBytesB)[lane] = 0;
		reinterpret_cast<uint32_t *>(bufs + kBufBytesB + kLoadBytesB)[lane] = 0;
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();

	const uint64_t warps_total = (uint64_t) gridDim.x * (THREADS / 32);
	uint64_t tile = a.tile_lo + (uint64_t) warp * gridDim.x + blockIdx.x;
	Emitter em{&a, 0, 0, 0};
	Front fr;
	fr.init(s_front, a);
	uint32_t phase[2] = {0, 0};

	// issue the load of `t` into buffer b (lane 0 for TMA; all lanes for edge tiles)
	auto issue = [&](uint64_t t, int b) -> bool {
		uint8_t *dst = bufs + b * kBufBytesB;
		if (tile_is_interior_b(a, t)) {
			if (lane == 0) {
				// generic-proxy reads of this buffer are done (__syncwarp before us); order them
				// before the async-proxy write
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				mbar_expect_tx(&bars[b], kLoadBytesB);
				tma_bulk_g2s(dst, a.text16 + t * (uint64_t) kWarpTileB - kHaloBytes, kLoadBytesB, &bars[b]);
			}
			return true;
		}
		load_tile_edge_b(a, t, dst);
		return false;
	};

	bool cur_tma = false;
	int cur = 0;
	if (tile < a.tile_hi)
		cur_tma = issue(tile, 0);
	__syncwarp();

	for (; tile < a.tile_hi; tile += warps_total) {
		uint8_t *buf = bufs + cur * kBufBytesB;
		const uint64_t next = tile + warps_total;
		bool next_tma = false;
		if (next < a.tile_hi)
			next_tma = issue(next, cur ^ 1);
		if (cur_tma) {
			mbar_wait(&bars[cur], phase[cur]);
			phase[cur] ^= 1;
		}
		__syncwarp();

		const uint8_t *chunk = buf + kHaloBytes + lane * kLaneBytes;
		fr.begin(a, chunk);
		uint4 prev = *reinterpret_cast<const uint4 *>(chunk - 16);
		uint4 c;
#define ACWM_GROUP(G)                                          \
	c = *reinterpret_cast<const uint4 *>(chunk + 16 * G);      \
	fr.template group<G>(prev, c);                             \
	prev = c;
		ACWM_GROUP(0)
		ACWM_GROUP(1)
		ACWM_GROUP(2)
		ACWM_GROUP(3)
		ACWM_GROUP(4)
		ACWM_GROUP(5)
		ACWM_GROUP(6)
#undef ACWM_GROUP

		em.tile = tile;
		const uint64_t tile_start = tile * (uint64_t) kWarpTileB;
		if constexpr (EXACT) {
			const uint64_t end_lo = a.report_lo;
			const bool inner = tile_start >= end_lo && tile_start + kWarpTileB <= a.data_hi;
			if (!inner) {
				const uint64_t cs = tile_start + (uint64_t) lane * kLaneBytes;
				const uint32_t lo_s = cs >= end_lo ? 0u : (uint32_t) min((uint64_t) kLaneBytes, end_lo - cs);
				const uint32_t hi_s = cs >= a.data_hi ? 0u : (uint32_t) min((uint64_t) kLaneBytes, a.data_hi - cs);
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
#pragma unroll
					for (int g = 0; g < Front::kWords; g++) {
						uint32_t w = fr.hw[g];
						while (w) {
							const int b = __ffs(w) - 1;
							w &= w - 1;
							if (slot < a.cap)
								a.staging[slot] = encode_stage(tile, rank, lane * kLaneBytes + Front::sym_of(g, b));
							slot++;
							rank++;
						}
					}
				}
				em.tile_rank += total;
				em.warp_count += total;
			}
		} else {
			uint32_t cnt = fr.count();
			while (__any_sync(kFull, cnt != 0)) {
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
							queue[k++] = (uint16_t) (lane * kLaneBytes + Front::sym_of(g, b));
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
						const uint64_t win = window8(buf, kHaloBytes + pos);
						const uint32_t key = mix64(win >> (64 - 8 * a.prm.b2));
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
		__syncwarp();
		cur ^= 1;
		cur_tma = next_tma;
	}

	if (lane == 0 && em.warp_count)
		atomicAdd(&a.ctl->count, em.warp_count);
}

// ------------------------------------------------------------ dispatch
template <class Front, bool EXACT>
static cudaError_t launch_one_b(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	cudaError_t e;
#define ACWM_LAUNCH(T)                                                                                        \
	e = cudaFuncSetAttribute(scan_bytes_kernel<Front, EXACT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
			(int) smem);                                                                                      \
	if (e != cudaSuccess)                                                                                    \
		return e;                                                                                            \
	scan_bytes_kernel<Front, EXACT, T><<<grid, T, smem, st>>>(a);                                            \
	return cudaGetLastError();
	switch (threads) {
	case 512: ACWM_LAUNCH(512)
	case 384: ACWM_LAUNCH(384)
	case 256: ACWM_LAUNCH(256)
	default: ACWM_LAUNCH(128)
	}
#undef ACWM_LAUNCH
}

cudaError_t launch_scan_bytes(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	const acwm_scan_params &p = a.prm;
	if (p.algo == ACWM_ALGO_AC) {
		const bool ex = p.exact_front != 0;
		if (a.front_in_smem)
			return ex ? launch_one_b<FrontACB<true>, true>(a, threads, smem, grid, st)
					  : launch_one_b<FrontACB<true>, false>(a, threads, smem, grid, st);
		return ex ? launch_one_b<FrontACB<false>, true>(a, threads, smem, grid, st)
				  : launch_one_b<FrontACB<false>, false>(a, threads, smem, grid, st);
	}
	switch (p.stride) {
	case 16: return launch_one_b<FrontWMB<16>, false>(a, threads, smem, grid, st);
	case 8: return launch_one_b<FrontWMB<8>, false>(a, threads, smem, grid, st);
	case 4: return launch_one_b<FrontWMB<4>, false>(a, threads, smem, grid, st);
	case 2: return launch_one_b<FrontWMB<2>, false>(a, threads, smem, grid, st);
	case 1: return launch_one_b<FrontWMB<1>, false>(a, threads, smem, grid, st);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace acwm
