// Device-side pieces shared by the 2-bit and the bytes scan kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/acwm.h"
#include "geometry.hpp"

namespace acwm {

// Device control block of one matcher (zeroed before every scan).
struct Control {
	unsigned long long count;    // |M|
	unsigned long long cursor;   // staging slots handed out
	unsigned long long written;  // positions written by the finalize pass
	unsigned int bad_text;       // OR of (byte & 0xFC) over the text, 2-bit path
	unsigned int overflow;       // staging capacity exceeded
};

struct ScanArgs {
	const uint8_t *text16;       // 16-byte aligned base of the "virtual" text
	uint64_t data_lo, data_hi;   // real text occupies virtual [data_lo, data_hi), data_lo < 16
	uint64_t report_lo;          // matches ending before this virtual position are not reported
	uint64_t tile_lo, tile_hi;   // warp tiles [tile_lo, tile_hi) of the virtual text are scanned by this launch
	const uint8_t *front;        // front-end table (global copy)
	uint32_t front_bytes;
	uint32_t front_in_smem;
	const uint32_t *filter2;     // stage-2 bitmap (global copy)
	const uint32_t *bucket_start;
	const acwm_ventry *entries;
	const uint8_t *patterns;
	acwm_scan_params prm;
	Control *ctl;
	uint64_t *staging;           // [tile:28 | rank:22 | pos:13]
	uint64_t cap;
	uint32_t *tile_count;        // matches per tile (when want_positions)
	int want_positions;
};

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(kFull, v, d);
		if ((int) lane_id() >= d)
			v += t;
	}
	return v;
}

__device__ __forceinline__ uint4 ldg_stream16(const uint8_t *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
			: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

__device__ __forceinline__ uint64_t encode_stage(uint64_t tile, uint32_t rank, uint32_t pos) {
	return (tile << (kRankBits + kPosBits)) | ((uint64_t) rank << kPosBits) | pos;
}

// Warp-collective emission of matches found at `pos` (tile-relative) with
// multiplicity `mult` (0 = none) per lane, lanes in ascending position order.
// Returns the number emitted; advances tile_rank.
struct Emitter {
	const ScanArgs *a;
	uint64_t tile;
	uint32_t tile_rank;
	unsigned long long warp_count;

	__device__ __forceinline__ void emit(uint32_t mult, uint32_t pos) {
		const unsigned ball = __ballot_sync(kFull, mult != 0);
		if (!ball)
			return;
		uint32_t excl, total;
		if (__all_sync(kFull, mult <= 1)) {
			excl = __popc(ball & ((1u << lane_id()) - 1));
			total = __popc(ball);
		} else {
			const uint32_t incl = warp_incl_scan(mult);
			excl = incl - mult;
			total = __shfl_sync(kFull, incl, 31);
		}
		if (a->want_positions) {
			unsigned long long slot0 = 0;
			if (lane_id() == 0)
				slot0 = atomicAdd(&a->ctl->cursor, (unsigned long long) total);
			slot0 = __shfl_sync(kFull, slot0, 0);
			for (uint32_t i = 0; i < mult; i++) {
				const unsigned long long slot = slot0 + excl + i;
				if (slot < a->cap)
					a->staging[slot] = encode_stage(tile, tile_rank + excl + i, pos);
			}
		}
		tile_rank += total;
		warp_count += total;
	}

	__device__ __forceinline__ void end_tile() {
		if (a->want_positions && lane_id() == 0)
			a->tile_count[tile] = tile_rank;
		tile_rank = 0;
	}
};

// Exact check of the window ending at virtual position e against every pattern of
// the bucket `key` hashes to (the HASH -> PREFIX -> compare tail of Wu-Manber,
// wu/wu.c:81-99, also used for the hits of a depth-truncated AC automaton).
// Returns how many distinct patterns end at e.
__device__ __forceinline__ uint32_t verify_window(const ScanArgs &a, uint32_t key, uint64_t e) {
	const acwm_scan_params &p = a.prm;
	const uint32_t b = (uint32_t) (key * p.hb_mult) >> p.hb_sh;
	const uint32_t lo = __ldg(a.bucket_start + b), hi = __ldg(a.bucket_start + b + 1);
	uint32_t mult = 0;
	for (uint32_t i = lo; i < hi; i++) {
		const acwm_ventry en = a.entries[i];
		if (en.key != key)
			continue;
		const uint32_t len = en.len & 0x7fffffffu;
		if (e + 1 < a.data_lo + len || e >= a.data_hi || e < a.report_lo)
			continue; // window would start before the text / end after it / not ours to report
		bool ok = true;
		if (!(en.len >> 31)) {
			const uint8_t *t = a.text16 + (e + 1 - len);
			const uint8_t *q = a.patterns + en.offset;
			for (uint32_t k = 0; k < len; k++)
				if (t[k] != q[k]) {
					ok = false;
					break;
				}
		}
		mult += ok;
	}
	return mult;
}

} // namespace acwm
