// The scan kernel skeleton (sm_100a), shared by every front end.
//
// ONE launch does the whole job -- supersedes the reference's kernel + D2H of 7680 per-thread
// counters + host sum (cuda/cuda_ac.cu:654-673) -- and no CTA ever waits for the whole grid:
//
//   1. scan     every persistent CTA (one or two per SM) owns a contiguous span of warp tiles; every warp
//               runs its own TMA pipeline over the span (cp.async.bulk global -> shared,
//               mbarrier complete_tx, no block-wide barrier), walks its front end over the tile,
//               checks the candidates (warp-cooperatively when they are few) and stages each match as
//               [tile | rank-in-tile | pos-in-tile] in blocks it reserves from the launch's cursor
//               (doubling sizes, logged per warp in shared memory);
//   2. order    the CTA turns the per-tile counts of its span into exclusive offsets (shared memory),
//               publishes the span total as a tagged word and looks back over the totals of the spans
//               in front of it; then every warp moves its own staged matches to their sorted slots;
//   3. publish  a CTA's arrival is ONE 64-bit atomic that also carries its match count; the last
//               CTA to arrive writes the result block.  The working counters exist four times
//               (launch k uses copy k % 4 and zeroes copy k + 1 as it starts), the scratch arrays three
//               times (by launch number): no memset node between scans.
// Overlap mode (acwm_set_overlap): the launch is a programmatic dependent launch -- CTAs of scan k + 1 become
// resident beside (two half-size CTAs per SM) or right after (one CTA per SM) the CTAs of scan k and run their
// whole scan phase on their own copies of the scratch state; thread 0 waits for scan k (griddepcontrol.wait)
// only before the arrival.  With two CTAs per SM every SM always has one CTA scanning while the other is in its
// prologue or its ordering epilogue.  At most three scans are in flight (see Work in scan_common.cuh).
//   4. exchange (multi-GPU) the publishing thread stores the rank's count into every peer's mailbox
//               over NVLink (system-scope stores on peer-mapped memory) and, at its very end, sums
//               what the peers left in its own mailbox for the PREVIOUS scan: an all-reduce of the
//               8-byte count, pipelined by one scan, without another launch or a lock-step wait.
#pragma once
#include <atomic>
#include <cstdlib>

#include "scan_common.cuh"

namespace acwm {

// Careful loader for the first / last tiles of the text: never touches a byte outside
// [data_lo, data_hi); everything else in the buffer becomes zero.
static __device__ __noinline__ void load_tile_edge(const ScanArgs &a, uint64_t tile, uint8_t *buf) {
	const long long base = (long long) (tile * (uint64_t) kTile) - (long long) kHalo;
	for (int c = (int) lane_id(); c < (int) (kLoadBytes / 16); c += 32) {
		const long long off = base + 16ll * c;
		if (off >= (long long) a.data_lo && off + 16 <= (long long) a.data_hi) {
			// whole 16-byte pieces: asynchronous copies, all in flight together, no registers held
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + 16 * c)), "l"(a.text16 + off) : "memory");
			continue;
		}
		uint4 r = make_uint4(0, 0, 0, 0);
		if (off + 16 > (long long) a.data_lo && off < (long long) a.data_hi) { // straddles an end of the text
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
	asm volatile("cp.async.wait_all;" ::: "memory"); // the caller's __syncwarp publishes the tile to the warp
}

// Pre-packed text (4 symbols per byte, ScanArgs.packed_in): the slot receives [16 bytes of history][896 bytes of
// tile]; the device buffer is zero-padded to a multiple of 16 bytes, so only the first tile needs care.
static __device__ __noinline__ void load_tile_edge_packed(const ScanArgs &a, uint64_t tile, uint8_t *buf) {
	const long long base = (long long) (tile * (uint64_t) (kTile / 4)) - (long long) (kHalo / 4);
	const long long end = (long long) (((a.data_hi + 63) & ~(uint64_t) 63) / 4);
	for (int c = (int) lane_id(); c < (int) (kLoadBytes / 4 / 16); c += 32) {
		const long long off = base + 16ll * c;
		if (off >= 0 && off + 16 <= end)
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + 16 * c)), "l"(a.text16 + off) : "memory");
		else
			*reinterpret_cast<uint4 *>(buf + 16 * c) = make_uint4(0, 0, 0, 0);
	}
	asm volatile("cp.async.wait_all;" ::: "memory");
}

__device__ __forceinline__ bool tile_is_interior(const ScanArgs &a, uint64_t tile) {
	if (a.packed_in)
		return tile >= 1 && (tile + 1) * (uint64_t) kTile <= ((a.data_hi + 63) & ~(uint64_t) 63);
	return tile >= 1 && (tile + 1) * (uint64_t) kTile <= (a.data_hi & ~(uint64_t) 15);
}

template <class Front, bool EXACT, int THREADS, int MINB, int STAGES = Front::kPacked ? 1 : 2>
__global__ void __launch_bounds__(THREADS, MINB) scan_kernel(const __grid_constant__ ScanArgs a) {
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr uint32_t W = THREADS / 32;
	constexpr bool kPacked = Front::kPacked;
	constexpr bool kPk = kPacked && !EXACT; // the 2-bit copy of the tile exists for the verification windows only
	// [CTA scratch 1 KiB][per-warp areas][front table][offset masks][stage-2 bitmap][per-tile counts]
	// The scan loop addresses all of it by 32-bit shared-space addresses (see scan_common.cuh).
	// The dynamic shared memory of a kernel without static shared memory starts at shared address kDynSmemBase (the
	// first KiB of the window belongs to the system): a CONSTANT, and so are the per-warp areas (base + warp * size)
	// and the front table behind them: these addresses fold into the immediate field of their LDS / STS / SYNCS.
	// The tables behind the front table start where the host says (ScanArgs.s_rmask .. s_cnt: one constant-bank
	// read each).  Both checked here, loudly.
	constexpr uint32_t sb = kDynSmemBase;
	constexpr uint32_t stages = STAGES;
	static_assert(!kPacked || STAGES == 1, "the 2-bit path has one slot per warp");
	// ring depth of the per-warp tile pipeline: the 2-bit path copies a tile into registers and re-arms its ONE slot
	// while it walks; the bytes path walks the raw tile in place and either loads the next one into a second slot
	// meanwhile (12-16 warps per SM) or has one slot per warp and more warps to cover the load (20 warps)
	constexpr uint32_t s_warps = sb + kSmemReserve;
	constexpr uint32_t s_front = s_warps + W * warp_smem_bytes(stages, kPk);
	const uint32_t tab_bar = sb;
	const uint32_t s_next = sb + 16; // next unclaimed tile of this CTA's span
	uint32_t *s_bad = reinterpret_cast<uint32_t *>(smem + 20);
	unsigned long long *s_count = reinterpret_cast<unsigned long long *>(smem + 24); // matches of this CTA
	unsigned long long *s_misc = reinterpret_cast<unsigned long long *>(smem + 32);  // [0] cursor, [1] append base
	const uint32_t front_smem = a.front_in_smem ? ((a.front_bytes + 15u) & ~15u) : 0u;
	// tables too large for shared memory stay in global memory (L2-resident: access-policy window + evict_first text)
	const uint32_t rm_bytes = (EXACT || !a.prm.r_in_smem) ? 0u : ((a.prm.r_entries * a.prm.r_entry_bytes + 15u) & ~15u);
	const uint32_t f2_bytes = (EXACT || !a.prm.f2_in_smem) ? 0u : ((a.prm.f2_words * 4u + 15u) & ~15u);
	const uint32_t s_rmask = a.s_rmask, s_f2 = a.s_f2, s_cnt = a.s_cnt;
	if (smem_u32(smem) != sb || s_rmask != s_front + front_smem || s_f2 != s_rmask + rm_bytes || s_cnt != s_f2 + f2_bytes)
		__trap();

	const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
	const uint32_t wbase = s_warps + warp * warp_smem_bytes(stages, kPk);
	const uint32_t pk = wbase + stages * kBufBytes;
	const uint32_t bars = pk + (kPk ? kPackWords * 4 : 0);
	const uint32_t lst = bars + 16 * kMaxStages;      // match positions of the current tile (uint16)
	const uint32_t wlog = lst + 2 * kListCap;         // this warp's staging reservations (uint64)
	uint32_t *s_cnt_p = reinterpret_cast<uint32_t *>(smem + (s_cnt - sb)); // a.cnt_cap words

	// this CTA's span of warp tiles; the warps claim its tiles one at a time (shared-memory ticket)
	const uint64_t cta_lo = min(a.tile_lo + (uint64_t) blockIdx.x * a.tiles_per_cta, a.tile_hi);
	const uint64_t cta_hi = min(cta_lo + a.tiles_per_cta, a.tile_hi);
	const uint32_t n_b = (uint32_t) (cta_hi - cta_lo);
	const uint32_t tab_bytes = front_smem + rm_bytes + f2_bytes;

	// ---- prologue: every warp arms its own barriers and starts its first tile(s) at once (static claims
	// warp, W + warp, ..); the CTA meets only afterwards, under the copies
	if (threadIdx.x == 0) {
		trace_mark(a, 0);
		if (a.trace) { // which SM this CTA runs on (word 11): who shares an SM with whom in overlap mode
			uint32_t smid;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
			a.trace[(size_t) blockIdx.x * kTraceWords + 11] = smid;
		}
		mbar_init(tab_bar, 1);
		sts32(s_next, stages * W);
		*s_bad = 0;
		*s_count = 0;
	}
	if (lane == 0)
		for (uint32_t s = 0; s < stages; s++)
			mbar_init(bars + 8 * s, 1);
	if (lane < 4)
		for (uint32_t s = 0; s < stages; s++) // pad behind each buffer: read, never used
			sts32(wbase + s * kBufBytes + kLoadBytes + 4 * lane, 0);
	if (kPk && lane < 3)
		sts32(pk + 4 * (kPackWords - 3 + lane), 0);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();

	// Per-warp pipeline state, all in registers: slot s holds (or is receiving) tile slot_idx[s] of the span.  Every
	// tile completes one phase of its slot's mbarrier: a tile that lies inside the text arrives by ONE TMA bulk copy
	// (complete_tx), the first / last tiles of the text go through the careful loader, which then arrives itself.
	uint32_t slot_idx[stages] = {}, slot_phase[stages] = {}, slot_buf[stages] = {}, slot_bar[stages] = {};
	// what a tile's copy needs, computed once: tiles [1, int_hi) lie inside the text, tile t is loaded from
	// tile0 + t * tile_stride (load_bytes bytes: history + tile)
	const bool packed_in = kPacked && a.packed_in;
	const uint32_t int_hi = (uint32_t) ((packed_in ? ((a.data_hi + 63) & ~(uint64_t) 63) : (a.data_hi & ~(uint64_t) 15)) / kTile);
	const uint32_t int_span = int_hi ? int_hi - 1u : 0u; // number of interior tiles
	const uint32_t tile_stride = packed_in ? kTile / 4 : kTile, load_bytes = packed_in ? kLoadBytes / 4 : kLoadBytes;
	const uint8_t *tile0 = a.text16 - (packed_in ? kHalo / 4 : kHalo);
	const uint32_t cta_lo32 = (uint32_t) cta_lo; // tile numbers fit 28 bits (staging entries)
	// start loading tile idx of the span (if there is one) into ring slot s.  The warp has read the slot's previous
	// tile (into registers, or walked it) and met at a __syncwarp: the copy may overwrite it (the same write-after-
	// read hand-over as a consumer-release / producer-acquire TMA pipeline; no proxy fence, whose MEMBAR would make
	// lane 0 wait for all its stores in flight)
	auto issue = [&](uint32_t s, uint32_t idx) {
		slot_idx[s] = idx;
		if (idx >= n_b)
			return;
		const uint32_t t = cta_lo32 + idx;
		if (t - 1u < int_span) { // 1 <= t < int_hi, one unsigned compare
			if (lane == 0) {
				mbar_expect_tx(slot_bar[s], load_bytes);
				tma_bulk_g2s(slot_buf[s], tile0 + (uint64_t) t * tile_stride, load_bytes, slot_bar[s], kPolicyEvictFirst);
			}
		} else {
			uint8_t *dst = smem + (slot_buf[s] - sb);
			if (packed_in)
				load_tile_edge_packed(a, t, dst);
			else
				load_tile_edge(a, t, dst);
			__syncwarp(); // every lane's part of the tile is written before lane 0 releases it
			if (lane == 0)
				mbar_arrive(slot_bar[s]);
		}
	};
	// claim the next tile of the span for ring slot s and start loading it
	// (lane 0 draws the ticket; the other lanes' adds go to words of the warp's match list, which holds nothing
	// between two tiles)
	const uint32_t tick = lane == 0 ? s_next : lst + 4 * lane;
	auto refill = [&](uint32_t s) { issue(s, __shfl_sync(kFull, atoms_add1(tick), 0)); };
#pragma unroll
	for (uint32_t s = 0; s < stages; s++) {
		slot_phase[s] = 0;
		slot_buf[s] = wbase + s * kBufBytes;
		slot_bar[s] = bars + 8 * s;
		issue(s, s * W + warp);
	}

	// tables: global -> shared with TMA bulk copies, overlapped with the first text tiles
	if (threadIdx.x == 0) {
		if (tab_bytes) {
			mbar_expect_tx(tab_bar, tab_bytes);
#pragma unroll 1
			for (uint32_t off = 0; off < front_smem; off += 16384)
				tma_bulk_g2s(s_front + off, a.front + off, min(16384u, front_smem - off), tab_bar, kPolicyEvictLast);
#pragma unroll 1
			for (uint32_t off = 0; off < rm_bytes; off += 16384)
				tma_bulk_g2s(s_rmask + off, a.rmask + off, min(16384u, rm_bytes - off), tab_bar, kPolicyEvictLast);
#pragma unroll 1
			for (uint32_t off = 0; off < f2_bytes; off += 16384)
				tma_bulk_g2s(s_f2 + off, reinterpret_cast<const uint8_t *>(a.filter2) + off, min(16384u, f2_bytes - off), tab_bar,
						kPolicyEvictLast);
		}
		trace_mark(a, 1);
	}
	if (blockIdx.x == 0 && threadIdx.x == THREADS - 1) {
		// the successor's counters (see Work): zero in L2 before the successor may start (it may once every CTA
		// of ours has passed the trigger below); done by a lane that has nothing to issue, under the first copies
		Work *nxt = &a.ctl->work[(a.epoch + 1u) % kWorkRing];
		__stcg(&nxt->arrive, 0ull);
		__stcg(&nxt->cursor, 0ull);
		__stcg(&nxt->bad_text, 0u);
		__threadfence();
	}
	// per-tile counts: a tile without a match never writes its word
	if (a.want_positions)
		for (uint32_t i = threadIdx.x; i < min(n_b, a.cnt_cap); i += THREADS)
			s_cnt_p[i] = 0;
	__syncthreads(); // the ticket, the table barrier and the CTA counters are set up
	if (a.pdl)
		pdl_trigger(); // overlap mode: the next scan of the stream may take over SMs as our CTAs retire
	if (threadIdx.x == 0) {
		trace_mark(a, 2);
		trace_mark(a, 9);
		trace_mark(a, 10); // 2 -> 9 -> 10: what a stamp itself costs
	}

	uint32_t badacc = 0;
	Work *wk = &a.ctl->work[a.epoch % kWorkRing];
	Emitter em;
	em.a = &a;
	em.wk = wk;
	em.s_cnt = s_cnt;
	em.idx = 0;
	em.tile = 0;
	em.warp_count = 0;
	em.blk_ptr = 0;
	em.blk_left = 0;
	em.log_s = wlog;
	em.n_log = em.lost = 0;
	em.tile_slots = nullptr;
	em.tile_word = 0;
	Front fr;
	const TabRef f2{s_f2, reinterpret_cast<const uint8_t *>(a.filter2), a.prm.f2_in_smem != 0};
	fr.init(s_front, TabRef{s_rmask, a.rmask, a.prm.r_in_smem != 0}, a);
	// candidates of a tile are checked by the lanes that found them while they are few (<= lane_local), by the
	// whole warp from a compacted list while they fit it (<= kListCap), and lane by lane again beyond that
	const uint32_t lane_local = a.lane_local;

	bool tab_ready = tab_bytes == 0; // the tables are first needed by walk(): the first tile is loaded and packed under their copy
	for (;;) {
		// the oldest slot is slot 0 (the bytes path rotates its two slots at the end of the iteration)
		const uint32_t idx = slot_idx[0];
		if (idx >= n_b)
			break; // the span is exhausted (claims are handed out in ascending order)
		const uint64_t tile = cta_lo + idx;
		mbar_wait(slot_bar[0], slot_phase[0]); // every lane sees the phase flip itself: the tile is visible to it
		slot_phase[0] ^= 1u;

		const uint32_t buf = slot_buf[0];
		fr.load(a, buf, pk, badacc);
		if constexpr (kPacked) { // the tile now lives in registers (+ pk): refill the slot while we walk
			__syncwarp();
			refill(0);
		}
		if (!tab_ready) {
			if (lane == 0)
				trace_mark(a, 16 + warp); // first tile in registers
			mbar_wait(tab_bar, 0);
			tab_ready = true;
			if (threadIdx.x == 0)
				trace_mark(a, 3);
		}
		fr.walk(a);

		em.tile = tile;
		em.idx = idx;
		// overlap mode: the previous scan of the stream may still be ordering its matches -- in ITS copy of
		// the scratch arrays and counters; what we write until we arrive goes to ours
		const uint64_t tile_start = tile * (uint64_t) kTile;
		uint32_t total = 0;
		if constexpr (EXACT) {
			const uint64_t end_lo = a.report_lo; // first end position this scan reports (>= data_lo + m_min - 1)
			const bool inner = tile_start >= end_lo && tile_start + kTile <= a.data_hi;
			if (!inner) { // first / last tiles: drop ends outside [end_lo, data_hi)
				const uint64_t cs = tile_start + (uint64_t) lane * kLane;
				const uint32_t lo_s = cs >= end_lo ? 0u : (uint32_t) min((uint64_t) kLane, end_lo - cs);
				const uint32_t hi_s = cs >= a.data_hi ? 0u : (uint32_t) min((uint64_t) kLane, a.data_hi - cs);
				fr.mask_range(lo_s, hi_s);
			}
			const uint32_t cnt = fr.count();
			const uint32_t b1 = __ballot_sync(kFull, cnt != 0);
			if (b1) {
				// rank of a lane's first match = matches of the lanes in front of it.  The usual tile of a sparse text has
				// at most two matches per lane: two ballots give the prefix (cnt = [cnt >= 1] + [cnt >= 2]) and every lane
				// stages its own matches; dense tiles take the shuffle scan and leave through the shared list
				// (full-warp stores).
				const uint32_t b2 = __ballot_sync(kFull, cnt > 1);
				const bool sparse = __ballot_sync(kFull, cnt > 2) == 0;
				uint32_t k;
				if (sparse) {
					uint32_t lt;
					asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
					k = __popc(b1 & lt) + __popc(b2 & lt);
					total = __popc(b1) + __popc(b2);
				} else {
					const uint32_t incl = warp_incl_scan(cnt);
					total = __shfl_sync(kFull, incl, 31);
					k = incl - cnt;
				}
				if (a.want_positions) {
					em.reserve(total);
					if (!sparse && total <= kListCap) { // positions -> shared list, then full-warp stores
#pragma unroll
						for (int g = 0; g < Front::kWords; g++)
							for (uint32_t w = fr.hw[g]; w; w &= w - 1)
								sts16(lst + 2 * k++, lane * kLane + Front::sym_of(g, __ffs(w) - 1));
						__syncwarp();
						for (uint32_t i = lane; i < total; i += 32)
							em.put(i, lds_u16(lst + 2 * i));
						__syncwarp(); // the list is free again
					} else {
#pragma unroll
						for (int g = 0; g < Front::kWords; g++)
							for (uint32_t w = fr.hw[g]; w; w &= w - 1)
								em.put(k++, lane * kLane + Front::sym_of(g, __ffs(w) - 1));
					}
				}
			}
		} else {
			// (2-bit path: the copy of the tile in pk is complete, the warp met after load())
			// candidates -> offset mask -> stage-2 bitmap -> buckets; survivors become bits of mw (chunk-relative
			// end positions) of the lane that owns the chunk
			uint32_t mw0 = 0, mw1 = 0, mw2 = 0, mw3 = 0, multi = 0;
			auto set_mw = [&](uint32_t p) {
				const uint32_t bit = 1u << (p & 31);
				if (p < 32)
					mw0 |= bit;
				else if (p < 64)
					mw1 |= bit;
				else if (p < 96)
					mw2 |= bit;
				else
					mw3 |= bit;
			};
			// how many distinct patterns end at tile symbol tp (0: none)
			auto check_end = [&](uint32_t tp) -> uint32_t {
				const uint32_t key = Front::key_at(a, buf, pk, tp);
				const uint32_t i2 = (uint32_t) (key * a.prm.f2_mult) >> a.prm.f2_sh;
				if ((f2.u32(i2 >> 5) >> (i2 & 31)) & 1u)
					return verify_window(a, key, tile_start + tp);
				return 0u;
			};
			const uint32_t ncand = fr.count();
			const uint32_t tot_c = __reduce_add_sync(kFull, ncand);
			if (tot_c == 0) {
				// the usual tile of a selective filter
			} else if (tot_c > lane_local && tot_c <= kListCap) {
				// candidates unevenly spread over the lanes: compact them into the warp's list and let
				// lane i check candidate i (one pass instead of max-per-lane divergent iterations)
				uint32_t k = warp_incl_scan(ncand) - ncand;
#pragma unroll
				for (int g = 0; g < Front::kWords; g++)
					for (uint32_t w = fr.hw[g]; w; w &= w - 1)
						sts16(lst + 2 * k++, lane * kLane + Front::sym_of(g, __ffs(w) - 1));
				__syncwarp();
				for (uint32_t base = 0; base < tot_c; base += 32) {
					const uint32_t i = base + lane;
					uint32_t cpos = 0, sv = 0, mu = 0;
					if (i < tot_c) {
						cpos = lds_u16(lst + 2 * i);
						uint32_t rm = fr.probe_mask(a, buf, pk, cpos);
						while (rm) {
							const uint32_t r = (uint32_t) (__ffs(rm) - 1);
							rm &= rm - 1;
							const uint32_t mult = check_end(cpos + r);
							if (mult) {
								sv |= 1u << r;
								mu |= mult > 1;
							}
						}
					}
					// survivors (rare) go back to the lane whose chunk they end in
					uint32_t any = __ballot_sync(kFull, sv != 0);
					while (any) {
						const int src = __ffs(any) - 1;
						any &= any - 1;
						const uint32_t cp = __shfl_sync(kFull, cpos, src);
						uint32_t sbits = __shfl_sync(kFull, sv, src);
						const uint32_t m2 = __shfl_sync(kFull, mu, src);
						const uint32_t owner = cp / kLane;
						if (lane == owner) {
							const uint32_t c = cp - owner * kLane;
							for (; sbits; sbits &= sbits - 1)
								set_mw(c + (uint32_t) (__ffs(sbits) - 1));
							multi |= m2;
						}
					}
				}
				__syncwarp(); // the list is reused for the match positions below
			} else {
				// few candidates (a couple per lane at most) or very many: every lane checks its own
#pragma unroll
				for (int g = 0; g < Front::kWords; g++) {
					uint32_t w = fr.hw[g];
					while (w) {
						const int b = __ffs(w) - 1;
						w &= w - 1;
						const uint32_t c = Front::sym_of(g, b);
						uint32_t rm = fr.probe_mask(a, buf, pk, lane * kLane + c);
						while (rm) {
							const uint32_t p = c + (uint32_t) (__ffs(rm) - 1);
							rm &= rm - 1;
							const uint32_t mult = check_end(lane * kLane + p);
							if (mult) {
								set_mw(p);
								multi |= mult > 1;
							}
						}
					}
				}
			}
			// how many entries this lane emits (a position with several patterns ending there
			// emits one entry per pattern: mixed-length sets only; recounted on demand)
			auto mult_at = [&](uint32_t p) {
				const uint32_t key = Front::key_at(a, buf, pk, lane * kLane + p);
				return verify_window(a, key, tile_start + lane * kLane + p);
			};
			const uint32_t mw[4] = {mw0, mw1, mw2, mw3};
			uint32_t cnt = 0;
			if (tot_c) {
				cnt = __popc(mw0) + __popc(mw1) + __popc(mw2) + __popc(mw3);
				if (multi) {
					cnt = 0;
#pragma unroll
					for (int g = 0; g < 4; g++)
						for (uint32_t w = mw[g]; w; w &= w - 1)
							cnt += mult_at(32 * g + __ffs(w) - 1);
				}
			}
			if (tot_c && __any_sync(kFull, cnt != 0)) {
				const uint32_t incl = warp_incl_scan(cnt);
				total = __shfl_sync(kFull, incl, 31);
				if (a.want_positions) {
					em.reserve(total);
					uint32_t k = incl - cnt;
					const bool listed = total <= kListCap;
#pragma unroll
					for (int g = 0; g < 4; g++)
						for (uint32_t w = mw[g]; w; w &= w - 1) {
							const uint32_t p = 32 * g + __ffs(w) - 1;
							const uint32_t reps = multi ? mult_at(p) : 1u;
							for (uint32_t i = 0; i < reps; i++, k++) {
								if (listed)
									sts16(lst + 2 * k, lane * kLane + p);
								else
									em.put(k, lane * kLane + p);
							}
						}
					if (listed) {
						__syncwarp();
						for (uint32_t i = lane; i < total; i += 32)
							em.put(i, lds_u16(lst + 2 * i));
					}
				}
			}
			__syncwarp(); // every lane is done with this slot (bytes path), pk and the list before they are rewritten
		}
		em.end_tile(total);
		if constexpr (!kPacked) { // refill the slot just walked, then the other slot (loaded meanwhile) becomes slot 0
			if constexpr (EXACT)
				__syncwarp();
			refill(0);
			if constexpr (stages > 1) {
			const uint32_t ti = slot_idx[0], tp = slot_phase[0], tb = slot_buf[0], tr = slot_bar[0];
			slot_idx[0] = slot_idx[stages - 1], slot_phase[0] = slot_phase[stages - 1];
			slot_buf[0] = slot_buf[stages - 1], slot_bar[0] = slot_bar[stages - 1];
			slot_idx[stages - 1] = ti, slot_phase[stages - 1] = tp;
			slot_buf[stages - 1] = tb, slot_bar[stages - 1] = tr;
			}
		}
	}
	if (a.trace && lane == 0)
		trace_mark(a, 48 + warp);
	if (!tab_ready)
		mbar_wait(tab_bar, 0); // a warp without a tile: the CTA must not retire under its own table copy
	em.finish();

	// ---- per-CTA totals (shared memory)
	if (lane == 0 && em.warp_count)
		atomicAdd(s_count, em.warp_count);
	if (em.lost && lane == 0)
		atomicOr(s_bad, 4u);
	if constexpr (kPacked) {
		badacc &= 0xFCFCFCFCu;
		if (__any_sync(kFull, badacc != 0) && lane == 0)
			atomicOr(s_bad, 1u);
	}
	__syncthreads();
	if (threadIdx.x == 0)
		trace_mark(a, 4);

	// ---- order, part 1: exclusive prefix of the per-tile counts inside this CTA's span (in place: shared memory,
	// and global memory for the tiles of a span too long for it)
	const uint32_t G = gridDim.x;
	const unsigned long long my_total = *s_count;
	if (a.want_positions && my_total) {
		if (warp == 0) { // one warp, 8 tiles per lane and pass: no block-wide barrier inside
			uint32_t carry = 0;
			for (uint32_t base = 0; base < n_b; base += 256) {
				uint32_t v[8], sum = 0;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const uint32_t i = base + lane * 8 + k;
					v[k] = 0;
					if (i < n_b)
						v[k] = i < a.cnt_cap ? s_cnt_p[i] : __ldcg(a.tile_count + cta_lo + i);
					sum += v[k];
				}
				const uint32_t incl = warp_incl_scan(sum);
				uint32_t run = carry + incl - sum;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const uint32_t i = base + lane * 8 + k;
					if (i < n_b) {
						if (i < a.cnt_cap)
							s_cnt_p[i] = run;
						else
							a.tile_count[cta_lo + i] = run;
					}
					run += v[k];
				}
				carry += __shfl_sync(kFull, incl, 31);
			}
		}
	}

	// ---- publish: the CTA's total goes out first (the spans behind us wait for nothing else), then ONE atomic is
	// count, arrival and exit ticket: the last CTA to arrive writes the result block.  Nobody waits for the grid.
	bool publisher = false;
	// the previous scan's counts have had a whole scan to arrive (the last one is collected by acwm_fetch_global_count)
	auto collect_peers = [&]() {
		if (a.xepoch < 2)
			return;
		unsigned long long sum;
		if (collect_mailbox(a.peers[a.rank], a.world, a.xepoch - 1, sum)) {
			a.ctl->result.global_count = sum;
			a.ctl->result.global_epoch = a.xepoch - 1;
		} else
			a.ctl->result.exchange_failed = a.xepoch - 1;
	};
	const unsigned long long tag = (unsigned long long) ((a.epoch + 1u) & 0xffffffu) << kTotalShift;
	if (threadIdx.x == 0) {
		if (a.want_positions && !((a.tune & kTuneFaultHideTotal) && blockIdx.x == 0))
			__stcg(a.cta_total + blockIdx.x, tag | my_total); // one 8-byte word: tag and value arrive together
		trace_mark(a, 5);
		// overlap mode: from here on we touch what the previous scan of the stream publishes (result block,
		// positions) -- and no CTA leaves before that scan is complete (the invariant behind Work)
		if (a.pdl)
			pdl_wait();
		Result *res = &a.ctl->result;
		unsigned long long old_count = 0, old_written = 0;
		unsigned int old_bad = 0, old_ovf = 0;
		if (a.append) { // read before anybody can publish (nobody publishes before we have arrived)
			old_count = __ldcg(&res->count);
			old_written = __ldcg(&res->written);
			old_bad = __ldcg(&res->bad_text);
			old_ovf = __ldcg(&res->overflow);
		}
		s_misc[1] = old_written;
		if (*s_bad)
			atomicOr(&wk->bad_text, *s_bad);
		__threadfence();
		const unsigned long long before = atomicAdd(&wk->arrive, (1ull << kArriveShift) + my_total);
		trace_mark(a, 6);
		if ((before >> kArriveShift) == G - 1) { // everybody has arrived: totals are final
			__threadfence();
			const unsigned long long cnt = (before & kArriveCountMask) + my_total;
			const unsigned long long cur = __ldcg(&wk->cursor);
			const unsigned int bad = __ldcg(&wk->bad_text);
			unsigned long long r_written = old_written;
			unsigned int r_ovf = old_ovf | ((bad >> 2) & 1u);
			if (a.want_positions) {
				if (old_written + cnt > a.cap || cur > a.stage_cap) {
					r_ovf = 1;
					r_written = min(old_written + cnt, (unsigned long long) a.cap);
				} else
					r_written += cnt;
			}
			res->count = old_count + cnt;
			res->written = r_written;
			res->bad_text = old_bad | (bad & 3u);
			res->overflow = r_ovf;
			if (a.world > 1) { // hand this launch's count to every rank (ours included)
				const unsigned long long tagged = ((unsigned long long) (a.xepoch & 0xffffu) << kMailShift) | cnt;
				const uint32_t slot = (a.xepoch & (kPeerRing - 1)) * a.world + a.rank;
				for (uint32_t r = 0; r < a.world; r++)
					st_relaxed_sys_u64(a.peers[r] + slot, tagged);
				publisher = true;
			} else
				res->global_count = old_count + cnt;
		}
	}
	if (!a.want_positions || !my_total) { // nothing of ours to place
		if (threadIdx.x == 0)
			trace_mark(a, 8);
		if (publisher)
			collect_peers();
		return;
	}

	// ---- order, part 2: our span starts where the spans in front of us end (their totals: a look-back over
	// at most G - 1 tagged words, normally all there already -- CTAs start in span order)
	uint32_t *s_stall = reinterpret_cast<uint32_t *>(smem + 48);
	if (warp == 1) { // not warp 0: its lane 0 may still be waiting for the previous scan
		unsigned long long sum = 0;
		bool stalled = false;
		const unsigned long long t0 = globaltimer_ns();
		for (uint32_t j = lane; j < blockIdx.x; j += 32) {
			unsigned long long v;
			uint32_t spins = 0;
			while (((v = ld_acquire_u64(a.cta_total + j)) & ~kTotalMask) != tag) {
				if ((++spins & 1023u) == 0
						&& globaltimer_ns() - t0 > ((a.tune & kTuneFaultHideTotal) ? 20ull * 1000 * 1000 : kLookbackTimeoutNs)) {
					stalled = true;
					break;
				}
			}
			if (stalled)
				break;
			sum += v & kTotalMask;
		}
#pragma unroll
		for (int d = 16; d; d >>= 1)
			sum += __shfl_xor_sync(kFull, sum, d);
		stalled = __any_sync(kFull, stalled);
		if (lane == 0) {
			s_misc[0] = sum;
			*s_stall = stalled ? 1u : 0u;
			if (stalled) // positions of this scan are incomplete: acwm_fetch reports it
				atomicMax(&a.ctl->result.order_failed, a.epoch + 1u);
		}
	}
	__syncthreads(); // also: thread 0 has passed its wait for the previous scan, s_misc[1] is set
	if (threadIdx.x == 0)
		trace_mark(a, 7);
	if (!*s_stall) {
		const unsigned long long span_base = s_misc[1] + s_misc[0];
		// every warp moves what it staged: entry -> positions[span base + tile offset + rank]
		constexpr int kBatch = 4;
		for (uint32_t d = 0; d < em.n_log; d++) {
			const unsigned long long desc = lds64(em.log_s + 8u * d);
			const unsigned long long ptr = desc & ((1ull << 40) - 1);
			const uint32_t len = (uint32_t) (desc >> 40);
			for (uint32_t i0 = lane; i0 < len; i0 += 32 * kBatch) {
				uint64_t e[kBatch];
#pragma unroll
				for (int u = 0; u < kBatch; u++) {
					const uint32_t i = i0 + 32 * u;
					e[u] = (i < len && ptr + i < a.stage_cap) ? __ldcg(a.staging + ptr + i) : ~0ull;
				}
#pragma unroll
				for (int u = 0; u < kBatch; u++) {
					if (e[u] == ~0ull)
						continue;
					const uint64_t tile = e[u] >> (kRankBits + kPosBits);
					const uint32_t rank = (uint32_t) (e[u] >> kPosBits) & ((1u << kRankBits) - 1);
					const uint32_t pos = (uint32_t) e[u] & ((1u << kPosBits) - 1);
					const uint64_t ti = tile - cta_lo;
					const uint32_t off = ti < a.cnt_cap ? s_cnt_p[ti] : __ldcg(a.tile_count + tile);
					const unsigned long long at = span_base + off + rank;
					if (at < a.cap)
						a.positions[at] = tile * kTile + pos - a.data_lo;
				}
			}
		}
	}
	if (threadIdx.x == 0)
		trace_mark(a, 8);
	if (publisher)
		collect_peers();
}

// ------------------------------------------------------------ launch helper
template <class Front, bool EXACT, int THREADS, int MINB, int STAGES = Front::kPacked ? 1 : 2>
static cudaError_t launch_shape(const ScanArgs &a, uint32_t smem, uint32_t grid, cudaStream_t st) {
	auto kern = scan_kernel<Front, EXACT, THREADS, MINB, STAGES>;
	static std::atomic<bool> attr_set[64]; // per device; shard threads of one process launch concurrently
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
		e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxSmem);
		if (e != cudaSuccess)
			return e;
		if (dev >= 0 && dev < 64)
			attr_set[dev].store(true, std::memory_order_release);
	}
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	// Plain scans are launched cooperatively (the runtime checks that the grid is resident at once: the span
	// look-back then never waits for a CTA that has not started); overlap mode trades that for a programmatic
	// dependent launch -- used only for grids of one CTA per SM, whose CTAs all find room as the CTAs of the scan
	// two launches back retire (the two attributes together serialise, profiles/README.md session i).  The
	// cooperative attribute costs nothing measurable (stand-alone c2 scan 38.1 us with it, 37.6 without: session r02u).
	cudaLaunchAttribute attr[1];
	if (a.pdl) {
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
	} else {
		attr[0].id = cudaLaunchAttributeCooperative;
		attr[0].val.cooperative = 1;
	}
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, a);
}

// 2-bit path: 1024 / 768 threads own their SM; 512 / 384 threads are compiled for two CTAs per SM (64 / 80
// registers) and serve both shapes.  The bytes path needs two raw slots per warp: at most 16 warps fit.
template <class Front, bool EXACT>
static cudaError_t launch_front(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st) {
	if constexpr (Front::kPacked) {
		switch (threads) {
		case 1024: return launch_shape<Front, EXACT, 1024, 1>(a, smem, grid, st);
		case 768: return launch_shape<Front, EXACT, 768, 1>(a, smem, grid, st);
		case 512: return launch_shape<Front, EXACT, 512, 2>(a, smem, grid, st);
		case 384: return launch_shape<Front, EXACT, 384, 2>(a, smem, grid, st);
		case 256: return launch_shape<Front, EXACT, 256, 2>(a, smem, grid, st);
		case 128: return launch_shape<Front, EXACT, 128, 2>(a, smem, grid, st);
		default: return cudaErrorInvalidValue;
		}
	} else if (a.stages == 1) { // one raw slot per warp, more warps
		switch (threads) {
		case 768: return launch_shape<Front, EXACT, 768, 1, 1>(a, smem, grid, st);
		case 640: return launch_shape<Front, EXACT, 640, 1, 1>(a, smem, grid, st);
		case 512: return launch_shape<Front, EXACT, 512, 1, 1>(a, smem, grid, st);
		default: return cudaErrorInvalidValue;
		}
	} else {
		switch (threads) {
		case 512: return launch_shape<Front, EXACT, 512, 1>(a, smem, grid, st);
		case 384: return launch_shape<Front, EXACT, 384, 1>(a, smem, grid, st);
		case 256: return launch_shape<Front, EXACT, 256, 1>(a, smem, grid, st);
		case 128: return launch_shape<Front, EXACT, 128, 1>(a, smem, grid, st);
		default: return cudaErrorInvalidValue;
		}
	}
}

} // namespace acwm
