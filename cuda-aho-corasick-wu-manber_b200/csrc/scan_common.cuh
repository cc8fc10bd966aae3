// Device-side pieces shared by the 2-bit and the bytes front ends (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/acwm.h"
#include "geometry.hpp"

namespace acwm {

// What a finished scan leaves behind for acwm_fetch / the multi-GPU layer.
struct Result {
	unsigned long long count;    // |M| (cumulative over the launches of one search)
	unsigned long long written;  // positions placed in the output (cumulative)
	unsigned int bad_text;       // OR of (byte & 0xFC) over the text, 2-bit path
	unsigned int overflow;       // more matches than position capacity
	unsigned long long global_count; // sum of `count` over the ranks of the peer exchange (== count without peers)
	unsigned int global_epoch;   // exchange epoch global_count belongs to
	unsigned int order_failed;   // launch number + 1 of the last scan whose position ordering gave up (never in practice)
	unsigned int exchange_failed; // exchange epoch whose counts did not all arrive in time (a peer skipped a scan or died)
	unsigned int pad;
};

// Working counters of the launch in flight.  kWorkRing copies: launch k uses work[k % 4] and clears
// work[(k + 1) % 4] when it STARTS (before it lets its successor in).  In overlap mode consecutive launches run
// at the same time, at most THREE of them: every launch of the chain puts one CTA on every SM (grid = #SMs) and
// an SM holds at most two CTAs of this kernel (shared memory: a CTA takes more than a third of the SM), so CTAs
// of launch k + 2 find room only as CTAs of launch k exit -- and those exit only after launch k - 1 has
// completed (griddepcontrol.wait before the arrival).  Hence when launch k starts and clears the copy of launch
// k + 1 = the copy of launch k - 3, that launch is over, and the successor finds its copy zeroed.  No memset node
// between scans.  The scratch arrays the scan phase writes (staging, per-tile counts, per-CTA totals)
// exist kScratchRing = 3 times, selected by launch number, for the same reason.
struct Work {
	unsigned long long arrive;   // [CTAs arrived : 16 | matches : 48] -- one atomic per CTA is count, grid barrier, exit
	                             // ticket and (by arrival order) the CTA's role in the ordering epilogue
	unsigned long long cursor;   // staging slots handed out
	unsigned int bad_text;
	unsigned int pad[3];
};
constexpr unsigned kArriveShift = 48;
constexpr unsigned long long kArriveCountMask = (1ull << kArriveShift) - 1;
constexpr unsigned kMailShift = 48;   // mailbox word of the count exchange: [epoch tag : 16 | count : 48]
// Per-CTA totals carry the launch number, so a stale value is never mistaken for this launch's: [tag : 24 | matches : 40]
constexpr unsigned kTotalShift = 40;
constexpr unsigned long long kTotalMask = (1ull << kTotalShift) - 1;
constexpr unsigned long long kLookbackTimeoutNs = 2000ull * 1000 * 1000; // a predecessor that never shows up: report, do not hang
// Work.bad_text bits: 1 = text byte >= 4 on the 2-bit path, 4 = a warp's reservation log overflowed (reported as overflow)
// ScanArgs.tune bits
constexpr uint32_t kTuneCoopVerify = 1u;       // candidates of a tile may be checked cooperatively (compacted list, lane i checks candidate i)
constexpr uint32_t kTuneFaultHideTotal = 2u;    // fault injection (tests only): CTA 0 never publishes its span total and the look-back
                                               // gives up after 20 ms instead of 2 s -- the path a stuck predecessor would take
constexpr uint32_t kTuneLaneLocalDefault = 3;  // ... when the warp has more than this many (bits 8..15 of tune); fewer: every lane checks its own

constexpr uint32_t kWorkRing = 4, kScratchRing = 3;

struct Control {
	Result result;
	Work work[kWorkRing];
};

struct ScanArgs {
	const uint8_t *text16;       // 16-byte aligned base of the "virtual" text
	uint64_t data_lo, data_hi;   // real text occupies virtual [data_lo, data_hi), data_lo < 16
	uint64_t report_lo;          // matches ending before this virtual position are not reported
	uint64_t tile_lo, tile_hi;   // warp tiles [tile_lo, tile_hi) of the virtual text are scanned by this launch
	uint64_t tiles_per_cta;      // CTA b owns tiles [tile_lo + b*tiles_per_cta, +tiles_per_cta) (clipped to tile_hi)
	const uint8_t *front;        // front-end table (global copy)
	uint32_t front_bytes;
	uint32_t front_in_smem;
	const uint8_t *rmask;        // WM offset masks (global copy), r_entries * r_entry_bytes
	const uint32_t *filter2;     // stage-2 bitmap (global copy)
	const uint32_t *bucket_start;
	const acwm_ventry *entries;
	const uint8_t *patterns;
	const uint32_t *vdfa;        // filtered AC: the automaton that decides a candidate window (global memory)
	acwm_scan_params prm;
	Control *ctl;
	uint64_t *staging;           // [tile:28 | rank:22 | pos:14]
	uint64_t *positions;         // sorted output
	uint64_t cap;                // capacity of positions (entries)
	uint64_t stage_cap;          // capacity of staging: cap + one reservation block per warp of the grid
	uint32_t *tile_count;        // matches per tile -> (in the epilogue) exclusive prefix within the owning CTA
	unsigned long long *cta_total; // matches per CTA
	uint32_t stages;             // ring depth of the per-warp tile pipeline
	uint32_t s_rmask, s_f2, s_cnt; // shared addresses of the offset masks, the stage-2 bitmap and the per-tile counts (api.cu
	                             // computes the layout of scan_kernel.cuh once per launch; the kernel checks it)
	uint32_t cnt_cap;            // per-tile counts of the first cnt_cap tiles of a span live in shared memory
	uint32_t epoch;              // launch number of this matcher: selects the Work copy
	// multi-GPU count exchange over NVLink peer memory: every rank's mailbox is uint64[kPeerRing][world]
	uint32_t world, rank, xepoch;
	unsigned long long *peers[kMaxPeers];
	int want_positions;
	int append;                  // 1: add to ctl->result instead of replacing it (chunked host text)
	int pdl;                     // 1: launched as a programmatic dependent launch (consecutive scans may overlap)
	uint32_t tune;               // kTune* bits
	uint32_t lane_local;         // tiles with at most this many candidates: every lane checks its own (more: cooperatively, while they fit the list)
	uint32_t packed_in;          // 1: text16 holds the text already packed 4 symbols per byte (host-side packer of
	                             // acwm_search_host, alphabet <= 4): data_lo = 0, buffer zero-padded to 16 bytes
	unsigned long long *trace;   // instrumentation (acwm_set_trace): kTraceWords of %globaltimer stamps per CTA, else NULL
};

constexpr unsigned kFull = 0xffffffffu;

// The mailbox word carries its own tag, so relaxed system-scope accesses are enough.
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
constexpr unsigned long long kExchangeTimeoutNs = 2000ull * 1000 * 1000; // a peer that never publishes: report, do not hang
// Sum of the counts every rank left in this GPU's mailbox for exchange epoch `x`: spins until all are in, gives up
// (false) after kExchangeTimeoutNs -- a peer that skipped a scan, failed before its launch or died must not hang this
// GPU's stream for good.
__device__ __forceinline__ bool collect_mailbox(const unsigned long long *box, uint32_t world, uint32_t x,
		unsigned long long &sum) {
	sum = 0;
	const unsigned long long *slot = box + (x & (kPeerRing - 1)) * world;
	const unsigned long long t0 = globaltimer_ns();
	for (uint32_t r = 0; r < world; r++) {
		unsigned long long v;
		uint32_t spins = 0;
		while (((v = ld_relaxed_sys_u64(slot + r)) >> kMailShift) != (x & 0xffffu)) {
			__nanosleep(32);
			if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > kExchangeTimeoutNs)
				return false;
		}
		sum += v & ((1ull << kMailShift) - 1);
	}
	return true;
}

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

// ------------------------------------------------------------ shared memory by 32-bit shared-space address
// Everything the scan loop touches in shared memory is addressed by its 32-bit shared-space address (computed once
// from the dynamic shared-memory base): plain LDS / STS / SYNCS with register + immediate addressing, no generic
// pointers (whose conversion the compiler rematerialises at every use).  The asm statements are volatile: the slots
// they read are rewritten by TMA copies and by other lanes between iterations.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
constexpr uint32_t kDynSmemBase = 0x400; // shared address of the dynamic shared memory (scan_kernel checks it)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
	uint4 r;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
	return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds64(uint32_t a) {
	unsigned long long v;
	asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
	return v;
}
// One ticket per warp without a divergent region: EVERY lane adds 1 at its own address -- lane 0 at the ticket counter,
// the others at scratch words of their own -- so the compiler sees a plain full-warp ATOMS (an atomic under
// `if (lane == 0)` is rewritten into a vote / popc / elect sequence of a dozen instructions).
__device__ __forceinline__ uint32_t atoms_add1(uint32_t a) {
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(a) : "memory");
	return old;
}

// ------------------------------------------------------------ mbarrier / TMA bulk copy
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
			"{\n"
			".reg .pred p;\n"
			"WAIT_%=:\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
			"@p bra DONE_%=;\n"
			"bra WAIT_%=;\n"
			"DONE_%=:\n"
			"}\n" ::"r"(bar),
			"r"(parity)
			: "memory");
}
// L2 policies of createpolicy.fractional.L2::evict_first / evict_last (fraction 1.0), as the constants they encode to:
// the text stream is read once and evicted first (keeps L2-resident tables in place), tables are kept
constexpr unsigned long long kPolicyEvictFirst = 0x12F0000000000000ull, kPolicyEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, unsigned long long pol) {
	asm volatile(
			"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
			"l"(src), "r"(bytes), "r"(bar), "l"(pol)
			: "memory");
}

// A read-only table that lives in shared memory or (too large for it) in global memory: which one is a launch
// parameter, so the access is a uniform branch between LDS and LDG instead of a generic load (64-bit address
// arithmetic + address-space resolution on every candidate).
struct TabRef {
	uint32_t s;          // shared-memory address (valid when in_smem)
	const uint8_t *g;    // global copy
	bool in_smem;
	__device__ __forceinline__ uint32_t u32(uint32_t word) const {
		uint32_t v;
		if (in_smem)
			asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(s + 4u * word));
		else
			v = __ldg(reinterpret_cast<const uint32_t *>(g) + word);
		return v;
	}
	__device__ __forceinline__ uint32_t u16(uint32_t i) const {
		uint32_t v;
		if (in_smem)
			asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(s + 2u * i));
		else
			v = __ldg(reinterpret_cast<const uint16_t *>(g) + i);
		return v;
	}
	__device__ __forceinline__ uint32_t u8(uint32_t i) const {
		uint32_t v;
		if (in_smem)
			asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(s + i));
		else
			v = __ldg(g + i);
		return v;
	}
};

// ------------------------------------------------------------ arrival / look-back loads
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

// instrumentation: phase time stamps of every CTA (scripts/trace.py turns them into a timeline)
constexpr uint32_t kTraceWords = 16 + 3 * 32; // [16 CTA phases][32 warps: first tile ready][32: scan loop done][32: tiles scanned]
__device__ __forceinline__ void trace_mark(const ScanArgs &a, uint32_t slot) {
	if (a.trace)
		a.trace[(size_t) blockIdx.x * kTraceWords + slot] = globaltimer_ns();
}

// programmatic dependent launch: wait for the previous kernel of the stream / let the next one become resident
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Warp-level staging of the matches of one tile.  A warp reserves staging slots from the launch's cursor in
// blocks that double in size (one global atomic per block, not per tile) and logs its blocks in shared
// memory: when the scan is over the warp itself moves its matches to their sorted places.  The entries of one
// tile are contiguous: a tile that does not fit what is left of the current block leaves that tail unused (the
// block's log entry shrinks) and starts a new block.
struct Emitter {
	const ScanArgs *a;
	Work *wk;
	uint32_t s_cnt;              // shared address of the per-tile counts (first a->cnt_cap tiles of the span; zeroed in the prologue)
	uint32_t log_s;              // shared address of this warp's reservations: [slots : 24 | first slot : 40]
	uint64_t tile;
	uint32_t idx;                // span-relative index of the tile
	unsigned long long warp_count;
	unsigned long long blk_ptr;  // next free slot of the current block
	uint32_t blk_left;
	uint32_t n_log, lost;        // reservations logged / more reservations than the log holds (reported as overflow)
	uint64_t *tile_slots;        // where entry 0 of the current tile goes (nullptr: the staging array is full)
	uint64_t tile_word;          // the current tile's number in place: tile << (kRankBits + kPosBits)

	// warp-uniform: make room for `total` entries of the current tile
	__device__ __forceinline__ void reserve(uint32_t total) {
		if (total > blk_left) {
			const uint32_t grab = max(total, kStageBlock << min(n_log, kMaxGrabLog2 - kStageBlockLog2));
			unsigned long long p = 0;
			if (lane_id() == 0) {
				if (n_log && !lost) // the unused tail of the block we leave
					sts64(log_s + 8u * (n_log - 1), lds64(log_s + 8u * (n_log - 1)) - ((unsigned long long) blk_left << 40));
				p = atomicAdd(&wk->cursor, (unsigned long long) grab);
				if (n_log < kLogCap)
					sts64(log_s + 8u * n_log, ((unsigned long long) grab << 40) | p);
			}
			if (n_log < kLogCap)
				n_log++;
			else
				lost = 1;
			blk_ptr = __shfl_sync(kFull, p, 0);
			blk_left = grab;
		}
		tile_slots = blk_ptr + total <= a->stage_cap ? a->staging + blk_ptr : nullptr;
		tile_word = tile << (kRankBits + kPosBits);
		blk_ptr += total;
		blk_left -= total;
	}
	// entry k (= rank in the tile) of the current reservation
	__device__ __forceinline__ void put(uint32_t k, uint32_t pos) const {
		if (tile_slots)
			tile_slots[k] = tile_word + (((uint64_t) k << kPosBits) | pos);
	}
	__device__ __forceinline__ void end_tile(uint32_t total) {
		if (a->want_positions && lane_id() == 0) {
			if (idx >= a->cnt_cap)
				a->tile_count[tile] = total;
			else if (total)
				sts32(s_cnt + 4u * idx, total);
		}
		warp_count += total;
	}
	// the last block keeps an unused tail as well
	__device__ __forceinline__ void finish() const {
		if (lane_id() == 0 && n_log && !lost)
			sts64(log_s + 8u * (n_log - 1), lds64(log_s + 8u * (n_log - 1)) - ((unsigned long long) blk_left << 40));
	}
};

// Exact check of the window ending at virtual position e against every pattern of
// the bucket `key` hashes to (the HASH -> PREFIX -> compare tail of Wu-Manber,
// wu/wu.c:81-99, also used for the hits of a depth-truncated AC automaton).
// Returns how many distinct patterns end at e.
// Filtered AC (front_kind 1): the window of m symbols ending at virtual position e is walked through the full-depth
// automaton from its root (one symbol per lookup, ac/ac.c:207-219 with the failure function folded in); the walk ends
// on a final state iff the window is a pattern.  Reached by the few windows the block filter and the stage-2
// bitmap let through, i.e. hardly more often than there are matches.
static __device__ __noinline__ uint32_t verify_dfa(const ScanArgs &a, uint64_t e) {
	const uint32_t m = a.prm.m_min;
	if (e + 1 < a.data_lo + m || e >= a.data_hi || e < a.report_lo)
		return 0; // window would start before the text / end after it / not ours to report
	const uint64_t s0 = e + 1 - m;
	uint32_t ent = 0;
	for (uint32_t k = 0; k < m; k++) {
		const uint32_t sym = a.packed_in ? ((a.text16[(s0 + k) >> 2] >> (2 * ((s0 + k) & 3))) & 3u) : (a.text16[s0 + k] & 3u);
		ent = __ldg(a.vdfa + ((ent >> 1) << 2) + sym);
	}
	return ent & 1u;
}

__device__ __forceinline__ uint32_t verify_window(const ScanArgs &a, uint32_t key, uint64_t e) {
	const acwm_scan_params &p = a.prm;
	const uint32_t b = (uint32_t) (key * p.hb_mult) >> p.hb_sh;
	const uint32_t lo = __ldg(a.bucket_start + b), hi = __ldg(a.bucket_start + b + 1);
	uint32_t mult = 0;
	for (uint32_t i = lo; i < hi; i++) {
		const acwm_ventry en = a.entries[i];
		if (en.key != key)
			continue;
		// filtered AC: the automaton decides -- but only windows whose last symbols are those of some pattern get
		// that far (its walk is m dependent lookups; false positives of the stage-2 bitmap end here, two loads in)
		if (p.verify_kind)
			return verify_dfa(a, e);
		const uint32_t len = en.len & 0x7fffffffu;
		if (e + 1 < a.data_lo + len || e >= a.data_hi || e < a.report_lo)
			continue; // window would start before the text / end after it / not ours to report
		bool ok = true;
		if (!(en.len >> 31)) {
			const uint8_t *q = a.patterns + en.offset;
			if (a.packed_in) {
				const uint64_t s0 = e + 1 - len;
				for (uint32_t k = 0; k < len; k++)
					if (((a.text16[(s0 + k) >> 2] >> (2 * ((s0 + k) & 3))) & 3u) != q[k]) {
						ok = false;
						break;
					}
			} else {
				const uint8_t *t = a.text16 + (e + 1 - len);
				for (uint32_t k = 0; k < len; k++)
					if (t[k] != q[k]) {
						ok = false;
						break;
					}
			}
		}
		mult += ok;
	}
	return mult;
}

} // namespace acwm
