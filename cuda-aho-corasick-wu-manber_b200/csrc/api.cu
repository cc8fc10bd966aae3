// C ABI of libacwm_b200.so (see include/acwm.h): matcher handle, device residency,
// scan launches, host<->device pipeline.  The reference-shaped shims live in shims.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "hostpack.hpp"
#include "matcher.hpp"

namespace acwm {

cudaError_t launch_scan_packed(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st);
cudaError_t launch_scan_bytes(const ScanArgs &a, uint32_t threads, uint32_t smem, uint32_t grid, cudaStream_t st);

constexpr uint64_t kBounceEntries = 8192; // positions fetched together with the result block
constexpr uint64_t kBounceSlot = 1ull << 17; // positions per pinned slot of the pipelined fetch of a long result (1 MiB)

static thread_local std::string g_last_error;

int set_error(int code, const std::string &msg) {
	g_last_error = msg;
	return code;
}
int cuda_fail(cudaError_t e, const char *what) {
	g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
	return ACWM_ERR_CUDA;
}

#define CU(call)                                  \
	do {                                          \
		cudaError_t e_ = (call);                  \
		if (e_ != cudaSuccess)                    \
			return cuda_fail(e_, #call);          \
	} while (0)

template <class T>
static int dev_upload(T **dst, const void *src, size_t bytes) {
	*dst = nullptr;
	if (bytes == 0)
		bytes = 16; // keep pointers valid
	CU(cudaMalloc((void **) dst, (bytes + 15) & ~(size_t) 15));
	if (src)
		CU(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
	return ACWM_OK;
}

static int ensure_positions(acwm_matcher *mt, uint64_t cap) {
	if (cap <= mt->pos_cap)
		return ACWM_OK;
	if (mt->d_staging)
		cudaFree(mt->d_staging);
	if (mt->d_positions)
		cudaFree(mt->d_positions);
	mt->d_staging = mt->d_positions = nullptr;
	mt->pos_cap = 0;
	// staging: one reservation block of slack per warp the widest grid can hold
	// a warp's reservations double in size: at most as many slots idle as it fills, plus its first block
	mt->stage_cap = 2 * cap + (uint64_t) 2 * kStageBlock * 32 * (uint64_t) std::max(mt->sm_count, 1);
	mt->host_allocs++;
	CU(cudaMalloc((void **) &mt->d_staging, kScratchRing * mt->stage_cap * 8)); // one copy per scan in flight (see Work)
	CU(cudaMalloc((void **) &mt->d_positions, cap * 8));
	mt->pos_cap = cap;
	return ACWM_OK;
}

static int ensure_tiles(acwm_matcher *mt, uint64_t n_tiles) {
	if (n_tiles <= mt->tile_cap)
		return ACWM_OK;
	if (mt->d_tile_count)
		cudaFree(mt->d_tile_count);
	mt->d_tile_count = nullptr;
	mt->tile_cap = 0;
	const uint64_t want = n_tiles + n_tiles / 8 + 1024;
	mt->host_allocs++;
	CU(cudaMalloc((void **) &mt->d_tile_count, kScratchRing * want * 4)); // one copy per scan in flight
	mt->tile_cap = want;
	return ACWM_OK;
}

static int launch_scan(acwm_matcher *mt, const uint8_t *d_text, uint64_t n, uint64_t report_from, uint64_t tile_lo,
		uint64_t tile_hi, int want_positions, int append, int exchange, cudaStream_t st, int packed_in = 0);

static int do_upload(acwm_matcher *mt, int device, uint64_t pos_capacity) {
	if (device >= 0)
		CU(cudaSetDevice(device));
	CU(cudaGetDevice(&mt->device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, mt->device));
	if (prop.major < 10)
		return set_error(ACWM_ERR_UNSUPPORTED, "libacwm_b200 is built for sm_100a (B200) only");
	mt->sm_count = prop.multiProcessorCount;
	mt->l2_persist_max = (size_t) prop.persistingL2CacheMaxSize;
	mt->l2_window_max = (size_t) prop.accessPolicyMaxWindowSize;
	const Compiled &c = mt->c;
	int rc;
	// all tables in ONE device allocation, so that a single access-policy window keeps them in L2
	const struct {
		const void *src;
		size_t bytes;
	} parts[7] = {{c.front.data(), c.front.size()}, {c.rmask.data(), c.rmask.size()},
			{c.filter2.data(), c.filter2.size() * 4}, {c.bucket_start.data(), c.bucket_start.size() * 4},
			{c.entries.data(), c.entries.size() * sizeof(acwm_ventry)}, {mt->ps.bytes.data(), mt->ps.bytes.size()},
			{c.vdfa.data(), c.vdfa.size() * 4}};
	size_t off[7], total = 0;
	for (int i = 0; i < 7; i++) {
		off[i] = total;
		total += (std::max<size_t>(parts[i].bytes, 16) + 255) & ~(size_t) 255;
	}
	if ((rc = dev_upload(&mt->d_tables, nullptr, total)))
		return rc;
	CU(cudaMemset(mt->d_tables, 0, total));
	for (int i = 0; i < 7; i++)
		if (parts[i].bytes)
			CU(cudaMemcpy(mt->d_tables + off[i], parts[i].src, parts[i].bytes, cudaMemcpyHostToDevice));
	// the access-policy window covers what every tile reads; the verify DFA (last part) is touched by the rare
	// candidate windows only and stays outside
	mt->tables_bytes = off[6];
	mt->d_vdfa = reinterpret_cast<uint32_t *>(mt->d_tables + off[6]);
	mt->d_front = mt->d_tables + off[0];
	mt->d_rmask = mt->d_tables + off[1];
	mt->d_filter2 = reinterpret_cast<uint32_t *>(mt->d_tables + off[2]);
	mt->d_bucket_start = reinterpret_cast<uint32_t *>(mt->d_tables + off[3]);
	mt->d_entries = reinterpret_cast<acwm_ventry *>(mt->d_tables + off[4]);
	mt->d_patterns = mt->d_tables + off[5];
	CU(cudaMalloc((void **) &mt->d_ctl, sizeof(Control)));
	CU(cudaMemset(mt->d_ctl, 0, sizeof(Control)));
	CU(cudaMallocHost((void **) &mt->h_res, sizeof(Result)));
	CU(cudaMallocHost((void **) &mt->h_bounce, kBounceEntries * 8));
	CU(cudaMalloc((void **) &mt->d_cta_total, kScratchRing * kMaxScanBlocks * sizeof(unsigned long long)));
	// the span totals are recognised by their launch tag: recycled device memory must not hold a look-alike
	CU(cudaMemset(mt->d_cta_total, 0, kScratchRing * kMaxScanBlocks * sizeof(unsigned long long)));
	CU(cudaStreamCreateWithFlags(&mt->s_copy, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&mt->s_scan, cudaStreamNonBlocking));
	for (auto &e : mt->ev_copy)
		CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	// tables read from global memory by the kernels are kept L2-resident (access-policy window, apply_l2_window)
	mt->l2_tables = !c.info.table_in_smem || (c.prm.r_entries && !c.prm.r_in_smem) || (c.prm.f2_words && !c.prm.f2_in_smem);
	if (mt->l2_tables && mt->l2_persist_max) {
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(mt->tables_bytes, mt->l2_persist_max));
		(void) cudaGetLastError();
	}
	mt->uploaded = true;
	if (pos_capacity && (rc = ensure_positions(mt, pos_capacity)))
		return rc;
	// One empty scan now: the kernel's module is loaded and its attributes are set here, not inside the first search
	// (whose event-bracketed kernel time, acwm_last_kernel_seconds, is kernel-only like the reference's gpuTime,
	// cuda/cuda_wm.cu:264-289)
	if ((rc = launch_scan(mt, mt->d_tables, 0, 0, 0, 0, 0, 0, 0, mt->s_scan)))
		return rc;
	CU(cudaStreamSynchronize(mt->s_scan));
	return ACWM_OK;
}

static void apply_l2_window(acwm_matcher *mt, cudaStream_t st) {
	if (!mt->l2_tables || !mt->l2_window_max || mt->l2_window_stream == (void *) st)
		return;
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	const size_t win = std::min(mt->tables_bytes, mt->l2_window_max);
	v.accessPolicyWindow.base_ptr = mt->d_tables;
	v.accessPolicyWindow.num_bytes = win;
	v.accessPolicyWindow.hitRatio = mt->l2_persist_max ? (float) std::min(1.0, (double) mt->l2_persist_max / (double) win) : 1.0f;
	v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
	(void) cudaGetLastError();
	mt->l2_window_stream = (void *) st; // set once per stream, not once per scan
}

constexpr uint64_t kBigScanBytes = 256ull << 20; // launches of at least this much text: see launch_scan

// Kernel variants that can be switched off for A/B measurements (scripts/tune.py): ACWM_TUNE = OR of
// 1 (warp-cooperative candidate verification).
static uint32_t kernel_tune() {
	static const uint32_t v = [] {
		const char *e = getenv("ACWM_TUNE");
		return e && *e ? (uint32_t) strtoul(e, nullptr, 0) : (kTuneCoopVerify | (kTuneLaneLocalDefault << 8));
	}();
	return v;
}

// Launch the scan of warp tiles [tile_lo, tile_hi) of the text at d_text (n bytes): one
// kernel that scans, orders the positions and publishes the result block.
static int launch_scan(acwm_matcher *mt, const uint8_t *d_text, uint64_t n, uint64_t report_from, uint64_t tile_lo,
		uint64_t tile_hi, int want_positions, int append, int exchange, cudaStream_t st, int packed_in) {
	const Compiled &c = mt->c;
	ScanArgs a;
	memset(&a, 0, sizeof(a));
	const uint64_t mis = (uint64_t) (uintptr_t) d_text & 15u;
	a.text16 = d_text - mis;
	a.data_lo = mis;
	a.data_hi = mis + n;
	a.report_lo = mis + std::max<uint64_t>(c.prm.m_min - 1, report_from);
	a.tile_lo = tile_lo;
	a.tile_hi = tile_hi;
	a.front = mt->d_front;
	a.front_bytes = (uint32_t) c.front.size();
	a.front_in_smem = c.info.table_in_smem;
	a.rmask = mt->d_rmask;
	a.filter2 = mt->d_filter2;
	a.bucket_start = mt->d_bucket_start;
	a.entries = mt->d_entries;
	a.patterns = mt->d_patterns;
	a.vdfa = mt->d_vdfa;
	a.prm = c.prm;
	a.ctl = mt->d_ctl;
	const uint32_t par = mt->epoch % kScratchRing; // the scratch arrays of this launch (its two predecessors may still be using theirs)
	a.staging = mt->d_staging ? mt->d_staging + par * mt->stage_cap : nullptr;
	a.positions = mt->d_positions;
	a.cap = want_positions ? mt->pos_cap : 0;
	a.stage_cap = want_positions ? mt->stage_cap : 0;
	a.tile_count = mt->d_tile_count ? mt->d_tile_count + par * mt->tile_cap : nullptr;
	a.cta_total = mt->d_cta_total + par * kMaxScanBlocks;
	a.stages = c.info.stages;
	a.epoch = mt->epoch++;
	a.want_positions = want_positions;
	a.append = append;
	a.tune = kernel_tune();
	a.lane_local = (a.tune & kTuneCoopVerify) ? ((a.tune >> 8) & 0xffu) : kListCap;
	a.packed_in = packed_in ? 1u : 0u;
	a.trace = mt->d_trace;
	if (exchange && mt->peer_world > 1) {
		a.world = mt->peer_world;
		a.rank = mt->peer_rank;
		a.xepoch = ++mt->xepoch; // starts at 1: a zeroed mailbox never matches
		for (uint32_t r = 0; r < mt->peer_world; r++)
			a.peers[r] = reinterpret_cast<unsigned long long *>(mt->peer_ptrs[r]);
	}
	const uint64_t ntl = tile_hi - tile_lo;
	// Long scans with a verification stage leave the two-half-size-CTAs shape for one full-size CTA per SM: 32 warps
	// instead of 24 (c2 at 1 GiB per launch: 5.8 against 5.2 TB/s), and what two CTAs per SM hide -- one launch's
	// prologue and epilogue under its neighbour's scan -- is a few microseconds per launch.  ACWM_BIG_SHAPE=0/1 forces.
	static const int big_env = [] {
		const char *e = getenv("ACWM_BIG_SHAPE");
		return e && *e ? atoi(e) : -1;
	}();
	const bool exact = c.prm.algo == ACWM_ALGO_AC && !c.prm.front_kind && c.prm.exact_front;
	const bool big = c.alt_threads && (big_env >= 0 ? big_env != 0 : (!exact && ntl * kTile >= kBigScanBytes));
	const uint32_t threads = big ? c.alt_threads : c.info.threads, warps = threads / 32;
	const uint32_t shape_smem_bytes = big ? c.alt_smem_bytes : c.info.smem_bytes;
	const bool dual = !big && c.info.ctas_per_sm == 2;
	const uint32_t sms = (uint32_t) std::max(mt->sm_count, 1);
	// Overlap mode (device-resident scans only; exchange == "called from acwm_scan_device"): one CTA per SM, launched
	// as a programmatic dependent launch, so that the CTAs of consecutive scans share the SMs (two half-size CTAs of
	// two scans per SM) or follow each other on them without a grid-wide gap (one full-size CTA per SM).  What keeps
	// at most three scans in flight (see Work in scan_common.cuh): the grid fills every SM and an SM holds at most
	// two CTAs of this kernel -- dual CTAs take more than a third of its shared memory, single ones more than half.
	const bool chain = exchange && mt->overlap && ntl >= (uint64_t) sms * warps;
	// every CTA owns a contiguous span of whole "rounds" (one tile per warp)
	const uint32_t max_grid = std::min<uint32_t>(chain ? sms : sms * (dual ? 2u : 1u), kMaxScanBlocks);
	uint32_t grid = (uint32_t) std::min<uint64_t>(max_grid, (ntl + warps - 1) / warps);
	grid = std::max(grid, 1u);
	a.tiles_per_cta = std::max<uint64_t>(1, (ntl + grid - 1) / grid);
	grid = (uint32_t) std::max<uint64_t>(1, (ntl + a.tiles_per_cta - 1) / a.tiles_per_cta);
	// per-tile counts of a span stay in whatever shared memory the tables and the rings leave free
	const uint32_t smem_max = dual ? kMaxSmemDual : kMaxSmem;
	a.cnt_cap = (uint32_t) std::min<uint64_t>(a.tiles_per_cta, (smem_max - shape_smem_bytes) / 4);
	const uint32_t smem = shape_smem_bytes + a.cnt_cap * 4;
	{ // the shared-memory layout of scan_kernel.cuh: [1 KiB][per-warp areas][front][offset masks][stage-2 bitmap][per-tile counts]
		const bool pk_copy = c.prm.packed2bit && !exact;
		const uint32_t stages = c.prm.packed2bit ? 1u : c.info.stages;
		const uint32_t front = a.front_in_smem ? ((a.front_bytes + 15u) & ~15u) : 0u;
		const uint32_t rm = (exact || !c.prm.r_in_smem) ? 0u : ((c.prm.r_entries * c.prm.r_entry_bytes + 15u) & ~15u);
		const uint32_t f2 = (exact || !c.prm.f2_in_smem) ? 0u : ((c.prm.f2_words * 4u + 15u) & ~15u);
		a.s_rmask = kDynSmemBase + kSmemReserve + warps * warp_smem_bytes(stages, pk_copy) + front;
		a.s_f2 = a.s_rmask + rm;
		a.s_cnt = a.s_f2 + f2;
		if (a.s_cnt - kDynSmemBase + a.cnt_cap * 4 > smem) {
			return set_error(ACWM_ERR_INVALID, "scan kernel: shared-memory layout exceeds the launch size");
		}
	}
	const bool two_at_most = 3 * (smem + 1024) > kSmemPerSmTotal; // CTAs of this kernel per SM
	a.pdl = (chain && grid == sms && two_at_most) ? 1 : 0;
	if (grid > 256)
		a.trace = nullptr; // the trace buffer holds 256 CTAs
	cudaError_t e = c.prm.packed2bit ? launch_scan_packed(a, threads, smem, grid, st)
									 : launch_scan_bytes(a, threads, smem, grid, st);
	if (e != cudaSuccess)
		return cuda_fail(e, "scan kernel launch");
	mt->launches++;
	return ACWM_OK;
}

// Sum of the mailbox for the last exchange epoch (the scan kernels collect epoch x-1 while they run epoch x).
__global__ void collect_last_kernel(Control *ctl, const unsigned long long *box, uint32_t world, uint32_t x) {
	if (ctl->result.global_epoch != x) {
		unsigned long long sum;
		if (collect_mailbox(box, world, x, sum)) {
			ctl->result.global_count = sum;
			ctl->result.global_epoch = x;
		} else
			ctl->result.exchange_failed = x;
	}
}

// Host text of a 2-bit matcher (alphabet <= 4): the host cores pack it 4 symbols per byte into a ring of pinned
// buffers (hostpack.cpp), a quarter of the bytes crosses PCIe, and the scan takes the packed tiles as they are.
// The packer threads run ahead through the text; this thread takes the chunks in order (and packs along while it
// waits), sends each one and launches its scan.
constexpr uint64_t kHostPackMin = 4ull << 20;
constexpr unsigned kHostPackMinThreads = 3; // fewer: the text goes one byte per symbol (the ranks of a torchrun box share its cores)
constexpr uint64_t kPackChunk = 56 * HostPacker::kPieceSymbols; // 14 Mi symbols = 4096 tiles = 56 work items
constexpr unsigned kPackRing = 16;
static_assert(kPackChunk % kTile == 0 && kPackChunk % 64 == 0, "chunks are whole tiles and whole 16-byte pieces");
// Share of a pinned host text that is sent unpacked beside the packed rest.  ACWM_HOST_RAW_PERCENT fixes it
// (0 = none).  Otherwise the FIRST search of a matcher takes it from the packer's thread count: both transfers share the
// link, so link time n(r + (1-r)/4)/L and packing time n(1-r)/P meet at r = (1/P - 1/4L) / (3/4L + 1/P), with P = 5.2 GB/s
// per packer thread on text that comes from DRAM (9 with the AVX-512 inner loop) and L = 56 GB/s for the link -- and from
// then on the share CLIMBS on what the calls measure: every search records its own duration (setup excluded) under the
// number of raw chunks it ran with; the next one runs with the best number so far, or with a neighbour of it that has no
// figure yet (every 16th call re-tries a neighbour).  What the link and the cores really deliver depends
// on who else uses them (the ranks of one box share its cores, its memory and its PCIe root), so the share follows the
// box instead of a model of it.
constexpr double kMaxRawShare = 0.6;
static double host_raw_share_model(unsigned threads) {
	const double per_thread = __builtin_cpu_supports("avx512bw") ? 9.0 : 5.2;
	const double P = per_thread * std::max(1u, threads), L = 56.0;
	const double r = (1.0 / P - 0.25 / L) / (0.75 / L + 1.0 / P);
	return std::min(kMaxRawShare, std::max(0.0, r));
}
static int host_raw_percent_env() {
	const char *e = getenv("ACWM_HOST_RAW_PERCENT");
	return e && *e ? (int) std::min<long>(std::max<long>(atol(e), 0), 90) : -1;
}
// ACWM_HOST_PACK: 0 = never, 2 = always (tests), unset / 1 = when the host has the cores for it
static int host_pack_mode() {
	const char *e = getenv("ACWM_HOST_PACK");
	return e && *e ? atoi(e) : 1;
}

static int search_host_packed(acwm_matcher *mt, const uint8_t *text, uint64_t n, uint64_t *count, uint64_t *positions,
		uint64_t cap, uint64_t *n_written, int want_positions) {
	int rc;
	const uint64_t T = kTile;
	const uint64_t n_tiles = (n + T - 1) / T;
	const uint64_t packed_total = ((n + 63) / 64) * 16; // zero-padded to whole 16-byte pieces
	if (packed_total + 64 > mt->text_cap) {
		if (mt->d_text)
			cudaFree(mt->d_text);
		mt->d_text = nullptr;
		mt->text_cap = 0;
		mt->host_allocs++;
		CU(cudaMalloc((void **) &mt->d_text, packed_total + 64));
		mt->text_cap = packed_total + 64;
	}
	if (want_positions && (rc = ensure_tiles(mt, n_tiles)))
		return rc;
	const uint64_t chunk_tiles = kPackChunk / T, slot_bytes = kPackChunk / 4 + 64;
	const uint64_t n_chunks = std::max<uint64_t>(1, (n + kPackChunk - 1) / kPackChunk);
	if (!mt->h_pack_ring) {
		mt->host_allocs++;
		CU(cudaMallocHost((void **) &mt->h_pack_ring, kPackRing * slot_bytes));
		for (auto &e : mt->ev_pack)
			CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	while (mt->ev_time.size() < 2 * n_chunks) {
		cudaEvent_t e;
		CU(cudaEventCreate(&e));
		mt->ev_time.push_back(e);
	}
	apply_l2_window(mt, mt->s_scan);
	mt->first_epoch = mt->epoch;
	static const bool dbg = getenv("ACWM_DEBUG_TIMING") != nullptr;
	auto now = [] {
		timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec + 1e-9 * ts.tv_nsec;
	};
	const double t_begin = now();
	double t_wait = 0;
	std::vector<double> t_issued;
	HostPacker &pk = *mt->packer;
	struct Release { // error returns too
		HostPacker &p;
		~Release() { p.finish(); }
	} release{pk};
	// The link moves raw text by DMA while the cores pack: a prefix of the text (whole chunks; the share: see host_raw_share_model above, about
	// where link time and packing time meet) goes one byte per symbol from the caller's
	// PINNED buffer on a second copy stream, the rest is packed.  The packed part brings its own history (the 64-symbol
	// halo of its first tile and the reach of the longest pattern's compare), packed by this thread.
	uint64_t R = 0;
	const uint32_t H = (std::max<uint32_t>(64u, mt->c.prm.m_max) + 63u) & ~63u;
	const int raw_env = host_raw_percent_env();
	const uint64_t r_max = std::min<uint64_t>({n_chunks - 1, (uint64_t) (n_chunks * kMaxRawShare), (uint64_t) acwm_matcher::kRawTimes - 1});
	bool climbing = false;
	if (n_chunks >= 5 && H <= 4096) {
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, text) == cudaSuccess && at.type == cudaMemoryTypeHost) {
			if (raw_env >= 0)
				R = std::min<uint64_t>(n_chunks - 1, (uint64_t) (n_chunks * (raw_env / 100.0) + 0.5));
			else {
				climbing = true;
				if (mt->raw_chunks_of != n_chunks) { // another text size: start from the model
					mt->raw_chunks_of = n_chunks;
					mt->raw_chunks = std::min<uint64_t>(r_max, (uint64_t) (n_chunks * host_raw_share_model(mt->packer->threads()) + 0.5));
					mt->raw_calls = 0;
					for (auto &t : mt->raw_time)
						t = 0;
				}
				R = std::min<uint64_t>(mt->raw_chunks, r_max);
			}
		}
		(void) cudaGetLastError();
	}
	const uint64_t n_raw = R * kPackChunk;
	if (R) {
		if (n_raw + 64 > mt->raw_cap) {
			if (mt->d_raw)
				cudaFree(mt->d_raw);
			mt->d_raw = nullptr;
			mt->raw_cap = 0;
			mt->host_allocs++;
			// room for the largest share the adaptation may reach (kMaxRawShare), so that a moving share never reallocates
			const uint64_t want = std::max<uint64_t>(n_raw, ((uint64_t) (n_chunks * kMaxRawShare) + 1) * kPackChunk) + 64;
			CU(cudaMalloc((void **) &mt->d_raw, want));
			mt->raw_cap = want;
		}
		if (!mt->s_copy2) {
			CU(cudaStreamCreateWithFlags(&mt->s_copy2, cudaStreamNonBlocking));
			CU(cudaMallocHost((void **) &mt->h_hist, 4096 / 4 + 16));
			for (auto &e : mt->ev_hyb)
				CU(cudaEventCreate(&e));
		}
		CU(cudaEventRecord(mt->ev_hyb[0], mt->s_copy2));
	}
	const double t_work = now(); // everything above may have allocated
	const uint64_t allocs0 = mt->host_allocs;
	pk.begin(text + n_raw, n - n_raw, kPackChunk, mt->h_pack_ring, slot_bytes, kPackRing);
	for (uint64_t ci = 0; ci < R; ci++) { // the raw prefix: all copies and scans queued at once
		const uint64_t b0 = ci * kPackChunk;
		CU(cudaMemcpyAsync(mt->d_raw + b0, text + b0, kPackChunk, cudaMemcpyHostToDevice, mt->s_copy2));
		cudaEvent_t ev = mt->ev_copy[ci % mt->ev_copy.size()];
		CU(cudaEventRecord(ev, mt->s_copy2));
		CU(cudaStreamWaitEvent(mt->s_scan, ev, 0));
		CU(cudaEventRecord(mt->ev_time[2 * ci], mt->s_scan));
		if ((rc = launch_scan(mt, mt->d_raw, n_raw, mt->host_report_from, ci * chunk_tiles, (ci + 1) * chunk_tiles, want_positions, ci > 0, 0,
					 mt->s_scan, 0)))
			return rc;
		CU(cudaEventRecord(mt->ev_time[2 * ci + 1], mt->s_scan));
	}
	if (R) {
		CU(cudaEventRecord(mt->ev_hyb[1], mt->s_copy2));
		CU(cudaEventRecord(mt->ev_hyb[2], mt->s_copy));
		HostPacker::pack_now(text + n_raw - H, mt->h_hist, H);
		CU(cudaMemcpyAsync(mt->d_text + (n_raw - H) / 4, mt->h_hist, H / 4, cudaMemcpyHostToDevice, mt->s_copy));
	}
	uint64_t oldest = 0; // first packed chunk whose copy is not known to be complete (its ring slot is still taken)
	for (uint64_t ci = R; ci < n_chunks; ci++) {
		const uint64_t b0 = std::min<uint64_t>(n, ci * kPackChunk), b1 = std::min<uint64_t>(n, (ci + 1) * kPackChunk);
		const uint64_t cj = ci - R; // chunk number of the packing job
		while (cj >= oldest + kPackRing) { // the packers may not start this chunk before its slot is free
			CU(cudaEventSynchronize(mt->ev_pack[oldest % kPackRing]));
			pk.recycle(oldest++);
		}
		const double t0 = now();
		pk.wait_chunk(cj);
		t_wait += now() - t0;
		uint8_t *slot = mt->h_pack_ring + (cj % kPackRing) * slot_bytes;
		const uint64_t used = (b1 - b0 + 3) / 4, bytes = ((b1 - b0 + 63) / 64) * 16;
		if (bytes > used)
			memset(slot + used, 0, bytes - used); // the last piece of the text: zero padding
		if (dbg)
			t_issued.push_back((now() - t_begin) * 1e3);
		if (bytes)
			CU(cudaMemcpyAsync(mt->d_text + b0 / 4, slot, bytes, cudaMemcpyHostToDevice, mt->s_copy));
		cudaEvent_t ev = mt->ev_pack[cj % kPackRing];
		CU(cudaEventRecord(ev, mt->s_copy));
		CU(cudaStreamWaitEvent(mt->s_scan, ev, 0));
		CU(cudaEventRecord(mt->ev_time[2 * ci], mt->s_scan));
		if ((rc = launch_scan(mt, mt->d_text, n, mt->host_report_from, std::min(n_tiles, ci * chunk_tiles),
					 std::min(n_tiles, (ci + 1) * chunk_tiles), want_positions, ci > 0, 0, mt->s_scan, 1)))
			return rc;
		CU(cudaEventRecord(mt->ev_time[2 * ci + 1], mt->s_scan));
		while (oldest < cj && cudaEventQuery(mt->ev_pack[oldest % kPackRing]) == cudaSuccess)
			pk.recycle(oldest++); // let the packers run further ahead
	}
	(void) cudaGetLastError(); // cudaEventQuery's cudaErrorNotReady is not an error
	if (R)
		CU(cudaEventRecord(mt->ev_hyb[3], mt->s_copy));
	const uint64_t bad = pk.bad();
	mt->last_h2d_bytes = n_raw + (packed_total - n_raw / 4) + (R ? H / 4 : 0);
	mt->last_want_positions = want_positions;
	const double t_issue = now();
	rc = acwm_fetch(mt, count, positions, cap, n_written, mt->s_scan);
	pk.finish();
	if (climbing && (rc == ACWM_OK || rc == ACWM_ERR_OVERFLOW) && mt->host_allocs == allocs0) {
		// this call's duration under its number of raw chunks; the next call: the best of R - 1, R, R + 1, a neighbour
		// without a figure first, and every 16th call the neighbour whose figure is the oldest guess
		const double t_call = now() - t_work;
		double *T = mt->raw_time;
		T[R] = T[R] > 0 ? 0.5 * T[R] + 0.5 * t_call : t_call;
		uint64_t best = R; // the best point so far; its neighbours are what is worth knowing
		for (uint64_t r = 0; r <= r_max; r++)
			if (T[r] > 0 && T[r] < T[best])
				best = r;
		const uint64_t lo = best > 0 ? best - 1 : best, hi = std::min<uint64_t>(best + 1, r_max);
		uint64_t next = best;
		mt->raw_calls++;
		if (T[hi] <= 0)
			next = hi;
		else if (T[lo] <= 0)
			next = lo;
		else if (mt->raw_calls % 16 == 0)
			next = (mt->raw_calls / 16) % 2 ? hi : lo; // figures age: look at a neighbour again
		mt->raw_chunks = next;
		if (dbg)
			fprintf(stderr, "  hybrid: %llu of %llu chunks raw: %.3f ms -> next %llu\n", (unsigned long long) R, (unsigned long long) n_chunks,
					t_call * 1e3, (unsigned long long) next);
	}
	(void) cudaGetLastError();
	double secs = 0;
	for (uint64_t ci = 0; ci < n_chunks; ci++) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, mt->ev_time[2 * ci], mt->ev_time[2 * ci + 1]) == cudaSuccess)
			secs += ms * 1e-3;
	}
	mt->last_kernel_s = secs;
	if (dbg) {
		fprintf(stderr, "  copy of chunk i issued at (ms after the call):");
		for (double t : t_issued)
			fprintf(stderr, " %.3f", t);
		fprintf(stderr, "\n  scan of chunk i done at (ms after the first chunk had landed):");
		for (uint64_t ci = 0; ci < n_chunks; ci++) {
			float ms = 0;
			cudaEventElapsedTime(&ms, mt->ev_time[0], mt->ev_time[2 * ci + 1]);
			fprintf(stderr, " %.3f", ms);
		}
		fprintf(stderr, "\n");
	}
	if (dbg)
		fprintf(stderr, "acwm host-packed search: n %llu, %llu chunks (%llu of them sent unpacked), %u threads: waited for the packers %.3f ms, issue loop %.3f ms, fetch %.3f ms, kernels %.3f ms\n",
				(unsigned long long) n, (unsigned long long) n_chunks, (unsigned long long) R, pk.threads(), t_wait * 1e3,
				(t_issue - t_begin) * 1e3,
				(now() - t_issue) * 1e3, secs * 1e3);
	if (bad & 0xFCFCFCFCFCFCFCFCull)
		return set_error(ACWM_ERR_BAD_TEXT, "text holds a byte >= 4 but the matcher was built for alphabet <= 4");
	return rc;
}

} // namespace acwm

using namespace acwm;

extern "C" {

const char *acwm_last_error(void) { return g_last_error.c_str(); }

int acwm_build(int algo, const uint8_t *patterns, const uint32_t *lens, uint32_t m, uint32_t p, uint32_t alphabet,
		const acwm_options *opts, acwm_matcher **out) {
	if (!out)
		return set_error(ACWM_ERR_INVALID, "out == NULL");
	*out = nullptr;
	acwm_matcher *mt = new (std::nothrow) acwm_matcher();
	if (!mt)
		return set_error(ACWM_ERR_NOMEM, "host allocation failed");
	if (opts)
		mt->opts = *opts;
	std::string err;
	int rc;
	try { // the table compiler grows std::vectors: no exception leaves the C ABI
		rc = normalize_patterns(patterns, lens, m, p, alphabet, mt->ps, err);
		if (rc == ACWM_OK)
			rc = compile_tables(algo, mt->ps, mt->opts, mt->c, err);
	} catch (const std::bad_alloc &) {
		rc = ACWM_ERR_NOMEM;
		err = "host allocation failed while compiling the tables";
	} catch (const std::exception &e) {
		rc = ACWM_ERR_INVALID;
		err = std::string("table compiler: ") + e.what();
	}
	if (rc != ACWM_OK) {
		delete mt;
		return set_error(rc, err);
	}
	*out = mt;
	return ACWM_OK;
}

int acwm_upload(acwm_matcher *mt, int device, uint64_t pos_capacity) {
	if (!mt)
		return set_error(ACWM_ERR_INVALID, "matcher == NULL");
	if (mt->uploaded) {
		if (device >= 0 && device != mt->device)
			return set_error(ACWM_ERR_INVALID, "matcher already resident on another device");
		CU(cudaSetDevice(mt->device));
		return ensure_positions(mt, pos_capacity);
	}
	return do_upload(mt, device, pos_capacity);
}

int acwm_scan_device(acwm_matcher *mt, const uint8_t *d_text, uint64_t n, uint64_t report_from, int want_positions,
		void *stream) {
	if (!mt || (!d_text && n))
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	int rc;
	if (!mt->uploaded && (rc = do_upload(mt, -1, 0)))
		return rc;
	if (want_positions && mt->pos_cap == 0)
		return set_error(ACWM_ERR_INVALID, "positions requested but the matcher was uploaded with pos_capacity = 0");
	cudaStream_t st = (cudaStream_t) stream;
	const uint64_t mis = (uint64_t) (uintptr_t) d_text & 15u;
	const uint64_t T = kTile;
	const uint64_t n_tiles = (mis + n + T - 1) / T; // n == 0: no tile, the launch only publishes an empty result
	if (want_positions && (rc = ensure_tiles(mt, n_tiles)))
		return rc;
	apply_l2_window(mt, st);
	if (mt->profiling) {
		if (!mt->ev_prof[0])
			for (auto &e : mt->ev_prof)
				CU(cudaEventCreate(&e));
		CU(cudaEventRecord(mt->ev_prof[0], st));
	}
	mt->first_epoch = mt->epoch;
	if ((rc = launch_scan(mt, d_text, n, report_from, 0, n_tiles, want_positions, 0, 1, st)))
		return rc;
	if (mt->profiling)
		CU(cudaEventRecord(mt->ev_prof[1], st));
	mt->last_want_positions = want_positions;
	return ACWM_OK;
}

int acwm_fetch(acwm_matcher *mt, uint64_t *count, uint64_t *positions, uint64_t cap, uint64_t *n_written,
		void *stream) {
	if (!mt || !mt->uploaded)
		return set_error(ACWM_ERR_INVALID, "matcher not uploaded");
	cudaStream_t st = (cudaStream_t) stream;
	static const bool dbg = getenv("ACWM_DEBUG_TIMING") != nullptr;
	timespec ts0, ts1;
	clock_gettime(CLOCK_MONOTONIC, &ts0);
	CU(cudaMemcpyAsync(mt->h_res, &mt->d_ctl->result, sizeof(Result), cudaMemcpyDeviceToHost, st));
	// the first positions ride along (pinned bounce buffer): a sparse result needs no second round trip
	const uint64_t spec = (positions && mt->last_want_positions) ? std::min<uint64_t>({cap, mt->pos_cap, kBounceEntries}) : 0;
	if (spec)
		CU(cudaMemcpyAsync(mt->h_bounce, mt->d_positions, spec * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	clock_gettime(CLOCK_MONOTONIC, &ts1);
	if (dbg)
		fprintf(stderr, "  acwm_fetch: result block after %.3f ms\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
	const Result &h = *mt->h_res;
	if (count)
		*count = h.count;
	if (n_written)
		*n_written = 0;
	if (h.order_failed > mt->first_epoch)
		return set_error(ACWM_ERR_CUDA, "the position ordering gave up waiting for a CTA of its own launch");
	if (h.bad_text)
		return set_error(ACWM_ERR_BAD_TEXT, "text holds a byte >= 4 but the matcher was built for alphabet <= 4");
	if (positions && mt->last_want_positions) {
		const uint64_t have = std::min<uint64_t>(h.written, mt->pos_cap);
		const uint64_t w = std::min<uint64_t>(have, cap);
		const uint64_t have_b = std::min(w, spec);
		if (have_b)
			memcpy(positions, mt->h_bounce, have_b * 8);
		if (w > have_b) {
			// the rest: straight into pinned caller memory, else through two pinned 1 MiB slots (copy of slot k + 1 under
			// the memcpy of slot k) -- a plain cudaMemcpy into pageable memory moves 2-3 GB/s, this 20+
			cudaPointerAttributes at;
			const bool pinned_dst = cudaPointerGetAttributes(&at, positions) == cudaSuccess && at.type == cudaMemoryTypeHost;
			(void) cudaGetLastError();
			if (pinned_dst) {
				CU(cudaMemcpyAsync(positions + have_b, mt->d_positions + have_b, (w - have_b) * 8, cudaMemcpyDeviceToHost, st));
				CU(cudaStreamSynchronize(st));
			} else {
				if (!mt->h_bounce2) {
					mt->host_allocs++;
					CU(cudaMallocHost((void **) &mt->h_bounce2, 2 * kBounceSlot * 8));
					for (auto &e : mt->ev_bounce)
						CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				}
				uint64_t issued = have_b, copied = have_b;
				unsigned k_issue = 0, k_copy = 0;
				while (copied < w) {
					while (issued < w && k_issue < k_copy + 2) { // keep both slots busy
						const uint64_t cnt = std::min<uint64_t>(kBounceSlot, w - issued);
						CU(cudaMemcpyAsync(mt->h_bounce2 + (k_issue & 1) * kBounceSlot, mt->d_positions + issued, cnt * 8,
								cudaMemcpyDeviceToHost, st));
						CU(cudaEventRecord(mt->ev_bounce[k_issue & 1], st));
						issued += cnt;
						k_issue++;
					}
					const uint64_t cnt = std::min<uint64_t>(kBounceSlot, w - copied);
					CU(cudaEventSynchronize(mt->ev_bounce[k_copy & 1]));
					memcpy(positions + copied, mt->h_bounce2 + (k_copy & 1) * kBounceSlot, cnt * 8);
					copied += cnt;
					k_copy++;
				}
			}
		}
		if (dbg) {
			clock_gettime(CLOCK_MONOTONIC, &ts0);
			fprintf(stderr, "  acwm_fetch: %llu positions after another %.3f ms\n", (unsigned long long) w,
					(ts0.tv_sec - ts1.tv_sec) * 1e3 + (ts0.tv_nsec - ts1.tv_nsec) * 1e-6);
		}
		if (n_written)
			*n_written = w;
		if (h.overflow || h.count > cap)
			return set_error(ACWM_ERR_OVERFLOW, "more matches than position capacity");
	}
	return ACWM_OK;
}

int acwm_result_device_ptrs(acwm_matcher *mt, uint64_t **d_count, uint64_t **d_positions) {
	if (!mt || !mt->uploaded)
		return set_error(ACWM_ERR_INVALID, "matcher not uploaded");
	if (d_count)
		*d_count = reinterpret_cast<uint64_t *>(&mt->d_ctl->result.count);
	if (d_positions)
		*d_positions = mt->d_positions;
	return ACWM_OK;
}

static int search_host_raw(acwm_matcher *mt, const uint8_t *text, uint64_t n, uint64_t *count, uint64_t *positions, uint64_t cap,
		uint64_t *n_written, int want_positions);

int acwm_search_host(acwm_matcher *mt, const uint8_t *text, uint64_t n, uint64_t *count, uint64_t *positions,
		uint64_t cap, uint64_t *n_written) {
	if (!mt || (!text && n))
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	int rc;
	if (!mt->uploaded && (rc = do_upload(mt, -1, 0)))
		return rc;
	CU(cudaSetDevice(mt->device));
	const int want_positions = positions != nullptr && cap > 0;
	if (want_positions && (rc = ensure_positions(mt, cap)))
		return rc;
	const int pack_mode = host_pack_mode();
	if (mt->c.prm.packed2bit && n >= kHostPackMin && pack_mode > 0) {
		if (!mt->packer) {
			try { // thread creation may throw: no exception leaves the C ABI, the text then goes unpacked
				mt->packer = new HostPacker();
			} catch (...) {
				mt->packer = nullptr;
				return search_host_raw(mt, text, n, count, positions, cap, n_written, want_positions);
			}
		}
		// Packing pays when the cores out-run what the link carries raw (a core packs 5-7 GB/s, one PCIe link moves
		// ~55 GB/s of raw text) -- and when the box has the memory bandwidth for both: the first searches run the hybrid
		// transfer, then the plain copy (one call each to set up, one to measure), and from then on the faster of the
		// two runs (the other is re-tried every 16th call).
		if (pack_mode == 2)
			return search_host_packed(mt, text, n, count, positions, cap, n_written, want_positions);
		if (mt->packer->threads() >= kHostPackMinThreads) {
			// a call that had to allocate (the first of either kind, a longer text) does not count as a measurement
			int mode = mt->host_rate[1] <= 0 ? 1 : mt->host_rate[0] <= 0 ? 0 : (mt->host_rate[1] >= mt->host_rate[0] ? 1 : 0);
			if (mt->host_rate[0] > 0 && mt->host_rate[1] > 0 && mt->host_calls % 16 == 15)
				mode ^= 1;
			const uint64_t allocs = mt->host_allocs;
			timespec t0, t1;
			clock_gettime(CLOCK_MONOTONIC, &t0);
			rc = mode ? search_host_packed(mt, text, n, count, positions, cap, n_written, want_positions)
					  : search_host_raw(mt, text, n, count, positions, cap, n_written, want_positions);
			clock_gettime(CLOCK_MONOTONIC, &t1);
			const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
			if ((rc == ACWM_OK || rc == ACWM_ERR_OVERFLOW) && secs > 0 && mt->host_allocs == allocs) {
				// the plain copy is one thing (smoothed); the hybrid transfer changes as its raw share climbs: its last call counts
				const double rate = (double) n / secs;
				mt->host_rate[mode] = (mode == 0 && mt->host_rate[0] > 0) ? 0.5 * mt->host_rate[0] + 0.5 * rate : rate;
				mt->host_calls++;
			}
			return rc;
		}
	}
	return search_host_raw(mt, text, n, count, positions, cap, n_written, want_positions);
}

// The text crosses the link one byte per symbol: chunks of 32 MiB, copy of chunk i under the scan of chunk i - 1.
static int search_host_raw(acwm_matcher *mt, const uint8_t *text, uint64_t n, uint64_t *count, uint64_t *positions, uint64_t cap,
		uint64_t *n_written, int want_positions) {
	int rc;
	if (n + 64 > mt->text_cap) {
		if (mt->d_text)
			cudaFree(mt->d_text);
		mt->d_text = nullptr;
		mt->text_cap = 0;
		mt->host_allocs++;
		CU(cudaMalloc((void **) &mt->d_text, n + 64));
		mt->text_cap = n + 64;
	}
	const uint64_t T = kTile;
	const uint64_t n_tiles = (n + T - 1) / T;
	if (want_positions && (rc = ensure_tiles(mt, n_tiles)))
		return rc;
	// chunks of ~32 MiB, whole tiles each: H2D on s_copy, scan on s_scan as soon as the chunk
	// (and, through stream order, everything before it -- the halo) has landed; every launch
	// after the first appends its count and its (sorted) positions to the result block
	const uint64_t chunk_tiles = std::max<uint64_t>(1, (32ull << 20) / T);
	const uint64_t n_chunks = std::max<uint64_t>(1, (n_tiles + chunk_tiles - 1) / chunk_tiles);
	while (mt->ev_time.size() < 2 * n_chunks) {
		cudaEvent_t e;
		CU(cudaEventCreate(&e));
		mt->ev_time.push_back(e);
	}
	apply_l2_window(mt, mt->s_scan);
	mt->first_epoch = mt->epoch;
	for (uint64_t ci = 0; ci < n_chunks; ci++) {
		const uint64_t b0 = std::min<uint64_t>(n, ci * chunk_tiles * T), b1 = std::min<uint64_t>(n, (ci + 1) * chunk_tiles * T);
		if (b1 > b0)
			CU(cudaMemcpyAsync(mt->d_text + b0, text + b0, b1 - b0, cudaMemcpyHostToDevice, mt->s_copy));
		cudaEvent_t ev = mt->ev_copy[ci % mt->ev_copy.size()];
		CU(cudaEventRecord(ev, mt->s_copy));
		CU(cudaStreamWaitEvent(mt->s_scan, ev, 0));
		CU(cudaEventRecord(mt->ev_time[2 * ci], mt->s_scan));
		if ((rc = launch_scan(mt, mt->d_text, n, mt->host_report_from, std::min(n_tiles, ci * chunk_tiles),
					 std::min(n_tiles, (ci + 1) * chunk_tiles), want_positions, ci > 0, 0, mt->s_scan)))
			return rc;
		CU(cudaEventRecord(mt->ev_time[2 * ci + 1], mt->s_scan));
	}
	mt->last_want_positions = want_positions;
	mt->last_h2d_bytes = n;
	rc = acwm_fetch(mt, count, positions, cap, n_written, mt->s_scan);
	double secs = 0;
	for (uint64_t ci = 0; ci < n_chunks; ci++) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, mt->ev_time[2 * ci], mt->ev_time[2 * ci + 1]) == cudaSuccess)
			secs += ms * 1e-3;
	}
	mt->last_kernel_s = secs;
	return rc;
}

double acwm_last_kernel_seconds(const acwm_matcher *mt) { return mt ? mt->last_kernel_s : 0.0; }
uint64_t acwm_last_h2d_bytes(const acwm_matcher *mt) { return mt ? mt->last_h2d_bytes : 0; }

int acwm_pack_text_2bit(const uint8_t *text, uint64_t n, uint8_t *packed, int *bad_text) {
	if ((!text && n) || !packed)
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	static HostPacker pool; // shared by callers of this entry point; matchers own theirs
	static std::mutex mu;
	std::lock_guard<std::mutex> g(mu);
	const uint64_t bad = pool.pack(text, packed, n);
	if (bad_text)
		*bad_text = (bad & 0xFCFCFCFCFCFCFCFCull) != 0;
	return ACWM_OK;
}

int acwm_set_overlap(acwm_matcher *mt, int on) {
	if (!mt)
		return set_error(ACWM_ERR_INVALID, "matcher == NULL");
	mt->overlap = on != 0;
	return ACWM_OK;
}

int acwm_set_peers(acwm_matcher *mt, uint32_t rank, uint32_t world, const uint64_t *mailboxes) {
	if (!mt)
		return set_error(ACWM_ERR_INVALID, "matcher == NULL");
	if (world <= 1 || !mailboxes) {
		mt->peer_world = mt->peer_rank = 0;
		return ACWM_OK;
	}
	if (world > kMaxPeers || rank >= world)
		return set_error(ACWM_ERR_INVALID, "acwm_set_peers: world > 16 or rank >= world");
	for (uint32_t r = 0; r < world; r++) {
		if (!mailboxes[r])
			return set_error(ACWM_ERR_INVALID, "acwm_set_peers: NULL mailbox");
		mt->peer_ptrs[r] = mailboxes[r];
	}
	mt->peer_world = world;
	mt->peer_rank = rank;
	mt->xepoch = 0;
	return ACWM_OK;
}

int acwm_fetch_global_count(acwm_matcher *mt, uint64_t *global_count, void *stream) {
	if (!mt || !mt->uploaded || !global_count)
		return set_error(ACWM_ERR_INVALID, "matcher not uploaded / NULL argument");
	cudaStream_t st = (cudaStream_t) stream;
	if (mt->peer_world > 1 && mt->xepoch) { // the scan kernels run the exchange one scan behind: finish the last one
		collect_last_kernel<<<1, 1, 0, st>>>(mt->d_ctl, reinterpret_cast<const unsigned long long *>(mt->peer_ptrs[mt->peer_rank]),
				mt->peer_world, mt->xepoch);
		CU(cudaGetLastError());
		mt->launches++;
	}
	CU(cudaMemcpyAsync(mt->h_res, &mt->d_ctl->result, sizeof(Result), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	*global_count = mt->h_res->global_count;
	if (mt->peer_world > 1 && mt->h_res->exchange_failed) {
		CU(cudaMemsetAsync(&mt->d_ctl->result.exchange_failed, 0, sizeof(unsigned int), st));
		return set_error(ACWM_ERR_CUDA, "count exchange: a peer's count did not arrive (it skipped a scan, failed before its launch, or died)");
	}
	return ACWM_OK;
}

int acwm_set_trace(acwm_matcher *mt, unsigned long long *d_trace) {
	if (!mt)
		return set_error(ACWM_ERR_INVALID, "matcher == NULL");
	mt->d_trace = d_trace;
	return ACWM_OK;
}

uint32_t acwm_trace_words_per_cta(void) { return kTraceWords; }

int acwm_set_profiling(acwm_matcher *mt, int on) {
	if (!mt)
		return set_error(ACWM_ERR_INVALID, "matcher == NULL");
	mt->profiling = on != 0;
	return ACWM_OK;
}

int acwm_profiled_seconds(acwm_matcher *mt, double *scan_s, double *finalize_s) {
	if (!mt || !mt->profiling || !mt->ev_prof[0])
		return set_error(ACWM_ERR_INVALID, "profiling is off or no scan was profiled");
	CU(cudaEventSynchronize(mt->ev_prof[1]));
	float a = 0, b = 0; // the ordering of the positions is fused into the scan kernel: no separate finalize time
	CU(cudaEventElapsedTime(&a, mt->ev_prof[0], mt->ev_prof[1]));
	if (scan_s)
		*scan_s = a * 1e-3;
	if (finalize_s)
		*finalize_s = b * 1e-3;
	return ACWM_OK;
}

unsigned long long acwm_launch_count(const acwm_matcher *mt) { return mt ? mt->launches : 0; }

int acwm_get_info(const acwm_matcher *mt, acwm_info *info) {
	if (!mt || !info)
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	*info = mt->c.info;
	return ACWM_OK;
}

void acwm_free(acwm_matcher *mt) {
	if (!mt)
		return;
	if (mt->uploaded) {
		cudaSetDevice(mt->device);
		cudaFree(mt->d_tables);
		cudaFree(mt->d_ctl);
		cudaFree(mt->d_cta_total);
		if (mt->d_staging)
			cudaFree(mt->d_staging);
		if (mt->d_positions)
			cudaFree(mt->d_positions);
		if (mt->d_tile_count)
			cudaFree(mt->d_tile_count);
		if (mt->d_text)
			cudaFree(mt->d_text);
		if (mt->h_res)
			cudaFreeHost(mt->h_res);
		if (mt->h_bounce)
			cudaFreeHost(mt->h_bounce);
		if (mt->h_pack_ring)
			cudaFreeHost(mt->h_pack_ring);
		if (mt->h_hist)
			cudaFreeHost(mt->h_hist);
		if (mt->d_raw)
			cudaFree(mt->d_raw);
		if (mt->d_mailbox)
			cudaFree(mt->d_mailbox);
		if (mt->s_copy2)
			cudaStreamDestroy(mt->s_copy2);
		for (auto e : mt->ev_pack)
			if (e)
				cudaEventDestroy(e);
		for (auto e : mt->ev_hyb)
			if (e)
				cudaEventDestroy(e);
		for (auto e : mt->ev_bounce)
			if (e)
				cudaEventDestroy(e);
		if (mt->h_bounce2)
			cudaFreeHost(mt->h_bounce2);
		if (mt->s_copy)
			cudaStreamDestroy(mt->s_copy);
		if (mt->s_scan)
			cudaStreamDestroy(mt->s_scan);
		for (auto e : mt->ev_copy)
			if (e)
				cudaEventDestroy(e);
		for (auto e : mt->ev_time)
			cudaEventDestroy(e);
		for (auto e : mt->ev_prof)
			if (e)
				cudaEventDestroy(e);
		(void) cudaGetLastError();
	}
	delete mt->packer;
	delete mt;
}

void acwm_shard_bounds(uint64_t n, uint32_t world, uint32_t rank, uint32_t halo, uint64_t *start, uint64_t *len) {
	if (world == 0)
		world = 1;
	const uint64_t chunk = (n + world - 1) / world; // main.c:375: ceil(nFull / commSize)
	uint64_t s = (uint64_t) rank * chunk, e = (uint64_t) (rank + 1) * chunk + halo; // main.c:469-471
	if (e > n)
		e = n; // main.c:472-473
	if (s > n)
		s = n;
	if (start)
		*start = s;
	if (len)
		*len = e > s ? e - s : 0;
}

constexpr uint32_t kMaxShards = 64;

int acwm_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

// One host thread per shard: shard r = the MPI rank r of main.c:467-477, its matcher mts[r] (resident on whatever
// device the caller uploaded it to; not uploaded yet -> device r mod #devices), its slice of the caller's text.
int acwm_search_host_sharded(acwm_matcher *const *mts, uint32_t world, const uint8_t *text, uint64_t n, uint64_t *count,
		uint64_t *positions, uint64_t cap, uint64_t *n_written, uint64_t *shard_counts) {
	if (!mts || world == 0 || world > kMaxShards || (!text && n))
		return set_error(ACWM_ERR_INVALID, "acwm_search_host_sharded: NULL argument or world not in 1..64");
	uint32_t m_max = 0;
	for (uint32_t r = 0; r < world; r++) {
		if (!mts[r])
			return set_error(ACWM_ERR_INVALID, "acwm_search_host_sharded: NULL matcher");
		for (uint32_t q = 0; q < r; q++)
			if (mts[q] == mts[r])
				return set_error(ACWM_ERR_INVALID, "acwm_search_host_sharded: one matcher per shard (searches on a matcher are not concurrent)");
		if (r == 0)
			m_max = mts[r]->c.prm.m_max;
		else if (mts[r]->ps.len != mts[0]->ps.len || mts[r]->ps.bytes != mts[0]->ps.bytes)
			return set_error(ACWM_ERR_INVALID, "acwm_search_host_sharded: the matchers hold different pattern sets");
	}
	const int n_dev = acwm_device_count();
	if (n_dev <= 0)
		return set_error(ACWM_ERR_CUDA, "acwm_search_host_sharded: no CUDA device");
	const bool want_positions = positions != nullptr && cap > 0;
	struct Shard {
		uint64_t start = 0, len = 0, count = 0, written = 0;
		std::unique_ptr<uint64_t[]> pos;
		int rc = ACWM_OK;
		std::string err;
	};
	std::vector<Shard> sh(world);
	unsigned hw = std::thread::hardware_concurrency();
	if (hw == 0)
		hw = 4;
	auto run = [&](uint32_t r) {
		Shard &s = sh[r];
		acwm_matcher *mt = mts[r];
		acwm_shard_bounds(n, world, r, m_max ? m_max - 1 : 0, &s.start, &s.len);
		if (s.len < mt->c.prm.m_min)
			return; // an empty tail shard (n not much larger than world): nothing can end in it
		if (!mt->uploaded && (s.rc = do_upload(mt, (int) (r % (uint32_t) n_dev), 0))) {
			s.err = acwm_last_error();
			return;
		}
		if (world > 1 && !mt->packer) // the shards of one process share its cores
			mt->packer = new HostPacker(std::max(1u, hw / world));
		if (want_positions) {
			s.pos.reset(new (std::nothrow) uint64_t[cap]); // untouched pages cost nothing
			if (!s.pos) {
				s.rc = ACWM_ERR_NOMEM;
				s.err = "no host memory for the shard's positions";
				return;
			}
		}
		// every match is reported by exactly one shard: ends below m_max-1 of a shard but the first belong to
		// its predecessor (with equal-length patterns a plain scan does that by itself, main.c:467-477)
		mt->host_report_from = r ? (uint64_t) (m_max - 1) : 0;
		s.rc = acwm_search_host(mt, text + s.start, s.len, &s.count, s.pos.get(), want_positions ? cap : 0, &s.written);
		mt->host_report_from = 0;
		if (s.rc != ACWM_OK)
			s.err = acwm_last_error();
	};
	{
		// no exception leaves the C ABI: a shard whose thread cannot be started runs on this one
		auto guarded = [&](uint32_t r) {
			try {
				run(r);
			} catch (const std::exception &e) {
				sh[r].rc = ACWM_ERR_NOMEM;
				sh[r].err = e.what();
			}
		};
		std::vector<std::thread> th;
		th.reserve(world);
		std::vector<uint32_t> here{0};
		for (uint32_t r = 1; r < world; r++) {
			try {
				th.emplace_back(guarded, r);
			} catch (const std::exception &) {
				here.push_back(r);
			}
		}
		for (uint32_t r : here)
			guarded(r);
		for (auto &t : th)
			t.join();
	}
	// the MPI_Reduce(SUM) of main.c:656, and the gather of the positions: shards are in text order and each
	// shard's positions are sorted, so concatenation (+ shard start) is globally sorted
	uint64_t total = 0, w = 0;
	int rc = ACWM_OK;
	std::string err;
	struct Move {
		uint64_t *to;
		const uint64_t *from;
		uint64_t n, add;
	};
	std::vector<Move> moves;
	for (uint32_t r = 0; r < world; r++) {
		const Shard &s = sh[r];
		if (s.rc != ACWM_OK && s.rc != ACWM_ERR_OVERFLOW) {
			if (rc == ACWM_OK || rc == ACWM_ERR_OVERFLOW) {
				rc = s.rc;
				err = "shard " + std::to_string(r) + ": " + s.err;
			}
			continue;
		}
		if (s.rc == ACWM_ERR_OVERFLOW && rc == ACWM_OK) {
			rc = ACWM_ERR_OVERFLOW;
			err = "shard " + std::to_string(r) + ": " + s.err;
		}
		total += s.count;
		if (shard_counts)
			shard_counts[r] = s.count;
		if (want_positions) {
			const uint64_t take = std::min<uint64_t>(s.written, cap - w);
			if (take) // (to, from, how many, what to add): copied below, one thread per shard when there is much to move
				moves.push_back({positions + w, s.pos.get(), take, s.start});
			w += take;
			if (take < s.count && rc == ACWM_OK) {
				rc = ACWM_ERR_OVERFLOW;
				err = "more matches than the positions buffer holds (the count is exact)";
			}
		}
	}
	auto move = [](const Move &mv) {
		for (uint64_t i = 0; i < mv.n; i++)
			mv.to[i] = mv.from[i] + mv.add;
	};
	if (w >= (1u << 20) && moves.size() > 1) {
		std::vector<std::thread> th;
		th.reserve(moves.size());
		move(moves[0]);
		for (size_t i = 1; i < moves.size(); i++) {
			try {
				th.emplace_back(move, moves[i]);
			} catch (const std::exception &) {
				move(moves[i]);
			}
		}
		for (auto &t : th)
			t.join();
	} else
		for (const Move &mv : moves)
			move(mv);
	if (count)
		*count = total;
	if (n_written)
		*n_written = w;
	return rc == ACWM_OK ? ACWM_OK : set_error(rc, err);
}

// ---- one process, several GPUs, device-resident shards: the count exchange of the scan kernels without torch / NCCL
static int check_shard_set(acwm_matcher *const *mts, uint32_t world, const char *who) {
	if (!mts || world == 0 || world > kMaxPeers)
		return set_error(ACWM_ERR_INVALID, std::string(who) + ": NULL argument or world not in 1..16");
	for (uint32_t r = 0; r < world; r++) {
		if (!mts[r] || !mts[r]->uploaded)
			return set_error(ACWM_ERR_INVALID, std::string(who) + ": every matcher must be uploaded (acwm_upload) first");
		for (uint32_t q = 0; q < r; q++)
			if (mts[q] == mts[r])
				return set_error(ACWM_ERR_INVALID, std::string(who) + ": one matcher per shard");
	}
	return ACWM_OK;
}

int acwm_peers_destroy(acwm_matcher *const *mts, uint32_t world) {
	if (!mts)
		return set_error(ACWM_ERR_INVALID, "acwm_peers_destroy: NULL argument");
	for (uint32_t r = 0; r < world; r++) {
		acwm_matcher *mt = mts[r];
		if (!mt)
			continue;
		mt->peer_world = mt->peer_rank = 0;
		mt->shard_stream = nullptr;
		if (mt->d_mailbox && mt->uploaded) {
			cudaSetDevice(mt->device);
			cudaStreamSynchronize(mt->s_scan);
			cudaFree(mt->d_mailbox);
			(void) cudaGetLastError();
		}
		mt->d_mailbox = nullptr;
	}
	return ACWM_OK;
}

int acwm_peers_create(acwm_matcher *const *mts, uint32_t world) {
	int rc = check_shard_set(mts, world, "acwm_peers_create");
	if (rc != ACWM_OK)
		return rc;
	acwm_peers_destroy(mts, world);
	for (uint32_t r = 0; r < world; r++)
		for (uint32_t q = 0; q < world; q++) {
			const int dr = mts[r]->device, dq = mts[q]->device;
			if (dr == dq)
				continue;
			int can = 0;
			CU(cudaDeviceCanAccessPeer(&can, dr, dq));
			if (!can)
				return set_error(ACWM_ERR_UNSUPPORTED, "acwm_peers_create: device " + std::to_string(dr)
						+ " cannot access the memory of device " + std::to_string(dq));
			CU(cudaSetDevice(dr));
			const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				return cuda_fail(e, "cudaDeviceEnablePeerAccess");
			(void) cudaGetLastError();
		}
	uint64_t ptrs[kMaxPeers] = {};
	const size_t box_bytes = (size_t) kPeerRing * world * sizeof(unsigned long long);
	for (uint32_t r = 0; r < world; r++) {
		acwm_matcher *mt = mts[r];
		CU(cudaSetDevice(mt->device));
		CU(cudaMalloc((void **) &mt->d_mailbox, box_bytes));
		CU(cudaMemset(mt->d_mailbox, 0, box_bytes));
		CU(cudaDeviceSynchronize());
		ptrs[r] = (uint64_t) (uintptr_t) mt->d_mailbox;
	}
	for (uint32_t r = 0; r < world; r++) {
		if ((rc = acwm_set_peers(mts[r], r, world, ptrs)))
			return rc;
		// matchers of one device share the stream of the first of them: their scans then run in rank order, so a
		// kernel that collects the previous exchange never waits for a kernel queued behind it on the same GPU
		mts[r]->shard_stream = mts[r]->s_scan;
		for (uint32_t q = 0; q < r; q++)
			if (mts[q]->device == mts[r]->device) {
				mts[r]->shard_stream = mts[q]->shard_stream;
				break;
			}
	}
	return ACWM_OK;
}

int acwm_scan_device_sharded(acwm_matcher *const *mts, uint32_t world, const uint8_t *const *d_shards,
		const uint64_t *shard_lens, int want_positions) {
	int rc = check_shard_set(mts, world, "acwm_scan_device_sharded");
	if (rc != ACWM_OK)
		return rc;
	if (!d_shards || !shard_lens)
		return set_error(ACWM_ERR_INVALID, "acwm_scan_device_sharded: NULL argument");
	const uint32_t m_max = mts[0]->c.prm.m_max;
	for (uint32_t r = 0; r < world; r++) {
		acwm_matcher *mt = mts[r];
		if (world > 1 && (mt->peer_world != world || mt->peer_rank != r))
			return set_error(ACWM_ERR_INVALID, "acwm_scan_device_sharded: call acwm_peers_create on these matchers first");
		CU(cudaSetDevice(mt->device));
		cudaStream_t st = mt->shard_stream ? mt->shard_stream : mt->s_scan;
		if ((rc = acwm_scan_device(mt, d_shards[r], shard_lens[r], r ? (uint64_t) (m_max - 1) : 0, want_positions, st)))
			return rc;
	}
	return ACWM_OK;
}

int acwm_fetch_sharded(acwm_matcher *const *mts, uint32_t world, uint64_t *global_count, uint64_t *shard_counts) {
	int rc = check_shard_set(mts, world, "acwm_fetch_sharded");
	if (rc != ACWM_OK)
		return rc;
	for (uint32_t r = 0; r < world; r++) { // every scan has published before anybody collects
		CU(cudaSetDevice(mts[r]->device));
		CU(cudaStreamSynchronize(mts[r]->shard_stream ? mts[r]->shard_stream : mts[r]->s_scan));
	}
	uint64_t total = 0, first = 0;
	for (uint32_t r = 0; r < world; r++) {
		acwm_matcher *mt = mts[r];
		CU(cudaSetDevice(mt->device));
		uint64_t g = 0;
		if ((rc = acwm_fetch_global_count(mt, &g, mt->shard_stream ? mt->shard_stream : mt->s_scan)))
			return rc;
		if (mt->h_res->order_failed > mt->first_epoch)
			return set_error(ACWM_ERR_CUDA, "the position ordering gave up waiting for a CTA of its own launch");
		if (mt->h_res->bad_text)
			return set_error(ACWM_ERR_BAD_TEXT, "text holds a byte >= 4 but the matcher was built for alphabet <= 4");
		if (shard_counts)
			shard_counts[r] = mt->h_res->count;
		total += mt->h_res->count;
		if (r == 0)
			first = g;
		else if (g != first)
			return set_error(ACWM_ERR_CUDA, "count exchange: the shards disagree on the global count");
	}
	if (first != total)
		return set_error(ACWM_ERR_CUDA, "count exchange: the exchanged sum differs from the sum of the shard counts");
	if (global_count)
		*global_count = first;
	return ACWM_OK;
}

int acwm_text_to_device(int device, const uint8_t *text, uint64_t n, uint8_t **d_text) {
	if (!d_text || (!text && n))
		return set_error(ACWM_ERR_INVALID, "acwm_text_to_device: NULL argument");
	*d_text = nullptr;
	CU(cudaSetDevice(device));
	CU(cudaMalloc((void **) d_text, n ? n : 16));
	if (n)
		CU(cudaMemcpy(*d_text, text, n, cudaMemcpyHostToDevice));
	return ACWM_OK;
}

void acwm_device_free(int device, void *d_ptr) {
	if (!d_ptr)
		return;
	if (cudaSetDevice(device) == cudaSuccess)
		cudaFree(d_ptr);
	(void) cudaGetLastError();
}

int acwm_table_blob(const acwm_matcher *mt, int which, const void **ptr, uint64_t *bytes) {
	if (!mt || !ptr || !bytes)
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	const Compiled &c = mt->c;
	switch (which) {
	case ACWM_BLOB_FRONT: *ptr = c.front.data(); *bytes = c.front.size(); break;
	case ACWM_BLOB_FILTER2: *ptr = c.filter2.data(); *bytes = c.filter2.size() * 4; break;
	case ACWM_BLOB_BUCKET_START: *ptr = c.bucket_start.data(); *bytes = c.bucket_start.size() * 4; break;
	case ACWM_BLOB_ENTRIES: *ptr = c.entries.data(); *bytes = c.entries.size() * sizeof(acwm_ventry); break;
	case ACWM_BLOB_PATTERNS: *ptr = mt->ps.bytes.data(); *bytes = mt->ps.bytes.size(); break;
	case ACWM_BLOB_PARAMS: *ptr = &c.prm; *bytes = sizeof(c.prm); break;
	case ACWM_BLOB_SYMCLASS: *ptr = c.symclass.data(); *bytes = c.symclass.size(); break;
	case ACWM_BLOB_RMASK: *ptr = c.rmask.data(); *bytes = c.rmask.size(); break;
	case ACWM_BLOB_VDFA: *ptr = c.vdfa.data(); *bytes = c.vdfa.size() * 4; break;
	default: return set_error(ACWM_ERR_INVALID, "unknown blob id");
	}
	return ACWM_OK;
}

} // extern "C"
