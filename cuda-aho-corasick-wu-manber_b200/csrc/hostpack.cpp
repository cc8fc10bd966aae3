// Host-side 2-bit packer in front of the H2D copy of acwm_search_host (alphabet <= 4).
//
// A text on the host costs PCIe time, not HBM time: one byte per symbol as the reference stores it
// (ac/ac.c:209, wu/wu.c:63) crosses the link at ~50 GB/s, the scan kernel reads it at > 4000 GB/s.  Packing 4
// symbols per byte on the host cores (PEXT: 8 symbols -> 16 bits per instruction) quarters the bytes on the link;
// the kernel takes the packed tiles as they are (ScanArgs.packed_in: the layout is its own in-register format).
// The packer also checks the text (a byte >= 4 is ACWM_ERR_BAD_TEXT, as on the device path).
#include "hostpack.hpp"

#include <cstdlib>
#include <cstring>
#include <ctime>

#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace acwm {

namespace {

// 8 symbols (one per byte, little endian) -> 16 bits, symbol j at bits [2j, 2j+1]
inline uint32_t pack8_generic(uint64_t x) {
	x = (x | (x >> 6)) & 0x000F000F000F000Full;
	x = (x | (x >> 12)) & 0x000000FF000000FFull;
	x = (x | (x >> 24)) & 0xFFFFull;
	return (uint32_t) x;
}

uint64_t pack_generic(const uint8_t *src, uint8_t *dst, uint64_t n_sym) {
	uint64_t bad = 0, i = 0;
	for (; i + 8 <= n_sym; i += 8) {
		uint64_t x;
		memcpy(&x, src + i, 8);
		bad |= x;
		const uint16_t y = (uint16_t) pack8_generic(x);
		memcpy(dst + i / 4, &y, 2);
	}
	for (; i < n_sym; i += 4) { // tail: up to 7 symbols, zero-filled
		uint8_t b = 0;
		for (uint64_t k = 0; k < 4 && i + k < n_sym; k++) {
			bad |= src[i + k];
			b |= (uint8_t) ((src[i + k] & 3u) << (2 * k));
		}
		dst[i / 4] = b;
	}
	return bad;
}

#if defined(__x86_64__)
__attribute__((target("bmi2"))) uint64_t pack_bmi2(const uint8_t *src, uint8_t *dst, uint64_t n_sym) {
	uint64_t bad = 0, i = 0;
	const uint64_t m = 0x0303030303030303ull;
	for (; i + 32 <= n_sym; i += 32) { // 32 symbols -> one 64-bit store
		uint64_t x0, x1, x2, x3;
		memcpy(&x0, src + i, 8);
		memcpy(&x1, src + i + 8, 8);
		memcpy(&x2, src + i + 16, 8);
		memcpy(&x3, src + i + 24, 8);
		bad |= (x0 | x1) | (x2 | x3);
		const uint64_t y = _pext_u64(x0, m) | (_pext_u64(x1, m) << 16) | (_pext_u64(x2, m) << 32) | (_pext_u64(x3, m) << 48);
		memcpy(dst + i / 4, &y, 8);
	}
	if (i < n_sym)
		bad |= pack_generic(src + i, dst + i / 4, n_sym - i);
	return bad;
}
#endif

#if defined(__x86_64__)
// AVX-512: 64 symbols -> 16 bytes in three instructions (VPMADDUBSW pairs the symbols: b0 + 4 b1; VPMADDWD pairs the
// pairs: + 16 (b2 + 4 b3) = one byte per four symbols in every 32-bit lane; VPMOVDB narrows the lanes), against eight PEXT
// and their shifts; the packed bytes leave with non-temporal stores when the destination allows (the pinned ring is read by
// the DMA engine only: no read-for-ownership of lines the cores never look at again).
__attribute__((target("avx512f,avx512bw,avx512vl"))) uint64_t pack_avx512(const uint8_t *src, uint8_t *dst, uint64_t n_sym) {
	const __m512i w1 = _mm512_set1_epi16(0x0401), w2 = _mm512_set1_epi32(0x00100001);
	__m512i acc = _mm512_setzero_si512();
	uint64_t i = 0;
	const bool stream = (((uintptr_t) dst) & 63) == 0;
	for (; i + 256 <= n_sym; i += 256) { // 256 symbols -> one 64-byte line
		__m128i q[4];
		for (int k = 0; k < 4; k++) {
			const __m512i x = _mm512_loadu_si512((const void *) (src + i + 64 * k));
			acc = _mm512_or_si512(acc, x);
			q[k] = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(x, w1), w2));
		}
		const __m512i line = _mm512_inserti32x4(_mm512_inserti32x4(_mm512_inserti32x4(_mm512_castsi128_si512(q[0]), q[1], 1), q[2], 2), q[3], 3);
		if (stream)
			_mm512_stream_si512((__m512i *) (dst + i / 4), line);
		else
			_mm512_storeu_si512((void *) (dst + i / 4), line);
	}
	if (stream)
		_mm_sfence();
	uint64_t bad = (uint64_t) _mm512_reduce_or_epi64(acc);
	if (i < n_sym)
		bad |= pack_bmi2(src + i, dst + i / 4, n_sym - i);
	return bad;
}
#endif

using PackFn = uint64_t (*)(const uint8_t *, uint8_t *, uint64_t);
PackFn pick_pack() {
#if defined(__x86_64__)
	const char *e = getenv("ACWM_PACK_IMPL"); // generic | bmi2 | avx512 (measurements; default: the best the CPU has)
	if (e && !strcmp(e, "generic"))
		return pack_generic;
	const bool bmi2 = __builtin_cpu_supports("bmi2");
	if (bmi2 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && !(e && !strcmp(e, "bmi2")))
		return pack_avx512;
	if (bmi2)
		return pack_bmi2;
#endif
	return pack_generic;
}

constexpr uint64_t kPiece = HostPacker::kPieceSymbols; // symbols per work item (a multiple of 32)

} // namespace

HostPacker::HostPacker(unsigned threads) {
	fn_ = (void *) pick_pack();
	if (threads == 0) {
		const unsigned hw = std::thread::hardware_concurrency();
		threads = hw ? hw : 4;
		unsigned allowed = 0;
#if defined(__linux__)
		cpu_set_t set;
		if (sched_getaffinity(0, sizeof(set), &set) == 0)
			allowed = (unsigned) CPU_COUNT(&set);
#endif
		if (allowed && allowed < threads)
			threads = allowed; // the launcher gave this process its own cores (bench.py pins every rank to a slice)
		else if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
			// one process per GPU (torchrun), all on the same cores: the ranks of a box share them
			const unsigned w = (unsigned) strtoul(e, nullptr, 10);
			if (w > 1)
				threads = threads / w ? threads / w : 1;
		}
		if (threads > 4)
			threads -= 1; // leave a core to whatever else the process runs (the caller's own threads, the driver's)
		if (threads > 32)
			threads = 32;
	}
	for (unsigned t = 1; t < threads; t++) // the caller's thread is the first worker
		workers_.emplace_back([this] { worker(); });
}

HostPacker::~HostPacker() {
	{
		std::lock_guard<std::mutex> g(mu_);
		quit_ = true;
	}
	cv_.notify_all();
	for (auto &w : workers_)
		w.join();
}

// Pieces are claimed through one 64-bit ticket [job : 32 | next piece : 32]: a worker that wakes up late for a
// job that is already over can never take a piece of the next one with the old job's pointers.  A piece of chunk c
// may be claimed once c < gate (its ring slot is free).  Returns false when there is nothing to claim right now.
bool HostPacker::pack_one(const Job &j) {
	PackFn fn = (PackFn) fn_;
	for (;;) {
		uint64_t t = ticket_.load(std::memory_order_acquire);
		const uint64_t p = t & 0xffffffffu;
		if ((t >> 32) != (j.id & 0xffffffffu) || p >= j.pieces)
			return false;
		const uint64_t c = p / j.ppc;
		if (c >= gate_.load(std::memory_order_acquire))
			return false;
		inflight_.fetch_add(1, std::memory_order_acq_rel); // before the claim: finish() never sees a claimed piece it does not wait for
		if (!ticket_.compare_exchange_weak(t, t + 1, std::memory_order_acq_rel)) {
			inflight_.fetch_sub(1, std::memory_order_acq_rel);
			continue;
		}
		const uint64_t lo = p * kPiece, hi = lo + kPiece < j.n ? lo + kPiece : j.n;
		uint8_t *dst = j.ring + (c % j.ring_chunks) * j.slot_bytes + (p % j.ppc) * (kPiece / 4);
		const uint64_t bad = fn(j.src + lo, dst, hi - lo);
		if (bad)
			bad_.fetch_or(bad, std::memory_order_relaxed);
		done_[c % kMaxRing].fetch_add(1, std::memory_order_release);
		inflight_.fetch_sub(1, std::memory_order_acq_rel);
		return true;
	}
}

void HostPacker::worker() {
	uint64_t seen = 0;
	for (;;) {
		Job j;
		{
			std::unique_lock<std::mutex> g(mu_);
			cv_.wait(g, [&] { return quit_ || job_.id != seen; });
			if (quit_)
				return;
			j = job_;
			seen = j.id;
		}
		// stay on the job until its last piece is taken (gated pieces: the H2D copies are catching up)
		for (;;) {
			if (pack_one(j))
				continue;
			const uint64_t t = ticket_.load(std::memory_order_acquire);
			if ((t >> 32) != (j.id & 0xffffffffu) || (t & 0xffffffffu) >= j.pieces)
				break;
			std::this_thread::yield();
		}
		// Stay busy until the caller has its results (finish()) or the next job arrives: on the measured host the
		// DMA copies still in flight take ~0.5 ms longer once the packer cores go idle (profiles/README.md, session n)
		while (active_.load(std::memory_order_acquire) && seq_.load(std::memory_order_acquire) == seen)
			std::this_thread::yield();
	}
}

void HostPacker::begin(const uint8_t *src, uint64_t n_sym, uint64_t chunk_sym, uint8_t *ring, uint64_t slot_bytes,
		unsigned ring_chunks) {
	std::lock_guard<std::mutex> g(mu_);
	job_.id++;
	job_.src = src;
	job_.n = n_sym;
	job_.pieces = (n_sym + kPiece - 1) / kPiece;
	job_.ppc = chunk_sym / kPiece;
	job_.ring = ring;
	job_.slot_bytes = slot_bytes;
	job_.ring_chunks = ring_chunks < kMaxRing ? ring_chunks : kMaxRing;
	cur_ = job_;
	for (auto &d : done_)
		d.store(0, std::memory_order_relaxed);
	bad_.store(0, std::memory_order_relaxed);
	gate_.store(job_.ring_chunks, std::memory_order_relaxed);
	ticket_.store((job_.id & 0xffffffffu) << 32, std::memory_order_release);
	active_.store(true, std::memory_order_release);
	seq_.store(job_.id, std::memory_order_release);
	if (job_.pieces > 1)
		cv_.notify_all();
}

// End of a job, also on an error path of the caller with pieces still unclaimed: the ticket is voided (no worker claims
// another piece of this job, none spins on its gated pieces), and the pieces in flight are waited for -- when finish()
// returns nobody reads the caller's text or writes the ring any more.
void HostPacker::finish() {
	ticket_.store(~0ull, std::memory_order_release);
	while (inflight_.load(std::memory_order_acquire))
		std::this_thread::yield();
	active_.store(false, std::memory_order_release);
}

bool HostPacker::chunk_ready(uint64_t c) const {
	const uint64_t first = c * cur_.ppc, last = first + cur_.ppc < cur_.pieces ? first + cur_.ppc : cur_.pieces;
	return done_[c % kMaxRing].load(std::memory_order_acquire) >= last - first;
}

void HostPacker::wait_chunk(uint64_t c) {
	while (!chunk_ready(c))
		if (!pack_one(cur_))
			std::this_thread::yield();
}

void HostPacker::recycle(uint64_t c) {
	done_[c % kMaxRing].store(0, std::memory_order_relaxed);
	gate_.store(c + cur_.ring_chunks + 1, std::memory_order_release);
}

uint64_t HostPacker::pack_now(const uint8_t *src, uint8_t *dst, uint64_t n_sym) { return pick_pack()(src, dst, n_sym); }

uint64_t HostPacker::pack(const uint8_t *src, uint8_t *dst, uint64_t n_sym) {
	if (n_sym == 0)
		return 0;
	// one chunk that covers everything, written straight to dst
	const uint64_t chunk = ((n_sym + kPiece - 1) / kPiece) * kPiece;
	begin(src, n_sym, chunk, dst, 0, 1);
	wait_chunk(0);
	finish();
	return bad();
}

} // namespace acwm
