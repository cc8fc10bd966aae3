// Thread pool that packs symbol-coded text (one byte per symbol, codes 0..3) to 2 bits per symbol; see hostpack.cpp.
#pragma once
#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace acwm {

class HostPacker {
public:
	static constexpr unsigned kMaxRing = 16;
	static constexpr uint64_t kPieceSymbols = 256 * 1024; // work item; chunks are whole multiples of it

	explicit HostPacker(unsigned threads = 0); // 0 = one per host core (shared between the ranks of a torchrun box)
	~HostPacker();
	HostPacker(const HostPacker &) = delete;
	HostPacker &operator=(const HostPacker &) = delete;

	// dst[i] = src[4i] | src[4i+1] << 2 | src[4i+2] << 4 | src[4i+3] << 6 for ceil(n_sym / 4) bytes (missing symbols = 0).
	// Returns the OR of all source bytes: a bit above the low two = a symbol >= 4 in the text.
	uint64_t pack(const uint8_t *src, uint8_t *dst, uint64_t n_sym);

	// The same on the calling thread alone (small pieces).
	static uint64_t pack_now(const uint8_t *src, uint8_t *dst, uint64_t n_sym);

	// Streaming form: the text is packed chunk by chunk (chunk_sym symbols, a multiple of kPieceSymbols) into a ring
	// of ring_chunks slots of slot_bytes; the workers run ahead of the caller, who takes the chunks in order:
	//   begin(); for c: wait_chunk(c) [helps packing meanwhile]; <copy slot c % ring>; recycle(c - ring + 1) once that
	//   slot's copy is done; bad() at the end.
	void begin(const uint8_t *src, uint64_t n_sym, uint64_t chunk_sym, uint8_t *ring, uint64_t slot_bytes, unsigned ring_chunks);
	bool chunk_ready(uint64_t c) const;
	void wait_chunk(uint64_t c);
	void recycle(uint64_t c); // the slot of chunk c is free again: chunk c + ring may be packed
	uint64_t bad() const { return bad_.load(std::memory_order_relaxed); }
	void finish(); // the caller has what it waited for: the workers may go to sleep
	unsigned threads() const { return (unsigned) workers_.size() + 1; }

private:
	struct Job {
		uint64_t id = 0;
		const uint8_t *src = nullptr;
		uint64_t n = 0, pieces = 0, ppc = 1; // symbols, work items, work items per chunk
		uint8_t *ring = nullptr;
		uint64_t slot_bytes = 0;
		unsigned ring_chunks = 1;
	};
	void worker();
	bool pack_one(const Job &j);

	void *fn_ = nullptr;
	std::vector<std::thread> workers_;
	std::mutex mu_;
	std::condition_variable cv_;
	bool quit_ = false;
	Job job_, cur_;
	std::atomic<uint64_t> ticket_{0}, gate_{0}, bad_{0}, seq_{0};
	std::atomic<uint32_t> inflight_{0}; // pieces claimed and not yet written: finish() waits for them
	std::atomic<bool> active_{false};
	std::array<std::atomic<uint64_t>, kMaxRing> done_{};
};

} // namespace acwm
