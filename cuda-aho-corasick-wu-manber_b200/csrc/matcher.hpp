// The matcher handle behind the C ABI (include/acwm.h).
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <string>
#include <vector>

#include "scan_common.cuh"
#include "tables.hpp"

namespace acwm {

constexpr uint32_t kMaxScanBlocks = 4096;
class HostPacker;

int set_error(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

} // namespace acwm

struct acwm_matcher {
	acwm::PatternSet ps;
	acwm::Compiled c;
	acwm_options opts{};
	// device residency
	bool uploaded = false;
	int device = -1, sm_count = 0;
	size_t l2_persist_max = 0, l2_window_max = 0;
	uint8_t *d_tables = nullptr; // one allocation: front | offset masks | stage-2 bitmap | buckets | entries | patterns
	size_t tables_bytes = 0;
	bool l2_tables = false;
	void *l2_window_stream = nullptr;
	uint8_t *d_front = nullptr;
	uint8_t *d_rmask = nullptr;
	uint32_t *d_filter2 = nullptr;
	uint32_t *d_bucket_start = nullptr;
	acwm_ventry *d_entries = nullptr;
	uint8_t *d_patterns = nullptr;
	uint32_t *d_vdfa = nullptr;
	acwm::Control *d_ctl = nullptr;
	acwm::Result *h_res = nullptr;
	uint64_t *h_bounce = nullptr; // pinned: the first positions of a result
	uint64_t *d_staging = nullptr, *d_positions = nullptr;
	uint64_t pos_cap = 0, stage_cap = 0;
	uint32_t *d_tile_count = nullptr;
	uint64_t tile_cap = 0;
	unsigned long long *d_cta_total = nullptr;
	// host-text pipeline
	uint8_t *d_text = nullptr;
	uint64_t text_cap = 0;
	cudaStream_t s_copy = nullptr, s_scan = nullptr;
	std::array<cudaEvent_t, 4> ev_copy{};
	acwm::HostPacker *packer = nullptr;      // 2-bit host packer (alphabet <= 4 host texts), created on first use
	uint8_t *h_pack_ring = nullptr;          // pinned ring the packer writes and the H2D copies read
	uint8_t *h_hist = nullptr;               // pinned: packed history in front of the packed part of a hybrid transfer
	uint8_t *d_raw = nullptr;                // device copy of the unpacked prefix of a hybrid transfer
	uint64_t raw_cap = 0;
	cudaStream_t s_copy2 = nullptr;          // its copy stream
	std::array<cudaEvent_t, 16> ev_pack{};    // one per ring slot: its copy is done
	uint64_t *h_bounce2 = nullptr;            // pinned: two slots of the pipelined fetch of a long position list
	std::array<cudaEvent_t, 2> ev_bounce{};
	std::array<cudaEvent_t, 4> ev_hyb{};      // hybrid transfer: start / end of the raw copies, start / end of the packed copies
	// acwm_search_host picks between the hybrid (packing) transfer and the plain one-byte-per-symbol copy by what
	// each DELIVERED on this box (text bytes per second of the whole call): eight ranks that share one box's memory
	// and PCIe root are served best by the plain copy, a single rank with all cores by the hybrid one
	double host_rate[2] = {0, 0};             // [0] plain copy, [1] hybrid; 0 = not tried yet
	uint32_t host_calls = 0;                  // measured calls
	uint64_t host_allocs = 0;                 // device / pinned allocations made by the host-text paths (a call that allocates is not a measurement)
	static constexpr int kRawTimes = 64;      // raw chunks the share may climb to (a 128 MiB text has 10 chunks)
	uint64_t raw_chunks = 0, raw_chunks_of = 0; // chunks of a pinned text sent unpacked in the next hybrid transfer / of a text of so many chunks
	uint32_t raw_calls = 0;
	double raw_time[kRawTimes] = {};          // smoothed duration of the searches that ran with R raw chunks (0: none yet)
	double raw_share = -1;                    // share of a pinned text sent unpacked, adapted from call to call (< 0: not yet measured)
	std::vector<cudaEvent_t> ev_time;
	std::array<cudaEvent_t, 2> ev_prof{};
	bool profiling = false;
	bool overlap = false;
	unsigned long long *d_trace = nullptr; // acwm_set_trace (caller-owned device buffer)
	uint32_t epoch = 0;
	uint32_t first_epoch = 0; // launch number of the first launch of the last search (acwm_fetch: did its ordering fail?)
	// multi-GPU count exchange (acwm_set_peers)
	uint32_t peer_world = 0, peer_rank = 0, xepoch = 0;
	uint64_t peer_ptrs[acwm::kMaxPeers] = {};
	unsigned long long *d_mailbox = nullptr; // acwm_peers_create: this matcher's mailbox (owned)
	cudaStream_t shard_stream = nullptr;     // acwm_scan_device_sharded: the stream of this matcher's device
	double last_kernel_s = 0;
	uint64_t last_h2d_bytes = 0; // bytes of text the last acwm_search_host sent over the link
	int last_want_positions = 0;
	uint64_t host_report_from = 0; // acwm_search_host_sharded: match ends below it belong to the shard in front
	unsigned long long launches = 0;
};
