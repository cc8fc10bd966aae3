// Tile geometry shared by the table compiler (shared-memory budgeting) and the kernels.
//
// One persistent CTA per SM; every WARP owns whole tiles of the text and double- or
// triple-buffers them RAW (one byte per symbol, as the reference stores the text) in
// shared memory, filled by one TMA bulk copy per tile.  A lane reads its 112-byte
// chunk with 7 LDS.128: lane stride 112 B = 16 B * 7 (odd) -> conflict-free.
#pragma once
#include <cstdint>

namespace acwm {

constexpr uint32_t kLane = 112;                     // symbols (= bytes) per lane per tile
constexpr uint32_t kTile = 32 * kLane;              // 3584 symbols per warp tile
constexpr uint32_t kHalo = 64;                      // history kept in front of the tile
constexpr uint32_t kLoadBytes = kHalo + kTile;      // one TMA bulk copy: 3648 B
constexpr uint32_t kBufBytes = kLoadBytes + 16;     // + pad read (never used) by the last window
constexpr uint32_t kMaxDepthPacked = 60;            // AC 2-bit path: warm-up (depth-1 + 2) must fit the halo
constexpr uint32_t kMaxDepthBytes = 64;             // AC bytes path: warm-up depth-1 <= 63
constexpr uint32_t kQueueCap = 256;                 // candidate queue entries per warp (uint16)
constexpr uint32_t kPackWords = 1 + kTile / 16 + 3; // 2-bit copy of the tile (+16 symbols of history, + pad) = 228 words
constexpr uint32_t kMaxStages = 4;

// per-warp shared memory: ring of raw tiles, 2-bit copy of the current tile (verification
// windows), candidate queue, one mbarrier per ring slot
constexpr uint32_t warp_smem_bytes(uint32_t stages) {
	return stages * kBufBytes + kPackWords * 4 + kQueueCap * 2 + kMaxStages * 8;
}

constexpr uint32_t kMaxSmem = 227 * 1024;
constexpr uint32_t kSmemReserve = 1024;             // CTA-level scratch (barriers, finalize scan), alignment slack

// Staging entry: [tile:28 | rank:22 | pos_in_tile:14]
constexpr uint32_t kPosBits = 14, kRankBits = 22;

// (warps, stages) the scan kernel is launched with, in order of preference.
struct LaunchShape {
	uint32_t warps, stages;
};
constexpr LaunchShape kShapes[] = {{16, 3}, {16, 2}, {12, 3}, {12, 2}, {8, 3}, {8, 2}, {4, 2}};

inline bool shape_fits(uint32_t table_bytes, LaunchShape s) {
	return table_bytes + s.warps * warp_smem_bytes(s.stages) + kSmemReserve <= kMaxSmem;
}
inline LaunchShape shape_for_tables(uint32_t table_bytes) {
	for (const LaunchShape &s : kShapes)
		if (shape_fits(table_bytes, s))
			return s;
	return LaunchShape{0, 0};
}

} // namespace acwm
