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
constexpr uint32_t kPackWords = 1 + kTile / 16 + 3; // 2-bit copy of the tile (+16 symbols of history, + pad) = 228 words
constexpr uint32_t kMaxStages = 4;
constexpr uint32_t kMaxPeers = 16;                  // ranks of the in-kernel count exchange
constexpr uint32_t kPeerRing = 4;                   // mailbox slots per rank pair: epochs in flight
constexpr uint32_t kListCap = 128;                  // per-warp list of match positions of one tile (cooperative stores)
constexpr uint32_t kLogCap = 24;                    // per-warp log of its staging reservations (they double in size)
constexpr uint32_t kMaxGrabLog2 = 23;               // largest single reservation: 2^23 slots

// per-warp shared memory: ring of raw tiles, (2-bit path) 2-bit copy of the current tile for the
// verification windows, one mbarrier + one tile id per ring slot, match list, reservation log
constexpr uint32_t warp_smem_bytes(uint32_t stages, bool packed) {
	return stages * kBufBytes + (packed ? kPackWords * 4 : 0) + kMaxStages * 16 + kListCap * 2 + kLogCap * 8;
}

constexpr uint32_t kMaxSmem = 227 * 1024;
constexpr uint32_t kSmemReserve = 1024;             // CTA-level scratch (barriers, finalize scan), alignment slack

// Staging entry: [tile:28 | rank:22 | pos_in_tile:14]
constexpr uint32_t kPosBits = 14, kRankBits = 22;
constexpr uint32_t kStageBlockLog2 = 7;
constexpr uint32_t kStageBlock = 1u << kStageBlockLog2; // staging slots of a warp's first reservation (each further one doubles)

// (warps, stages) the scan kernel is launched with, in order of preference.  The 2-bit path
// copies a tile into registers first and refills its slot while it walks, so one slot per
// warp already overlaps load and scan; the bytes path reads the raw tile throughout.
struct LaunchShape {
	uint32_t warps, stages;
};
constexpr LaunchShape kShapesPacked[] = {{32, 1}, {24, 1}, {16, 2}, {16, 1}, {12, 2}, {12, 1}, {8, 2}, {8, 1}, {4, 2}, {4, 1}};
constexpr LaunchShape kShapesBytes[] = {{16, 2}, {12, 2}, {8, 2}, {4, 2}};

inline bool shape_fits(uint32_t table_bytes, LaunchShape s, bool packed) {
	return table_bytes + s.warps * warp_smem_bytes(s.stages, packed) + kSmemReserve <= kMaxSmem;
}
inline LaunchShape shape_for_tables(uint32_t table_bytes, bool packed) {
	if (packed) {
		for (const LaunchShape &s : kShapesPacked)
			if (shape_fits(table_bytes, s, true))
				return s;
	} else {
		for (const LaunchShape &s : kShapesBytes)
			if (shape_fits(table_bytes, s, false))
				return s;
	}
	return LaunchShape{0, 0};
}

} // namespace acwm
