// Tile geometry shared by the table compiler (shared-memory budgeting) and the kernels.
//
// Persistent CTAs (one full-size or two half-size ones per SM); every WARP owns whole tiles of the text and keeps
// them RAW (one byte per symbol, as the reference stores the text) in one or two slots of shared memory, each filled
// by one TMA bulk copy per tile.  A lane reads its 112-byte chunk with 7 LDS.128: lane stride 112 B = 16 B * 7
// (odd) -> conflict-free.
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

// per-warp shared memory: ring of raw tiles, (2-bit path with a verification stage) 2-bit copy of the current tile
// for the verification windows, one mbarrier + one tile id per ring slot, match list, reservation log
constexpr uint32_t warp_smem_bytes(uint32_t stages, bool pk_copy) {
	return stages * kBufBytes + (pk_copy ? kPackWords * 4 : 0) + kMaxStages * 16 + kListCap * 2 + kLogCap * 8;
}

constexpr uint32_t kMaxSmem = 227 * 1024;
constexpr uint32_t kSmemReserve = 1024;             // CTA-level scratch (barriers, finalize scan), alignment slack
// Two CTAs per SM (the shape overlap mode prefers: consecutive scans share every SM, one scans while the other is in
// its prologue or its ordering epilogue): 228 KiB per SM, 1 KiB of it reserved per resident CTA.  A dual CTA asks for
// MORE than a third of the SM so that never three are resident (the in-flight bound behind the Work ring).
constexpr uint32_t kSmemPerSmTotal = 228 * 1024;
constexpr uint32_t kMaxSmemDual = kSmemPerSmTotal / 2 - 1024;
constexpr uint32_t kMinSmemDual = kSmemPerSmTotal / 3 - 1024 + 256;

// Staging entry: [tile:28 | rank:22 | pos_in_tile:14]
constexpr uint32_t kPosBits = 14, kRankBits = 22;
constexpr uint32_t kStageBlockLog2 = 7;
constexpr uint32_t kStageBlock = 1u << kStageBlockLog2; // staging slots of a warp's first reservation (each further one doubles)

// (warps, stages, CTAs per SM) the scan kernel is launched with, in order of preference.  The 2-bit path
// copies a tile into registers first and refills its slot while it walks, so one slot per
// warp already overlaps load and scan; the bytes path reads the raw tile throughout: two slots, or one and more
// warps.  The ring depth is a template parameter of the kernels (1 / 2).
struct LaunchShape {
	uint32_t warps, stages, ctas;
};
constexpr LaunchShape kShapesPacked[] = {{32, 1, 1}, {24, 1, 1}, {16, 1, 1}, {12, 1, 1}, {8, 1, 1}, {4, 1, 1}};
constexpr LaunchShape kShapesPackedDual[] = {{16, 1, 2}, {12, 1, 2}};
// bytes path: two raw slots per warp while 16 warps fit beside the tables; else one slot per warp and 20-24 warps
// (the load of a warp's next tile is then covered by the other warps instead of its own second slot: BASELINE
// configs[3], 128 KB of tables, runs 20 warps x 1 slot against 12 x 2)
constexpr LaunchShape kShapesBytes[] = {{16, 2, 1}, {24, 1, 1}, {20, 1, 1}, {12, 2, 1}, {16, 1, 1}, {8, 2, 1}, {4, 2, 1}};

inline uint32_t shape_smem(uint32_t table_bytes, LaunchShape s, bool pk_copy) {
	return table_bytes + s.warps * warp_smem_bytes(s.stages, pk_copy) + kSmemReserve;
}
inline bool shape_fits(uint32_t table_bytes, LaunchShape s, bool pk_copy) {
	return shape_smem(table_bytes, s, pk_copy) <= (s.ctas == 2 ? kMaxSmemDual : kMaxSmem);
}
// packed: the 2-bit path; pk_copy: its kernels keep a 2-bit copy of the tile (everything but the exact automaton)
inline LaunchShape shape_for_tables(uint32_t table_bytes, bool packed, bool pk_copy, bool dual = false) {
	if (packed && dual) {
		for (const LaunchShape &s : kShapesPackedDual)
			if (shape_fits(table_bytes, s, pk_copy))
				return s;
		return LaunchShape{0, 0, 0};
	}
	if (packed) {
		for (const LaunchShape &s : kShapesPacked)
			if (shape_fits(table_bytes, s, pk_copy))
				return s;
	} else {
		for (const LaunchShape &s : kShapesBytes)
			if (shape_fits(table_bytes, s, false))
				return s;
	}
	return LaunchShape{0, 0, 0};
}

} // namespace acwm
