// Tile geometry shared by the table compiler (shared-memory budgeting) and the kernels.
#pragma once
#include <cstdint>

namespace acwm {

// ---- 2-bit ("DNA", alphabet <= 4) path: one warp owns one tile ----
constexpr uint32_t kLaneSyms = 192;                 // symbols per lane: 48 B packed = 16 B * odd -> conflict-free LDS.128
constexpr uint32_t kWarpTile = 32 * kLaneSyms;      // 6144 symbols per warp tile
constexpr uint32_t kHaloSyms = 64;                  // history kept before the tile (>= D-1, >= B-1, >= 15)
constexpr uint32_t kTileWords = kWarpTile / 16;     // 384 packed words
constexpr uint32_t kHaloWords = kHaloSyms / 16;     // 4
constexpr uint32_t kBufWords = kHaloWords + kTileWords + 4; // + pad (funnel reads one word past the end)
constexpr uint32_t kQueueCap = 256;                 // candidate queue entries per warp (uint16)
constexpr uint32_t kWarpSmemPacked = 2 * kBufWords * 4 + kQueueCap * 2;

// ---- bytes path (alphabet > 4): raw text tiles in shared memory ----
constexpr uint32_t kLaneBytes = 112;                // 16 B * 7 -> conflict-free LDS.128
constexpr uint32_t kWarpTileB = 32 * kLaneBytes;    // 3584 bytes per warp tile
constexpr uint32_t kHaloBytes = 64;                 // history before the tile
constexpr uint32_t kBufBytesB = kHaloBytes + kWarpTileB + 16;
constexpr uint32_t kWarpSmemBytes = 2 * kBufBytesB + kQueueCap * 2 + 16 /* 2 mbarriers */;

constexpr uint32_t kMaxSmem = 227 * 1024;
constexpr uint32_t kSmemReserve = 1024;             // control words, alignment slack

// Staging entry: [tile:28 | rank:22 | pos_in_tile:13]
constexpr uint32_t kPosBits = 13, kRankBits = 22;

inline uint32_t threads_for_tables_packed(uint32_t table_bytes) {
	for (uint32_t warps : {32u, 24u, 16u, 8u})
		if (table_bytes + warps * kWarpSmemPacked + kSmemReserve <= kMaxSmem)
			return warps * 32;
	return 0;
}
inline uint32_t threads_for_tables_bytes(uint32_t table_bytes) {
	for (uint32_t warps : {16u, 12u, 8u, 4u})
		if (table_bytes + warps * kWarpSmemBytes + kSmemReserve <= kMaxSmem)
			return warps * 32;
	return 0;
}

} // namespace acwm
