// Host-side table compiler: pattern set -> scan tables (no CUDA in this file).
//
// Replaces the table-building half of the reference's preprocessing
// (/root/reference/ac/ac.c:224-245 preproc_ac, wu/wu.c:109-149 preproc_wu): the
// reference fills flat goto/failure/final arrays and dense SHIFT/PREFIX arrays that
// its kernels walk one symbol at a time; here the same pattern set is compiled into
//   * AC : a dense DFA with the failure function folded in, K symbols per lookup
//          on the 2-bit-compressed alphabet (alphabet <= 4) or one class-compressed
//          symbol per lookup (bytes path), end-anchored and depth-truncated when the
//          full automaton does not fit (hits are then verified);
//   * WM : a SHIFT-derived block bitmap sampled every s symbols (stage 1), a suffix
//          bitmap (HASH stage 2) and CSR verification buckets (PREFIX/compare).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/acwm.h"

namespace acwm {

struct PatternSet {
	uint32_t alphabet = 0;
	uint32_t p_in = 0;            // patterns handed in (with duplicates)
	uint32_t m_min = 0, m_max = 0;
	std::vector<uint8_t> bytes;   // distinct patterns back to back, first-occurrence order
	std::vector<uint64_t> off;
	std::vector<uint32_t> len;
	uint32_t size() const { return (uint32_t) len.size(); }
	const uint8_t *pat(uint32_t j) const { return bytes.data() + off[j]; }
};

struct Compiled {
	acwm_scan_params prm{};
	acwm_info info{};
	std::vector<uint8_t> front;          // AC: uint16 (smem) or uint32 (global) DFA entries; WM: uint32 bitmap words
	std::vector<uint8_t> rmask;          // WM, stride > 1: which offsets r < s a candidate block can sit at (uint8 / uint16)
	std::vector<uint32_t> filter2;       // stage-2 suffix bitmap
	std::vector<uint32_t> bucket_start;  // n_buckets + 1
	std::vector<acwm_ventry> entries;
	std::vector<uint8_t> symclass;       // bytes-path AC: 256 entries
	std::vector<uint32_t> vdfa;          // filtered AC: full-depth one-symbol DFA that decides candidate windows
	uint32_t front_entry_bytes = 2;
	// a second launch shape for the same tables (0 = none): one full-size CTA per SM where info names two half-size
	// ones.  Scans with a verification stage run faster that way on long texts (more warps per SM; per-launch
	// prologue and epilogue no longer matter there), api.cu picks per launch.
	uint32_t alt_threads = 0, alt_smem_bytes = 0;
};

// Geometry shared by the builder's cost model and the kernels.
constexpr uint32_t kSmemPerSM = 227 * 1024;
constexpr uint32_t kDefaultTableBudget = 128 * 1024;

int normalize_patterns(const uint8_t *patterns, const uint32_t *lens, uint32_t m, uint32_t p, uint32_t alphabet,
		PatternSet &out, std::string &err);

int compile_tables(int algo, const PatternSet &ps, const acwm_options &opts, Compiled &out, std::string &err);

// Packed-symbol helpers (2 bits per symbol, older symbol at lower bits).
inline uint32_t pack2_tail(const uint8_t *pat, uint32_t len, uint32_t nsym, uint32_t skip_from_end = 0) {
	uint32_t v = 0;
	for (uint32_t i = 0; i < nsym; i++)
		v |= (uint32_t) (pat[len - skip_from_end - nsym + i] & 3u) << (2 * i);
	return v;
}
// Byte helpers (8 bits per symbol, older symbol at lower bytes), up to 8 symbols.
inline uint64_t pack8_tail(const uint8_t *pat, uint32_t len, uint32_t nsym, uint32_t skip_from_end = 0) {
	uint64_t v = 0;
	for (uint32_t i = 0; i < nsym; i++)
		v |= (uint64_t) pat[len - skip_from_end - nsym + i] << (8 * i);
	return v;
}
// 64-bit block -> 32-bit mix used by the bytes-path stage-1 filter (device mirrors it).
inline uint32_t mix64to32(uint64_t v) {
	uint32_t lo = (uint32_t) v, hi = (uint32_t) (v >> 32);
	return lo * 0x9E3779B1u + hi * 0x85EBCA77u;
}

// Reference-layout flat tables for the shims (same content the reference leaves in
// its caller's arrays: ac/ac.c:61-62,114,162,186 and wu/wu.c:109-149).
void fill_reference_ac_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_supply, unsigned *state_final, unsigned *n_states, unsigned *n_distinct);
// Sibling algorithms behind the same matcher (same result set for equal-length patterns): the caller's flat tables
// as the reference's preproc_sh (sh/sh.c:78-149: trie of the REVERSED patterns), preproc_sbom (sbom/sbom.c:51-150:
// factor oracle of the reversed patterns, supply links resolved while the states are created, F(q) lists of 200
// cells per state) and preproc_sog8 (sog/sog8.c:113-170: 3-gram position masks, pattern hashes, two-level hash
// bitmap) leave them.
void fill_reference_sh_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_final, unsigned *n_states, unsigned *n_distinct);
void fill_reference_sbom_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_final_multi, unsigned *n_states, unsigned *n_patterns);
void fill_reference_sog8_tables(const uint8_t *const *rows, int p, uint8_t *T8, uint32_t *scanner_hs, int *scanner_index,
		uint8_t *scanner_hs2);
unsigned reference_wu_shiftsize(int alphabet);
void fill_reference_wu_tables(const uint8_t *const *rows, const uint8_t *flat, int m, int p, int B, int nbits,
		int *SHIFT, int *PREFIX_value, int *PREFIX_index, int *PREFIX_size);

} // namespace acwm
