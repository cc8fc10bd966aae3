// Host-side table compiler (see tables.hpp).  Pure C++, no CUDA.
#include "tables.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "geometry.hpp"

namespace acwm {

// ------------------------------------------------------------------ pattern set
static uint64_t fnv1a(const uint8_t *p, size_t n) {
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; i++) {
		h ^= p[i];
		h *= 1099511628211ull;
	}
	return h;
}

int normalize_patterns(const uint8_t *patterns, const uint32_t *lens, uint32_t m, uint32_t p, uint32_t alphabet,
		PatternSet &out, std::string &err) {
	if (!patterns || p == 0) {
		err = "empty pattern set";
		return ACWM_ERR_INVALID;
	}
	if (alphabet < 2 || alphabet > 256) {
		err = "alphabet must be in [2, 256] (symbols are unsigned char codes, ac/ac.c:136)";
		return ACWM_ERR_INVALID;
	}
	if (!lens && m == 0) {
		err = "pattern length m must be >= 1";
		return ACWM_ERR_INVALID;
	}
	out = PatternSet();
	out.alphabet = alphabet;
	out.p_in = p;
	std::unordered_multimap<uint64_t, uint32_t> seen;
	seen.reserve((size_t) p * 2);
	uint64_t at = 0;
	for (uint32_t j = 0; j < p; j++) {
		const uint32_t L = lens ? lens[j] : m;
		if (L == 0) {
			err = "zero-length pattern";
			return ACWM_ERR_INVALID;
		}
		const uint8_t *s = patterns + at;
		at += L;
		for (uint32_t i = 0; i < L; i++)
			if (s[i] >= alphabet) {
				err = "pattern symbol >= alphabet";
				return ACWM_ERR_INVALID;
			}
		const uint64_t h = fnv1a(s, L) ^ ((uint64_t) L << 48);
		bool dup = false;
		auto range = seen.equal_range(h);
		for (auto it = range.first; it != range.second; ++it) {
			const uint32_t k = it->second;
			if (out.len[k] == L && memcmp(out.pat(k), s, L) == 0) {
				dup = true;
				break;
			}
		}
		if (dup)
			continue; // identical patterns collapse onto one terminal state (ac/ac.c:183)
		seen.emplace(h, out.size());
		out.off.push_back(out.bytes.size());
		out.len.push_back(L);
		out.bytes.insert(out.bytes.end(), s, s + L);
	}
	out.m_min = *std::min_element(out.len.begin(), out.len.end());
	out.m_max = *std::max_element(out.len.begin(), out.len.end());
	return ACWM_OK;
}

// ------------------------------------------------------------------ helpers
static uint32_t ceil_log2(uint64_t v) {
	uint32_t b = 0;
	while (((uint64_t) 1 << b) < v)
		b++;
	return b;
}

static const uint32_t kMultF1 = 0x9E3779B1u;  // stage-1 block hash
static const uint32_t kMultF2 = 0x85EBCA77u;  // stage-2 suffix hash
static const uint32_t kMultHB = 0xC2B2AE3Du;  // bucket hash
static const uint32_t kMultR = 0x27D4EB2Fu;   // offset-mask hash

static inline void set_bit(std::vector<uint32_t> &bm, uint32_t idx) { bm[idx >> 5] |= 1u << (idx & 31); }

// Key of a pattern for stage 2 / buckets: the last b2 symbols.
static uint32_t pattern_key(const PatternSet &ps, uint32_t j, bool packed, uint32_t b2) {
	if (packed)
		return pack2_tail(ps.pat(j), ps.len[j], b2);
	return mix64to32(pack8_tail(ps.pat(j), ps.len[j], b2));
}

// How much shared memory the tables of one CTA may take: a CTA that owns its SM, or one of two that share it.
struct TableProfile {
	uint32_t stage1_bytes;   // WM stage-1 bitmap
	uint32_t rmask_bytes;    // WM offset masks
	uint32_t f2_bits;        // log2 of the stage-2 bitmap (bits)
};
static const TableProfile kProfileSingle = {64 * 1024, 32 * 1024, 18};
static const TableProfile kProfileDual = {32 * 1024, 8 * 1024, 16};

// Stage 2 (suffix bitmap) + verification buckets, shared by WM and truncated AC.
static void build_verify(const PatternSet &ps, bool packed, const acwm_options &opts, Compiled &c,
		const TableProfile &prof = kProfileSingle) {
	acwm_scan_params &prm = c.prm;
	const uint32_t pd = ps.size();
	const uint32_t b2 = packed ? std::min<uint32_t>(ps.m_min, 16) : std::min<uint32_t>(ps.m_min, 8);
	prm.b2 = b2;
	// stage-2 bitmap: direct index when the key space is small, hashed otherwise
	const uint32_t key_bits = packed ? 2 * b2 : 32;
	// ~1.5 % of the probes pass (a pass costs dependent L2 loads in the buckets); a larger bitmap
	// measured no faster (profiles/r01d_tune.csv) and every CTA has to load it
	// shared memory up to 2^18 bits (32 KiB); larger sets get an L2-resident bitmap of up to 2^26 bits
	uint32_t f2bits = std::min<uint32_t>(std::max<uint32_t>(ceil_log2((uint64_t) pd * 64), 13),
			(opts.force_smem_tables || pd <= 16384) ? prof.f2_bits : 26);
	if (opts.force_f2_bits)
		f2bits = std::min<uint32_t>(std::max<uint32_t>(opts.force_f2_bits, 13), 26);
	if (packed && key_bits <= f2bits) {
		f2bits = std::max<uint32_t>(key_bits, 5);
		prm.f2_mult = 1;
		prm.f2_sh = 0;
	} else {
		prm.f2_mult = kMultF2;
		prm.f2_sh = 32 - f2bits;
	}
	prm.f2_in_smem = f2bits <= 18 ? 1 : 0; // (a forced size may exceed the profile: the launch shape decides then)
	prm.f2_words = (uint32_t) (((uint64_t) 1 << f2bits) / 32);
	c.filter2.assign(prm.f2_words, 0);
	// buckets
	const uint32_t hbits = std::max<uint32_t>(ceil_log2((uint64_t) pd * 2), 4);
	prm.hb_mult = kMultHB;
	prm.hb_sh = 32 - hbits;
	prm.n_buckets = 1u << hbits;
	prm.n_entries = pd;
	c.bucket_start.assign((size_t) prm.n_buckets + 1, 0);
	std::vector<uint32_t> key(pd), bkt(pd);
	for (uint32_t j = 0; j < pd; j++) {
		key[j] = pattern_key(ps, j, packed, b2);
		set_bit(c.filter2, (uint32_t) ((uint32_t) (key[j] * prm.f2_mult) >> prm.f2_sh));
		bkt[j] = (uint32_t) (key[j] * prm.hb_mult) >> prm.hb_sh;
		c.bucket_start[bkt[j] + 1]++;
	}
	for (uint32_t b = 0; b < prm.n_buckets; b++)
		c.bucket_start[b + 1] += c.bucket_start[b];
	c.entries.assign(pd, acwm_ventry{});
	std::vector<uint32_t> fill(c.bucket_start.begin(), c.bucket_start.end() - 1);
	for (uint32_t j = 0; j < pd; j++) {
		acwm_ventry &e = c.entries[fill[bkt[j]]++];
		e.key = key[j];
		e.len = ps.len[j];
		if (packed && ps.len[j] <= b2)
			e.len |= 0x80000000u; // the key is the whole pattern
		e.offset = ps.off[j];
	}
}

// ------------------------------------------------------------------ AC: suffix trie -> DFA
struct Trie {
	uint32_t A = 0, D = 0;
	std::vector<int32_t> go;     // [nodes * A]
	std::vector<uint8_t> depth;
	uint32_t rows = 0, leaves = 0;
	bool overflow = false;
	uint32_t nodes() const { return (uint32_t) depth.size(); }
};

// Trie of the last D symbols of every pattern (symbols mapped through cls[] when given).
static void build_suffix_trie(const PatternSet &ps, uint32_t D, uint32_t A, const uint8_t *cls, uint32_t max_rows,
		Trie &t) {
	t = Trie();
	t.A = A;
	t.D = D;
	t.go.assign(A, -1);
	t.depth.assign(1, 0);
	t.rows = 1;
	for (uint32_t j = 0; j < ps.size(); j++) {
		const uint8_t *s = ps.pat(j) + (ps.len[j] - D);
		uint32_t st = 0;
		for (uint32_t i = 0; i < D; i++) {
			const uint32_t c = cls ? cls[s[i]] : s[i];
			int32_t nx = t.go[(size_t) st * A + c];
			if (nx < 0) {
				nx = (int32_t) t.nodes();
				t.go[(size_t) st * A + c] = nx;
				t.go.insert(t.go.end(), A, -1);
				t.depth.push_back((uint8_t) (i + 1));
				if (i + 1 < D) {
					if (++t.rows > max_rows) {
						t.overflow = true;
						return;
					}
				} else
					t.leaves++;
			}
			st = (uint32_t) nx;
		}
	}
}

// One-step DFA over the rows (nodes of depth < D), failure function folded in.
// next1[row*A + c] = (next_row << 1) | hit, where hit = the transition reached a
// depth-D node (its successor row is that node's failure state, which behaves
// identically from then on because a depth-D node has no children).
static void build_dfa1(const Trie &t, std::vector<uint32_t> &next1, uint32_t &n_rows) {
	const uint32_t A = t.A, D = t.D, N = t.nodes();
	std::vector<int32_t> row(N, -1), fail(N, 0);
	std::vector<uint32_t> order;
	order.reserve(N);
	std::vector<uint32_t> dnode((size_t) N * A, 0); // full delta over nodes (targets may be leaves)
	order.push_back(0);
	n_rows = 0;
	for (size_t qi = 0; qi < order.size(); qi++) {
		const uint32_t u = order[qi];
		if (t.depth[u] >= D)
			continue; // leaf: no row, no outgoing goto
		row[u] = (int32_t) n_rows++;
		for (uint32_t c = 0; c < A; c++) {
			const int32_t v = t.go[(size_t) u * A + c];
			if (v >= 0) {
				fail[v] = (u == 0) ? 0 : (int32_t) dnode[(size_t) fail[u] * A + c];
				dnode[(size_t) u * A + c] = (uint32_t) v;
				order.push_back((uint32_t) v);
			} else
				dnode[(size_t) u * A + c] = (u == 0) ? 0 : dnode[(size_t) fail[u] * A + c];
		}
	}
	next1.assign((size_t) n_rows * A, 0);
	for (uint32_t u = 0; u < N; u++) {
		if (row[u] < 0)
			continue;
		for (uint32_t c = 0; c < A; c++) {
			uint32_t v = dnode[(size_t) u * A + c];
			uint32_t hit = 0;
			if (t.depth[v] >= D) {
				hit = 1;
				v = (uint32_t) fail[v];
			}
			next1[(size_t) row[u] * A + c] = ((uint32_t) row[v] << 1) | hit;
		}
	}
}

static uint32_t full_trie_states(const PatternSet &ps, uint32_t A_hint) {
	// idcounter of the reference automaton (smatcher.h:50): nodes of the prefix trie.
	(void) A_hint;
	std::unordered_map<uint64_t, uint32_t> edge; // (state << 8 | sym) -> child
	edge.reserve((size_t) ps.bytes.size() * 2);
	uint32_t n = 1;
	for (uint32_t j = 0; j < ps.size(); j++) {
		uint32_t st = 0;
		const uint8_t *s = ps.pat(j);
		for (uint32_t i = 0; i < ps.len[j]; i++) {
			auto it = edge.find(((uint64_t) st << 8) | s[i]);
			if (it == edge.end()) {
				edge.emplace(((uint64_t) st << 8) | s[i], n);
				st = n++;
			} else
				st = it->second;
		}
	}
	return n;
}

struct WmPlan;
static double wm_plan_cost(const PatternSet &ps, bool packed, uint32_t stage1_bytes, bool allow_global);
static int compile_wm(const PatternSet &ps, const acwm_options &opts, bool packed, const TableProfile &prof, Compiled &c,
		std::string &err);

// The automaton that decides a candidate window of the filtered AC front (front_kind 1): the full-depth DFA of the
// pattern set, failure function folded in, one symbol per lookup, uint32 entries (next row << 1 | final) in global
// memory.  Walked from the root over the m symbols of a window it ends on a final state iff the window is a pattern.
static bool build_verify_dfa(const PatternSet &ps, uint32_t A, const uint8_t *cls, Compiled &c) {
	const uint64_t kMaxBytes = 96ull << 20;
	Trie t;
	build_suffix_trie(ps, ps.m_min, A, cls, (uint32_t) (kMaxBytes / (4 * A)), t);
	if (t.overflow)
		return false;
	uint32_t n_rows = 0;
	build_dfa1(t, c.vdfa, n_rows);
	c.prm.v_rows = n_rows;
	return true;
}

static int compile_ac_packed(const PatternSet &ps, const acwm_options &opts, uint32_t budget, const TableProfile &prof,
		Compiled &c, std::string &err) {
	acwm_scan_params &prm = c.prm;
	const uint32_t m = ps.m_min;
	const uint32_t Dmax = std::min<uint32_t>(m, kMaxDepthPacked);
	// cost model (lane-instruction slots per symbol): 7 per DFA lookup, ~100 per candidate (checked by its own
	// lane while the rest of the warp waits)
	double best_cost = 1e30;
	uint32_t bestK = 0, bestD = 0;
	uint32_t d_lo = 1, d_hi = Dmax;
	if (opts.force_depth) {
		d_lo = d_hi = std::min<uint32_t>(std::max<uint32_t>(opts.force_depth, 1), Dmax);
	}
	// entry = byte offset of the next row | hit bits, 16 bits: rows * row_bytes < 65536
	const uint32_t max_rows_any = std::min<uint32_t>(budget / 8, (1u << 13) - 1); // K = 1
	for (uint32_t D = d_hi; D >= d_lo; D--) {
		Trie t;
		build_suffix_trie(ps, D, 4, nullptr, max_rows_any, t);
		if (t.overflow)
			continue;
		const bool exact = (D == m);
		const double rate = exact ? 0.0 : std::min(1.0, (double) t.leaves / std::pow(4.0, (double) D));
		for (uint32_t K = 3; K >= 1; K--) {
			if (opts.force_stride && opts.force_stride != K)
				continue;
			const uint64_t bytes = (uint64_t) t.rows << (2 * K + 1);
			if (bytes > budget || bytes >= 65536)
				continue;
			const double cost = 7.0 / K + rate * 100.0;
			if (cost < best_cost - 1e-9) {
				best_cost = cost;
				bestK = K;
				bestD = D;
			}
		}
	}
	// When no automaton that fits shared memory filters well (large sets: every short suffix occurs), the
	// automaton goes to global memory: uint32 entries, K <= 2, served from L2 (access-policy window).
	// One dependent L2 lookup per K symbols (~40 slots) against a candidate rate that drops 4x per level.
	bool global = false;
	const uint64_t kGlobalBytesMax = 48ull << 20;
	if (!opts.force_smem_tables && (best_cost > 12.0 || !bestK)) {
		for (uint32_t D = std::min<uint32_t>(d_hi, 16); D >= std::max<uint32_t>(d_lo, 2); D--) {
			Trie t;
			build_suffix_trie(ps, D, 4, nullptr, (uint32_t) (kGlobalBytesMax / 16), t);
			if (t.overflow)
				continue;
			const bool exact = (D == m);
			const double rate = exact ? 0.0 : std::min(1.0, (double) t.leaves / std::pow(4.0, (double) D));
			for (uint32_t K = 2; K >= 1; K--) {
				if (opts.force_stride && opts.force_stride != K)
					continue;
				if (((uint64_t) t.rows << (2 * K + 2)) > kGlobalBytesMax)
					continue;
				const double cost = 40.0 / K + rate * 100.0;
				if (cost < best_cost - 1e-9) {
					best_cost = cost;
					bestK = K;
					bestD = D;
					global = true;
				}
			}
		}
	}
	// A sampled block filter in front of the automaton (front_kind 1): the stage-1 bitmap of the WM compiler says where
	// a pattern CAN end, and only those windows are walked through the (global, L2-resident) automaton.  What large
	// sets need -- a dependent L2 lookup per K symbols is 6x slower than the filter (profiles/README.md, c3 vs c3wm) --
	// and what every set too large for a K = 3 automaton in shared memory gains from.
	if (opts.force_front != 1 && !opts.force_depth && m >= 3) {
		const double fcost = wm_plan_cost(ps, true, std::min<uint32_t>(prof.stage1_bytes, budget), !opts.force_smem_tables) + 0.25;
		if (opts.force_front == 2 || fcost < best_cost - 1e-9) {
			acwm_options wo = opts;
			wo.force_stride = 0; // the AC meaning of force_stride (K) does not apply to the filter
			int rc = compile_wm(ps, wo, true, prof, c, err);
			if (rc != ACWM_OK)
				return rc;
			prm.front_kind = 1;
			// What decides a candidate window: the exact compare of the hash buckets.  A walk of the full-depth automaton
			// (ACWM_VERIFY_DFA=1) gives the same answer in m DEPENDENT lookups from L2 -- ~10 us for one lane at m = 32,
			// and a random text of 128 MiB holds hundreds of windows that share their last 16 symbols with one of
			// 100 000 patterns: the stragglers cost a quarter of the scan (profiles/README.md, session l).
			prm.verify_kind = (getenv("ACWM_VERIFY_DFA") && atoi(getenv("ACWM_VERIFY_DFA")) && build_verify_dfa(ps, 4, nullptr, c)) ? 1 : 0;
			return ACWM_OK;
		}
	}
	if (!bestK) {
		err = "AC: no (stride, depth) fits the table budget";
		return ACWM_ERR_UNSUPPORTED;
	}
	Trie t;
	build_suffix_trie(ps, bestD, 4, nullptr, global ? (uint32_t) (kGlobalBytesMax / 16) : max_rows_any, t);
	std::vector<uint32_t> next1;
	uint32_t n_rows = 0;
	build_dfa1(t, next1, n_rows);
	const uint32_t K = bestK, cols = 1u << (2 * K);
	const uint32_t eb = global ? 4 : 2; // entry = byte offset of the next row | hit bits
	c.front_entry_bytes = eb;
	c.front.assign((size_t) n_rows * cols * eb, 0);
	for (uint32_t r = 0; r < n_rows; r++)
		for (uint32_t idx = 0; idx < cols; idx++) {
			uint32_t st = r, hits = 0;
			for (uint32_t i = 0; i < K; i++) {
				const uint32_t e = next1[(size_t) st * 4 + ((idx >> (2 * i)) & 3)];
				hits |= (e & 1) << i;
				st = e >> 1;
			}
			// the hit bits sit above the entry's log2(entry bytes) clear low bits, i.e. where the symbols of the next
			// lookup go: the kernel forms the next address with one bitwise select (scan_packed.cu, FrontAC::step_of)
			if (global)
				reinterpret_cast<uint32_t *>(c.front.data())[(size_t) r * cols + idx] = (st * cols * 4) | (hits << 2);
			else
				reinterpret_cast<uint16_t *>(c.front.data())[(size_t) r * cols + idx] =
						(uint16_t) ((st << (2 * K + 1)) | (hits << 1));
		}
	prm.stride = K;
	prm.depth = bestD;
	prm.exact_front = (bestD == m) ? 1 : 0;
	prm.n_rows = n_rows;
	// two chains per lane (scan_packed.cu): the second chain repeats ceil((depth - 1) / K) warm-up lookups
	prm.ilp = (K == 3 && !global && prm.exact_front && bestD <= 13 && getenv("ACWM_NO_ILP") == nullptr) ? 2 : 1;
	if (!prm.exact_front)
		build_verify(ps, true, opts, c, prof);
	c.info.table_in_smem = global ? 0 : 1;
	return ACWM_OK;
}

static int compile_ac_bytes(const PatternSet &ps, const acwm_options &opts, uint32_t budget, Compiled &c,
		std::string &err) {
	acwm_scan_params &prm = c.prm;
	const uint32_t m = ps.m_min;
	// columns: symbol codes clamped to `alphabet` (one extra class for out-of-alphabet bytes)
	uint32_t ncols = 1;
	while (ncols < std::min<uint32_t>(ps.alphabet + 1, 256))
		ncols <<= 1;
	// class(b) = min(b, alphabet): every out-of-alphabet text byte shares the one class no
	// pattern uses (the kernel computes the same min arithmetically; this copy is for tests)
	c.symclass.resize(256);
	for (uint32_t b = 0; b < 256; b++)
		c.symclass[b] = (uint8_t) std::min<uint32_t>(b, std::min<uint32_t>(ps.alphabet, 255));
	const uint32_t Dmax = std::min<uint32_t>(m, kMaxDepthBytes);
	double best_cost = 1e30;
	uint32_t bestD = 0;
	bool best_smem = true;
	uint32_t d_lo = 1, d_hi = std::min<uint32_t>(Dmax, 8);
	if (m <= kMaxDepthBytes && m > d_hi)
		d_hi = m; // also try the exact automaton
	if (opts.force_depth)
		d_lo = d_hi = std::min<uint32_t>(std::max<uint32_t>(opts.force_depth, 1), Dmax);
	// rows are bounded by the 2^31-byte table limit below BEFORE the trie is built: a byte-alphabet trie costs 4 * ncols
	// bytes per node, so an unbounded one could take GiBs per depth tried
	const uint32_t max_rows_global = (uint32_t) std::min<uint64_t>(1u << 22, (1ull << 31) / (4ull * ncols));
	for (uint32_t D = d_hi; D >= d_lo; D--) {
		if (D > 8 && D != m && !opts.force_depth)
			continue;
		Trie t;
		build_suffix_trie(ps, D, ncols, c.symclass.data(), max_rows_global, t);
		if (t.overflow)
			continue;
		const bool exact = (D == m);
		const double rate = exact ? 0.0 : std::min(1.0, (double) t.leaves / std::pow((double) ps.alphabet, (double) D));
		if (!exact && rate > 0.25 && D < d_hi && !opts.force_depth)
			continue; // a front that lets a quarter of the positions through filters nothing: the divergent verify path would run the scan
		const uint64_t smem_bytes = (uint64_t) t.rows * ncols * 2;
		const bool fits = smem_bytes <= budget && t.rows < (1u << 15);
		// one lookup per symbol: ~8 lane-instructions from shared memory, ~40 from L2
		const double cost = (fits ? 8.0 : 40.0) + rate * 30.0;
		if (cost < best_cost - 1e-9 && ((uint64_t) t.rows * ncols * 4 <= (1ull << 31))) {
			best_cost = cost;
			bestD = D;
			best_smem = fits;
		}
	}
	if (!bestD) {
		err = "AC: automaton too large";
		return ACWM_ERR_UNSUPPORTED;
	}
	Trie t;
	build_suffix_trie(ps, bestD, ncols, c.symclass.data(), max_rows_global, t);
	std::vector<uint32_t> next1;
	uint32_t n_rows = 0;
	build_dfa1(t, next1, n_rows);
	if (best_smem) {
		c.front_entry_bytes = 2;
		c.front.assign((size_t) n_rows * ncols * 2, 0);
		uint16_t *tab = reinterpret_cast<uint16_t *>(c.front.data());
		for (size_t i = 0; i < (size_t) n_rows * ncols; i++)
			tab[i] = (uint16_t) next1[i];
	} else {
		c.front_entry_bytes = 4;
		c.front.assign((size_t) n_rows * ncols * 4, 0);
		memcpy(c.front.data(), next1.data(), c.front.size());
	}
	prm.stride = 1;
	prm.depth = bestD;
	prm.exact_front = (bestD == m) ? 1 : 0;
	prm.n_rows = n_rows;
	prm.n_classes = ncols;
	if (!prm.exact_front)
		build_verify(ps, false, opts, c);
	c.info.table_in_smem = best_smem ? 1 : 0;
	return ACWM_OK;
}

// ------------------------------------------------------------------ WM
// Stage 1 = the SHIFT table of Wu-Manber evaluated at a fixed stride s instead of a
// data-dependent skip: a sample position c is a candidate iff some pattern holds the
// block text[c-B+1..c] at distance r < s from its end, i.e. iff SHIFT[block] < s
// (wu/wu.c:126-128 computes the same minimum distance).  Every occurrence ending at
// e is covered by the sample c = e - r, r = e mod-aligned distance < s, because
// B <= m_min - s + 1 keeps the block inside the occurrence.
// Where a WM stage-1 bitmap for stride s would live and how selective it would be.
constexpr double kBloom2MaxLoad = 0.4; // entries per bit below which two bits per entry beat one
struct WmPlan {
	uint32_t s = 0, B = 0, fbits = 0;
	bool direct = false, in_smem = true;
	double cost = 1e30;
};

static WmPlan plan_wm_stride(const PatternSet &ps, bool packed, uint32_t s, uint32_t smem_fbits_max, bool allow_global) {
	const uint32_t pd = ps.size();
	const uint32_t Bcap = packed ? 16 : 8;
	const uint32_t B = std::min<uint32_t>(Bcap, ps.m_min - s + 1);
	const uint32_t key_bits = packed ? 2 * B : 8 * B; // bytes path: the block is mixed to 32 bits and always hashed
	const double entries = (double) s * pd;
	WmPlan best;
	auto consider = [&](uint32_t fbits, bool direct, bool in_smem) {
		// a sample is a candidate with probability `rate` and then costs its lane ~60 slots (the rest of the
		// warp waits) plus ~40 per offset it has to probe: about `load` offsets, at most s
		const double load = entries / std::pow(2.0, (double) std::min(fbits, key_bits));
		// a hashed bitmap with room to spare sets two bits per entry (see compile_wm): far fewer false candidates
		const double rate = (!direct && load < kBloom2MaxLoad) ? std::pow(1.0 - std::exp(-2.0 * load), 2.0) : 1.0 - std::exp(-load);
		// per symbol: ~8 lane-instructions per sample (+16 when the bitmap word comes from L2: measured on
		// BASELINE config 4, profiles/r01g_tune.csv)
		const double cost = ((packed ? 8.0 : 10.0) + (in_smem ? 0.0 : 16.0) + rate * 60.0
									+ std::min<double>(s, load) * 40.0) / s;
		if (cost < best.cost - 1e-9) {
			best.s = s;
			best.B = B;
			best.fbits = fbits;
			best.direct = direct;
			best.in_smem = in_smem;
			best.cost = cost;
		}
	};
	if (packed && key_bits <= smem_fbits_max)
		consider(std::max<uint32_t>(key_bits, 5), true, true);
	else if (!packed && key_bits < 13)
		consider(13, false, true);
	else
		consider(std::min<uint32_t>(std::max<uint32_t>(ceil_log2((uint64_t) (entries * 32)), 13), smem_fbits_max), false, true);
	if (allow_global) { // L2-resident bitmap: up to 2^26 bits = 8 MiB
		if (packed && key_bits <= 26 && key_bits > smem_fbits_max)
			consider(key_bits, true, false);
		else if (key_bits > smem_fbits_max)
			consider(std::min<uint32_t>(std::max<uint32_t>(ceil_log2((uint64_t) (entries * 256)), smem_fbits_max + 1), 26),
					false, false);
	}
	return best;
}

static uint32_t smem_fbits_for(uint32_t stage1_bytes) {
	return std::min<uint32_t>(ceil_log2((uint64_t) stage1_bytes * 8 + 1) - 1, 19);
}

static WmPlan plan_wm(const PatternSet &ps, bool packed, uint32_t force_stride, uint32_t smem_fbits_max, bool allow_global) {
	WmPlan plan;
	for (uint32_t s : {16u, 8u, 4u, 2u, 1u}) {
		if (force_stride && force_stride != s)
			continue;
		if (s > ps.m_min)
			continue;
		const WmPlan p = plan_wm_stride(ps, packed, s, smem_fbits_max, allow_global);
		if (p.cost < plan.cost - 1e-9)
			plan = p;
	}
	return plan;
}

static double wm_plan_cost(const PatternSet &ps, bool packed, uint32_t stage1_bytes, bool allow_global) {
	return plan_wm(ps, packed, 0, smem_fbits_for(stage1_bytes), allow_global).cost;
}

static int compile_wm(const PatternSet &ps, const acwm_options &opts, bool packed, const TableProfile &prof, Compiled &c,
		std::string &err) {
	acwm_scan_params &prm = c.prm;
	const uint32_t pd = ps.size();
	const uint32_t smem_fbits_max = smem_fbits_for(prof.stage1_bytes);
	const WmPlan plan = plan_wm(ps, packed, opts.force_stride, smem_fbits_max, !opts.force_smem_tables);
	if (!plan.s) {
		err = "WM: no sampling stride fits (pattern shorter than the forced stride?)";
		return ACWM_ERR_INVALID;
	}
	const uint32_t s = plan.s, B = plan.B, fbits = plan.fbits;
	prm.stride = s;
	prm.depth = B;
	prm.exact_front = 0;
	if (plan.direct) {
		prm.f1_mult = 1;
		prm.f1_sh2 = 0;
	} else {
		prm.f1_mult = kMultF1;
		prm.f1_sh2 = 32 - fbits;
	}
	prm.f1_sh1 = packed ? 32 - 2 * B : 64 - 8 * B; // drop symbols older than the block
	prm.f1_words = (uint32_t) (((uint64_t) 1 << fbits) / 32);
	// Hashed bitmaps (every candidate of a random text is a hash collision there): while the bitmap has room, an entry
	// sets TWO bits of its word -- bit (h >> sh2) & 31 and bit (h >> (sh2 - 5)) & 31 of word h >> (sh2 + 5), h = block *
	// mult: a blocked Bloom filter, one load per sample as before, false candidates ~ (2 load)^2 instead of load
	prm.f1_k = (!plan.direct && (double) s * pd / std::pow(2.0, (double) fbits) < kBloom2MaxLoad && prm.f1_sh2 >= 5) ? 2 : 1;
	std::vector<uint32_t> bm(prm.f1_words, 0);
	for (uint32_t j = 0; j < pd; j++)
		for (uint32_t r = 0; r < s; r++) {
			uint32_t v;
			if (packed)
				v = pack2_tail(ps.pat(j), ps.len[j], B, r);
			else
				v = mix64to32(pack8_tail(ps.pat(j), ps.len[j], B, r));
			const uint32_t h = v * prm.f1_mult, idx = h >> prm.f1_sh2;
			set_bit(bm, idx);
			if (prm.f1_k == 2)
				set_bit(bm, (idx & ~31u) | ((h >> (prm.f1_sh2 - 5)) & 31u));
		}
	c.front_entry_bytes = 4;
	c.front.resize((size_t) prm.f1_words * 4);
	memcpy(c.front.data(), bm.data(), c.front.size());
	// offset masks: a candidate block only has to be probed at the offsets r some pattern holds it at;
	// in shared memory while 2 entries per (pattern, offset) fit 32 KiB, else 8 per pair in L2
	if (s > 1) {
		const uint32_t eb = s > 8 ? 2 : 1;
		const uint32_t rbits_smem_max = ceil_log2(prof.rmask_bytes / eb + 1) - 1; // 32 KiB: 14 / 15 bits
		uint32_t rbits = std::max<uint32_t>(ceil_log2((uint64_t) s * pd * 2), 10);
		prm.r_in_smem = 1;
		if (rbits > rbits_smem_max) {
			// shared memory while at most ~2 (pattern, offset) pairs share an entry, else L2
			if (opts.force_smem_tables || (uint64_t) s * pd <= ((uint64_t) 2 << rbits_smem_max))
				rbits = rbits_smem_max;
			else {
				prm.r_in_smem = 0;
				rbits = std::min<uint32_t>(std::max<uint32_t>(ceil_log2((uint64_t) s * pd * 4), 16), 24);
			}
		}
		if (opts.force_r_bits)
			rbits = std::min<uint32_t>(std::max<uint32_t>(opts.force_r_bits, 10), 16);
		prm.r_mult = kMultR;
		prm.r_sh = 32 - rbits;
		prm.r_entries = 1u << rbits;
		prm.r_entry_bytes = eb;
		c.rmask.assign((size_t) prm.r_entries * prm.r_entry_bytes, 0);
		for (uint32_t j = 0; j < pd; j++)
			for (uint32_t r = 0; r < s; r++) {
				const uint32_t v = packed ? pack2_tail(ps.pat(j), ps.len[j], B, r)
										  : mix64to32(pack8_tail(ps.pat(j), ps.len[j], B, r));
				const uint32_t ri = (uint32_t) (v * prm.r_mult) >> prm.r_sh;
				if (prm.r_entry_bytes == 1)
					c.rmask[ri] |= (uint8_t) (1u << r);
				else
					reinterpret_cast<uint16_t *>(c.rmask.data())[ri] |= (uint16_t) (1u << r);
			}
	}
	build_verify(ps, packed, opts, c, prof);
	c.info.table_in_smem = plan.in_smem ? 1 : 0;
	return ACWM_OK;
}

// ------------------------------------------------------------------ entry point
static uint32_t smem_tables16(const Compiled &c) {
	// tables are rounded up to 16 bytes each in shared memory
	return (c.info.table_in_smem ? (((uint32_t) c.front.size() + 15u) & ~15u) : 0)
			+ (c.prm.r_in_smem ? (((uint32_t) c.rmask.size() + 15u) & ~15u) : 0)
			+ (c.prm.f2_in_smem ? (((uint32_t) c.filter2.size() * 4 + 15u) & ~15u) : 0);
}

// One compilation of the tables under a shared-memory profile (a CTA that owns its SM, or one of two sharing it).
static int compile_with_profile(int algo, const PatternSet &ps, const acwm_options &opts, bool packed, uint32_t budget,
		uint32_t ac_budget, TableProfile prof, Compiled &out, std::string &err) {
	out = Compiled();
	acwm_scan_params &prm = out.prm;
	prm.algo = (uint32_t) algo;
	prm.packed2bit = packed ? 1 : 0;
	prm.alphabet = ps.alphabet;
	prm.m_min = ps.m_min;
	prm.m_max = ps.m_max;
	prof.stage1_bytes = std::min(prof.stage1_bytes, budget / 2); // stage-1 bitmap; offset masks and stage 2 take <= 32 KiB each
	if (algo == ACWM_ALGO_AC) {
		return packed ? compile_ac_packed(ps, opts, ac_budget, prof, out, err) : compile_ac_bytes(ps, opts, ac_budget, out, err);
	}
	return compile_wm(ps, opts, packed, prof, out, err);
}

int compile_tables(int algo, const PatternSet &ps, const acwm_options &opts, Compiled &out, std::string &err) {
	const bool packed = ps.alphabet <= 4 && !opts.force_bytes_path;
	uint32_t budget = opts.smem_table_budget ? opts.smem_table_budget : kDefaultTableBudget;
	const uint32_t hard_cap = kMaxSmem - kSmemReserve - 4 * warp_smem_bytes(2, packed);
	budget = std::min(budget, hard_cap);
	if (algo != ACWM_ALGO_AC && algo != ACWM_ALGO_WM) {
		err = "unknown algorithm id";
		return ACWM_ERR_INVALID;
	}
	if (algo == ACWM_ALGO_AC && ps.m_min != ps.m_max) {
		err = "Aho-Corasick path takes equal-length patterns (preproc_ac has a single m, ac/ac.c:224); "
			  "use ACWM_ALGO_WM for mixed lengths";
		return ACWM_ERR_UNSUPPORTED;
	}
	// AC: leave room for the stage-2 bitmap in case the automaton gets truncated
	const uint32_t ac_budget = budget > 64 * 1024 ? budget - 32 * 1024 : budget / 2;
	int rc = compile_with_profile(algo, ps, opts, packed, budget, ac_budget, kProfileSingle, out, err);
	if (rc != ACWM_OK)
		return rc;
	const bool forced_shape = opts.force_threads || opts.force_stages;
	LaunchShape shape = shape_for_tables(smem_tables16(out), packed, packed && !out.prm.exact_front);
	// Two half-size CTAs per SM (2-bit path): consecutive scans of a stream then share every SM in overlap mode.
	// Taken when the tables fit twice WITHOUT a less selective plan: same stride, block and bitmap sizes as the
	// single-CTA compilation (the offset masks alone may shrink to one entry per (pattern, offset) pair).
	if (packed && opts.force_ctas != 1 && !forced_shape && ps.size() <= 16384) {
		Compiled dual;
		std::string derr;
		// an exact automaton needs no 2-bit copy of the tile: what 12 warps leave of half an SM (minus 1 KiB of per-tile counts)
		const uint32_t ac_dual = std::min(ac_budget, kMaxSmemDual - kSmemReserve - 12 * warp_smem_bytes(1, false) - 1024);
		if (compile_with_profile(algo, ps, opts, packed, std::min(budget, 2 * kProfileDual.stage1_bytes), ac_dual, kProfileDual,
					dual, derr) == ACWM_OK) {
			const acwm_scan_params &a = out.prm, &b = dual.prm;
			const bool same_plan = a.stride == b.stride && a.depth == b.depth && a.exact_front == b.exact_front
					&& a.front_kind == b.front_kind && dual.front.size() == out.front.size()
					&& dual.info.table_in_smem == out.info.table_in_smem && a.f2_words == b.f2_words
					&& a.r_in_smem == b.r_in_smem && a.f2_in_smem == b.f2_in_smem
					&& (b.r_entries == a.r_entries || (uint64_t) b.stride * ps.size() <= b.r_entries);
			const LaunchShape ds = shape_for_tables(smem_tables16(dual), true, !b.exact_front, true);
			if (ds.warps && (same_plan || opts.force_ctas == 2)) {
				out = std::move(dual);
				shape = ds;
			}
		}
		if (opts.force_ctas == 2 && shape.ctas != 2) {
			err = "force_ctas = 2: the scan tables do not fit shared memory twice";
			return ACWM_ERR_INVALID;
		}
	}
	const bool pk_copy = packed && !out.prm.exact_front;
	if (forced_shape) { // tuning / tests
		LaunchShape want{opts.force_threads ? opts.force_threads / 32 : shape.warps,
				opts.force_stages ? opts.force_stages : shape.stages, opts.force_ctas == 2 ? 2u : 1u};
		// the kernels that exist (launch_front, scan_kernel.cuh): 2-bit path 1 slot per warp; bytes path 2 slots with
		// 4..16 warps or 1 slot with 16..24 warps
		const uint32_t w = want.warps;
		const bool kernel_exists = packed ? (want.stages == 1 && (w == 32 || w == 24 || w == 16 || w == 12 || w == 8 || w == 4))
				: want.stages == 2 ? (w == 16 || w == 12 || w == 8 || w == 4)
								   : (want.stages == 1 && (w == 24 || w == 20 || w == 16));
		const bool ok = kernel_exists && want.warps * 32 == (opts.force_threads ? opts.force_threads : want.warps * 32)
				&& (want.ctas == 1 || (packed && (want.warps == 16 || want.warps == 12)))
				&& shape_fits(smem_tables16(out), want, pk_copy);
		if (!ok) {
			err = "forced launch shape (threads / stages / CTAs per SM) not available for this table size";
			return ACWM_ERR_INVALID;
		}
		shape = want;
	}
	if (!shape.warps) {
		err = "scan tables exceed shared memory";
		return ACWM_ERR_UNSUPPORTED;
	}
	const acwm_scan_params &prm = out.prm;
	acwm_info &inf = out.info;
	inf.algo = (uint32_t) algo;
	inf.alphabet = ps.alphabet;
	inf.n_patterns = ps.p_in;
	inf.n_distinct = ps.size();
	inf.m_min = ps.m_min;
	inf.m_max = ps.m_max;
	inf.packed2bit = prm.packed2bit;
	inf.stride = prm.stride;
	inf.depth = prm.depth;
	inf.exact_front = prm.exact_front;
	inf.n_rows = prm.front_kind ? prm.v_rows : prm.n_rows;
	inf.n_states = (algo == ACWM_ALGO_AC) ? full_trie_states(ps, ps.alphabet) : 0;
	inf.threads = shape.warps * 32;
	inf.stages = shape.stages;
	inf.ctas_per_sm = shape.ctas;
	inf.front_kind = prm.front_kind;
	inf.smem_bytes = shape_smem(smem_tables16(out), shape, pk_copy);
	if (shape.ctas == 2) { // never three CTAs on an SM: the bound on scans in flight (Work, scan_common.cuh) rests on it
		inf.smem_bytes = std::max(inf.smem_bytes, kMinSmemDual);
		// the same tables under one full-size CTA per SM: the shape of long scans with a verification stage (api.cu)
		const LaunchShape single = shape_for_tables(smem_tables16(out), packed, pk_copy, false);
		if (single.warps > shape.warps && !forced_shape && opts.force_ctas != 2) {
			out.alt_threads = single.warps * 32;
			out.alt_smem_bytes = shape_smem(smem_tables16(out), single, pk_copy);
		}
	}
	inf.table_bytes = out.front.size() + out.rmask.size() + out.filter2.size() * 4 + out.bucket_start.size() * 4
			+ out.entries.size() * sizeof(acwm_ventry) + ps.bytes.size() + out.vdfa.size() * 4;
	return ACWM_OK;
}

// ------------------------------------------------------------------ reference-layout tables (shims)
// Set-Horspool (sh/sh.c:78-149): the goto function of the trie of the reversed patterns, ids in creation order,
// root row zeroed first (sh_init, :58-59), terminal flag where a pattern ends (:131).  The trie is kept privately
// (flat child array), so what the caller's arrays held before does not matter.
void fill_reference_sh_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_final, unsigned *n_states, unsigned *n_distinct) {
	const size_t A = (size_t) alphabet;
	for (int c = 0; c < alphabet; c++)
		state_transition[c] = 0;
	std::vector<int> child(A, -1); // child[state * A + sym]
	std::vector<uint8_t> is_final(1, 0);
	unsigned distinct = 0;
	for (int i = 0; i < p; i++) {
		int state = 0;
		for (int j = m - 1; j >= 0; j--) {
			const unsigned c = rows[i][j];
			int nx = child[(size_t) state * A + c];
			if (nx < 0) {
				nx = (int) is_final.size();
				is_final.push_back(0);
				child.resize(child.size() + A, -1);
				child[(size_t) state * A + c] = nx;
				state_transition[(size_t) state * A + c] = nx;
			}
			state = nx;
		}
		if (!is_final[(size_t) state]) {
			is_final[(size_t) state] = 1;
			state_final[state] = 1;
			distinct++;
		}
	}
	if (n_states)
		*n_states = (unsigned) is_final.size();
	if (n_distinct)
		*n_distinct = distinct;
}

// Set Backward Oracle Matching (sbom/sbom.c:51-150): the factor oracle of the reversed patterns.  A pattern first
// follows whatever transitions exist -- trie edges AND the extra oracle edges earlier patterns left (:62-70) -- and
// creates states for the rest; each new state hangs extra edges on the supply chain of its parent (:104-113) and takes
// its own supply link from where that walk stops (:115-118).  F(q): cell 0 = number of patterns ending in q, cells
// 1.. = their row numbers (:145-146), 200 cells per state as the reference lays them out.
void fill_reference_sbom_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_final_multi, unsigned *n_states, unsigned *n_patterns) {
	const size_t A = (size_t) alphabet;
	for (int c = 0; c < alphabet; c++)
		state_transition[c] = 0;
	std::vector<int> next(A, -1), fail(1, -1);
	std::vector<unsigned> num(1, 0);
	for (int i = 0; i < p; i++) {
		int state = 0, j = m - 1;
		while (j >= 0) {
			const int nx = next[(size_t) state * A + rows[i][j]];
			if (nx < 0)
				break;
			state = nx;
			j--;
		}
		for (; j >= 0; j--) {
			const unsigned c = rows[i][j];
			const int nx = (int) fail.size();
			fail.push_back(0);
			num.push_back(0);
			next.resize(next.size() + A, -1);
			state_transition[(size_t) state * A + c] = nx;
			next[(size_t) state * A + c] = nx;
			int k = fail[(size_t) state];
			while (k >= 0 && next[(size_t) k * A + c] < 0) {
				next[(size_t) k * A + c] = nx;
				state_transition[(size_t) k * A + c] = nx;
				k = fail[(size_t) k];
			}
			fail[(size_t) nx] = k >= 0 ? next[(size_t) k * A + c] : 0;
			state = nx;
		}
		if (num[(size_t) state] < 199) { // the reference's F(q) holds 200 cells
			state_final_multi[(size_t) state * 200] = num[(size_t) state] + 1;
			state_final_multi[(size_t) state * 200 + num[(size_t) state] + 1] = (unsigned) i;
			num[(size_t) state]++;
		}
	}
	if (n_states)
		*n_states = (unsigned) fail.size();
	if (n_patterns)
		*n_patterns = (unsigned) p;
}

// Shift-Or with q-grams, 8-symbol patterns (sog/sog8.c:113-170): T8[3-gram] has bit i CLEAR when some pattern holds the
// 3-gram at offset i (0..5); scanner_hs = the 32-bit hash of each pattern (bytes 0..3 xor bytes 4..7, big endian),
// sorted, with scanner_index carrying the pattern rows; scanner_hs2 = the bitmap of the 16-bit fold of the hashes.
// Two places where this cannot be byte-identical: the reference folds an UNINITIALISED variable into the two-level
// hash (sog8.c:128) -- the fold of the pattern's own hash, which its search computes (:50-51), is stored here -- and
// its quicksort (:28-47) leaves equal hashes in an order of its own (here: by pattern row).
void fill_reference_sog8_tables(const uint8_t *const *rows, int p, uint8_t *T8, uint32_t *scanner_hs, int *scanner_index,
		uint8_t *scanner_hs2) {
	memset(T8, 0xff, (size_t) 1 << 24);
	memset(scanner_hs2, 0, 32 * 256);
	std::vector<std::pair<uint32_t, int>> hs((size_t) p);
	auto get32 = [](const uint8_t *a) { return ((uint32_t) a[0] << 24) + ((uint32_t) a[1] << 16) + ((uint32_t) a[2] << 8) + a[3]; };
	for (int i = 0; i < p; i++) {
		const uint8_t *q = rows[i];
		const uint32_t h = get32(q) ^ get32(q + 4);
		hs[(size_t) i] = {h, i};
		const uint16_t h2 = (uint16_t) ((h >> 16) ^ h);
		scanner_hs2[h2 >> 3] |= (uint8_t) (1u << (h2 & 7));
		for (int k = 0; k < 6; k++) {
			const uint32_t g = (uint32_t) q[k] + ((uint32_t) q[k + 1] << 8) + ((uint32_t) q[k + 2] << 16);
			T8[g] &= (uint8_t) (0xff - (1u << k));
		}
	}
	std::stable_sort(hs.begin(), hs.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
	for (int i = 0; i < p; i++) {
		scanner_hs[i] = hs[(size_t) i].first;
		scanner_index[i] = hs[(size_t) i].second;
	}
}

void fill_reference_ac_tables(const uint8_t *const *rows, int m, int p, int alphabet, int *state_transition,
		unsigned *state_supply, unsigned *state_final, unsigned *n_states, unsigned *n_distinct) {
	// Same observable content as preproc_ac (ac/ac.c:224-245): root row zeroed first
	// (:61-62), goto edges recorded as they are created with ids in creation order
	// (:159-162), terminal flags (:186), failure ids of depth >= 2 states (:114).
	// Cells the reference never writes are left as the caller initialised them; the
	// trie itself is kept privately so the caller's initial contents do not matter.
	const size_t A = (size_t) alphabet;
	for (int c = 0; c < alphabet; c++)
		state_transition[c] = 0;
	std::unordered_map<uint64_t, uint32_t> edge; // (state << 8 | sym) -> child
	edge.reserve((size_t) m * (size_t) p * 2);
	struct Edge {
		uint32_t from, sym, to;
	};
	std::vector<Edge> edges;
	std::vector<uint8_t> fin(1, 0);
	unsigned ns = 1, nd = 0;
	for (int j = 0; j < p; j++) {
		uint32_t st = 0;
		for (int i = 0; i < m; i++) {
			const uint32_t sym = rows[j][i];
			auto it = edge.find(((uint64_t) st << 8) | sym);
			if (it == edge.end()) {
				edge.emplace(((uint64_t) st << 8) | sym, ns);
				edges.push_back(Edge{st, sym, ns});
				state_transition[st * A + sym] = (int) ns;
				fin.push_back(0);
				st = ns++;
			} else
				st = it->second;
		}
		if (!fin[st]) {
			fin[st] = 1;
			state_final[st] = 1;
			nd++;
		}
	}
	// children of each state, ascending symbol (the reference's BFS visits i = 0..alphabet-1)
	std::sort(edges.begin(), edges.end(), [](const Edge &a, const Edge &b) {
		return a.from != b.from ? a.from < b.from : a.sym < b.sym;
	});
	std::vector<uint32_t> first(ns + 1, 0);
	for (const Edge &e : edges)
		first[e.from + 1]++;
	for (unsigned s = 0; s < ns; s++)
		first[s + 1] += first[s];
	auto go = [&](uint32_t st, uint32_t sym) -> int64_t {
		auto it = edge.find(((uint64_t) st << 8) | sym);
		return it == edge.end() ? -1 : (int64_t) it->second;
	};
	std::vector<uint32_t> queue;
	queue.reserve(ns);
	std::vector<uint32_t> fail(ns, 0);
	for (uint32_t k = first[0]; k < first[1]; k++)
		queue.push_back(edges[k].to);
	for (size_t qi = 0; qi < queue.size(); qi++) {
		const uint32_t cur = queue[qi];
		for (uint32_t k = first[cur]; k < first[cur + 1]; k++) {
			const uint32_t c = edges[k].sym, s = edges[k].to;
			queue.push_back(s);
			uint32_t st = fail[cur];
			int64_t nx;
			while ((nx = go(st, c)) < 0 && st != 0)
				st = fail[st];
			const uint32_t f = nx < 0 ? 0u : (uint32_t) nx; // root self-loop on an absent symbol
			fail[s] = f;
			state_supply[s] = f;
		}
	}
	if (n_states)
		*n_states = ns;
	if (n_distinct)
		*n_distinct = nd;
}

unsigned reference_wu_shiftsize(int alphabet) {
	switch (alphabet) { // wu/wu.c:18-47
	case 2: return 22;
	case 4: return 64;
	case 8: return 148;
	case 20: return 400;
	case 128: return 2668;
	case 256: return 5356;
	case 512: return 10732;
	case 1024: return 21484;
	default: return 0;
	}
}

void fill_reference_wu_tables(const uint8_t *const *rows, const uint8_t *flat, int m, int p, int B, int nbits,
		int *SHIFT, int *PREFIX_value, int *PREFIX_index, int *PREFIX_size) {
	// wu/wu.c:109-149 / 211-251; the caller pre-initialised SHIFT and PREFIX_size (main.c:444-449)
	for (int j = 0; j < p; j++) {
		const uint8_t *s = rows ? rows[j] : flat + (size_t) j * m;
		for (int q = m; q >= B; --q) {
			unsigned h = s[q - 3];
			h <<= nbits;
			h += s[q - 2];
			h <<= nbits;
			h += s[q - 1];
			const int shiftlen = m - q;
			if (shiftlen < SHIFT[h])
				SHIFT[h] = shiftlen;
			if (shiftlen == 0) {
				unsigned ph = s[0];
				ph <<= nbits;
				ph += s[1];
				PREFIX_value[(size_t) h * p + PREFIX_size[h]] = (int) ph;
				PREFIX_index[(size_t) h * p + PREFIX_size[h]] = j;
				PREFIX_size[h]++;
			}
		}
	}
}

} // namespace acwm
