// Reference-shaped entry points (same names and parameter lists as the reference's
// smatcher.h:89-106 and cuda/cuda_ac.cu / cuda/cuda_wm.cu wrappers), implemented on
// top of the native acwm_* API.  The reference's main.c links against these unchanged.
//
// Error convention: these signatures have no error channel; like the reference's
// checkCudaErrors / fail (cuda/cuda.h:26-47) they print to stderr and exit(1).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "matcher.hpp"

using namespace acwm;

extern "C" {
unsigned short m_nBitsInShift = 2; // smatcher.h:71 (main.c:431 sets 2)
unsigned int shiftsize = 0;        // smatcher.h:73
}

namespace {

[[noreturn]] void die(const char *where) {
	fprintf(stderr, "acwm %s: %s\n", where, acwm_last_error());
	exit(1);
}

std::mutex g_mu;
struct WmEntry {
	acwm_matcher *mt;
	uint64_t sig;
};
std::unordered_map<const void *, WmEntry> g_by_table; // SHIFT pointer / state_transition pointer -> matcher
unsigned long long g_last_count = 0;

uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
	const uint8_t *b = (const uint8_t *) p;
	for (size_t i = 0; i < n; i++) {
		h ^= b[i];
		h *= 1099511628211ull;
	}
	return h;
}

acwm_matcher *build_or_die(int algo, const uint8_t *flat, int m, int p, int alphabet, const char *where) {
	acwm_matcher *mt = nullptr;
	if (acwm_build(algo, flat, nullptr, (uint32_t) m, (uint32_t) p, (uint32_t) alphabet, nullptr, &mt) != ACWM_OK)
		die(where);
	return mt;
}

unsigned long long search_or_die(acwm_matcher *mt, const unsigned char *text, int n, const char *where) {
	uint64_t count = 0;
	const int rc = acwm_search_host(mt, text, n > 0 ? (uint64_t) n : 0, &count, nullptr, 0, nullptr);
	if (rc != ACWM_OK)
		die(where);
	return count;
}

std::vector<uint8_t> flatten_rows(unsigned char **pattern, int m, int p) {
	std::vector<uint8_t> flat((size_t) m * p);
	for (int j = 0; j < p; j++)
		memcpy(flat.data() + (size_t) j * m, pattern[j], (size_t) m);
	return flat;
}

void remember(const void *key, acwm_matcher *mt, uint64_t sig) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_by_table.find(key);
	if (it != g_by_table.end()) {
		acwm_free(it->second.mt);
		g_by_table.erase(it);
	}
	g_by_table[key] = WmEntry{mt, sig};
}

void forget(const void *key) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_by_table.find(key);
	if (it != g_by_table.end()) {
		acwm_free(it->second.mt);
		g_by_table.erase(it);
	}
}

acwm_matcher *lookup(const void *key, uint64_t sig) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_by_table.find(key);
	return (it != g_by_table.end() && it->second.sig == sig) ? it->second.mt : nullptr;
}

// Every terminal state of the flat goto table spells one pattern (ac/ac.c:162,186).
std::vector<uint8_t> patterns_from_goto(int m, int p_size, int alphabet, const int *tr, const unsigned *fin, int *p_out) {
	std::vector<uint8_t> flat;
	std::vector<uint8_t> path((size_t) m);
	struct Frame {
		int state, depth, next_sym;
	};
	std::vector<Frame> stack;
	stack.push_back(Frame{0, 0, 0});
	const long long max_state = (long long) m * p_size + 1;
	int found = 0;
	while (!stack.empty()) {
		Frame &f = stack.back();
		if (f.depth == m) {
			if (fin[f.state]) {
				flat.insert(flat.end(), path.begin(), path.end());
				found++;
			}
			stack.pop_back();
			continue;
		}
		if (f.next_sym >= alphabet) {
			stack.pop_back();
			continue;
		}
		const int c = f.next_sym++;
		const int t = tr[(size_t) f.state * alphabet + c];
		if (t > 0 && t < max_state) { // 0 = root self-loop, -1 = no edge
			path[(size_t) f.depth] = (uint8_t) c;
			const int d = f.depth + 1;
			stack.push_back(Frame{t, d, 0});
		}
	}
	*p_out = found;
	return flat;
}

void cuda_ac_common(int variant, int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final) {
	(void) state_supply;
	// the goto table IS the pattern set (two sets can share their terminal-state layout, e.g. any two sets of one
	// pattern): it is part of the signature, so refilled or recycled arrays never reach a stale matcher
	const size_t n_states = (size_t) m * p_size + 1;
	const uint64_t sig = fnv(state_transition, n_states * (size_t) alphabet * sizeof(int),
			fnv(state_final, n_states * sizeof(unsigned), fnv(&m, sizeof(m), fnv(&alphabet, sizeof(alphabet)))));
	acwm_matcher *mt = lookup(state_transition, sig);
	if (!mt) {
		int p = 0;
		std::vector<uint8_t> flat = patterns_from_goto(m, p_size, alphabet, state_transition, state_final, &p);
		if (p == 0) {
			fprintf(stderr, "acwm cuda_ac%d: the goto table holds no terminal state\n", variant);
			exit(1);
		}
		mt = build_or_die(ACWM_ALGO_AC, flat.data(), m, p, alphabet, "cuda_ac");
		remember(state_transition, mt, sig);
	}
	const unsigned long long matches = search_or_die(mt, text, n, "cuda_ac");
	g_last_count = matches;
	// cuda/cuda_ac.cu:675
	printf("Kernel %d matches \t%llu\t time \t%f\n", variant, matches, acwm_last_kernel_seconds(mt));
}

int cuda_wm_common(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int *SHIFT,
		double *gpuTime) {
	const uint64_t sig = fnv(pattern, (size_t) m * p_size, fnv(&m, sizeof(m)));
	acwm_matcher *mt = lookup(SHIFT, sig);
	if (!mt) {
		mt = build_or_die(ACWM_ALGO_WM, pattern, m, p_size, alphabet, "cuda_wm");
		remember(SHIFT, mt, sig);
	}
	const unsigned long long matches = search_or_die(mt, text, n, "cuda_wm");
	if (gpuTime)
		*gpuTime = acwm_last_kernel_seconds(mt); // cuda/cuda_wm.cu:300
	return (int) matches;
}

} // namespace

extern "C" {

unsigned long long acwm_shim_last_count(void) { return g_last_count; }

// ------------------------------------------------------------------ Aho-Corasick
struct ac_table *preproc_ac(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final) {
	struct ac_table *table = (struct ac_table *) malloc(sizeof(struct ac_table));
	if (!table) {
		fprintf(stderr, "Could not initialize table\n"); // ac/ac.c:235
		exit(1);
	}
	forget(state_transition); // a cuda_acN matcher compiled from the previous content of these arrays
	unsigned ns = 0, nd = 0;
	fill_reference_ac_tables((const uint8_t *const *) pattern, m, p_size, alphabet, state_transition, state_supply,
			state_final, &ns, &nd);
	std::vector<uint8_t> flat = flatten_rows(pattern, m, p_size);
	acwm_matcher *mt = build_or_die(ACWM_ALGO_AC, flat.data(), m, p_size, alphabet, "preproc_ac");
	table->idcounter = ns;
	table->patterncounter = nd;
	table->zerostate = (struct ac_state *) mt;
	return table;
}

unsigned search_ac(unsigned char *text, int n, struct ac_table *table) {
	return (unsigned) search_or_die((acwm_matcher *) table->zerostate, text, n, "search_ac");
}

void free_ac(struct ac_table *table, int alphabet) {
	(void) alphabet;
	if (!table)
		return;
	acwm_free((acwm_matcher *) table->zerostate);
	free(table);
}

// ------------------------------------------------------------------ Wu-Manber
void wu_determine_shiftsize(int alphabet) {
	const unsigned s = reference_wu_shiftsize(alphabet);
	if (!s) {
		fprintf(stderr, "The alphabet size is not supported by wu-manber\n"); // wu/wu.c:46
		exit(1);
	}
	shiftsize = s;
}

static void preproc_wu_common(unsigned char **rows, unsigned char *flat_in, int m, int p_size, int alphabet, int B,
		int *SHIFT, int *PREFIX_value, int *PREFIX_index, int *PREFIX_size) {
	fill_reference_wu_tables((const uint8_t *const *) rows, flat_in, m, p_size, B, (int) m_nBitsInShift, SHIFT,
			PREFIX_value, PREFIX_index, PREFIX_size);
	std::vector<uint8_t> flat = rows ? flatten_rows(rows, m, p_size)
									 : std::vector<uint8_t>(flat_in, flat_in + (size_t) m * p_size);
	acwm_matcher *mt = build_or_die(ACWM_ALGO_WM, flat.data(), m, p_size, alphabet, "preproc_wu");
	remember(SHIFT, mt, fnv(flat.data(), flat.size(), fnv(&m, sizeof(m))));
}

void preproc_wu(unsigned char **pattern, int m, int p_size, int alphabet, int B, int *SHIFT, int *PREFIX_value,
		int *PREFIX_index, int *PREFIX_size) {
	preproc_wu_common(pattern, nullptr, m, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size);
}

void preproc_wu2(unsigned char *pattern, int m, int p_size, int alphabet, int B, int *SHIFT, int *PREFIX_value,
		int *PREFIX_index, int *PREFIX_size) {
	preproc_wu_common(nullptr, pattern, m, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size);
}

static unsigned search_wu_common(const std::vector<uint8_t> &flat, int m, int p_size, unsigned char *text, int n,
		int *SHIFT) {
	const uint64_t sig = fnv(flat.data(), flat.size(), fnv(&m, sizeof(m)));
	acwm_matcher *mt = lookup(SHIFT, sig);
	if (!mt) {
		// preproc_wu was not called with these tables: compile from the patterns; the
		// alphabet is not a parameter of search_wu, the smallest one that holds them is used
		unsigned mx = 0;
		for (uint8_t b : flat)
			mx = b > mx ? b : mx;
		const int alphabet = mx < 4 ? 4 : 256;
		mt = build_or_die(ACWM_ALGO_WM, flat.data(), m, p_size, alphabet, "search_wu");
		remember(SHIFT, mt, sig);
	}
	return (unsigned) search_or_die(mt, text, n, "search_wu");
}

unsigned int search_wu(unsigned char **pattern, int m, int p_size, unsigned char *text, int n, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size) {
	(void) PREFIX_value, (void) PREFIX_index, (void) PREFIX_size;
	return search_wu_common(flatten_rows(pattern, m, p_size), m, p_size, text, n, SHIFT);
}

unsigned int search_wu2(unsigned char *pattern, int m, int p_size, unsigned char *text, int n, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size) {
	(void) PREFIX_value, (void) PREFIX_index, (void) PREFIX_size;
	return search_wu_common(std::vector<uint8_t>(pattern, pattern + (size_t) m * p_size), m, p_size, text, n, SHIFT);
}

// ------------------------------------------------------------------ sibling algorithms (sh/sh.c, sbom/sbom.c, sog/sog8.c)
// Set-Horspool, Set Backward Oracle Matching and Shift-Or with q-grams count the same thing as Aho-Corasick and
// Wu-Manber: the end positions whose window of m symbols is one of the (distinct) patterns.  Behind their entry points
// sits the same matcher; the caller's flat tables are filled as the reference fills them (tables.cpp).
struct ac_table *preproc_sh(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_final) {
	struct ac_table *table = (struct ac_table *) malloc(sizeof(struct ac_table));
	if (!table) {
		fprintf(stderr, "Could not initialize table\n"); // sh/sh.c:172
		exit(1);
	}
	unsigned ns = 0, nd = 0;
	fill_reference_sh_tables((const uint8_t *const *) pattern, m, p_size, alphabet, state_transition, state_final, &ns, &nd);
	std::vector<uint8_t> flat = flatten_rows(pattern, m, p_size);
	table->idcounter = ns;
	table->patterncounter = nd;
	table->zerostate = (struct ac_state *) build_or_die(ACWM_ALGO_AC, flat.data(), m, p_size, alphabet, "preproc_sh");
	return table;
}

unsigned search_sh(int m, unsigned char *text, int n, struct ac_table *table, int *bmBc) {
	(void) m, (void) bmBc; // the bad-character shifts steer the reference's skip loop (sh/sh.c:175), not the result
	return (unsigned) search_or_die((acwm_matcher *) table->zerostate, text, n, "search_sh");
}

void free_sh(struct ac_table *table, int alphabet) { free_ac(table, alphabet); }

struct sbom_table *preproc_sbom(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_final_multi) {
	struct sbom_table *table = (struct sbom_table *) malloc(sizeof(struct sbom_table));
	if (!table) {
		fprintf(stderr, "Could not initialize table\n"); // sbom/sbom.c:207
		exit(1);
	}
	unsigned ns = 0, np = 0;
	fill_reference_sbom_tables((const uint8_t *const *) pattern, m, p_size, alphabet, state_transition, state_final_multi, &ns,
			&np);
	std::vector<uint8_t> flat = flatten_rows(pattern, m, p_size);
	table->idcounter = ns;
	table->patterncounter = np;
	table->zerostate = (struct sbom_state *) build_or_die(ACWM_ALGO_AC, flat.data(), m, p_size, alphabet, "preproc_sbom");
	return table;
}

unsigned search_sbom(unsigned char **pattern, int m, unsigned char *text, int n, struct sbom_table *table) {
	(void) pattern, (void) m; // the reference verifies oracle hits against the rows (sbom/sbom.c:178); the matcher owns its copy
	return (unsigned) search_or_die((acwm_matcher *) table->zerostate, text, n, "search_sbom");
}

void free_sbom(struct sbom_table *table, int m) {
	(void) m;
	if (!table)
		return;
	acwm_free((acwm_matcher *) table->zerostate);
	free(table);
}

// sog8 has no handle: like the Wu-Manber entry points, the matcher is remembered by the caller's table (T8) and
// recognised by the patterns.  Patterns are 8 symbols (the reference compares 8 bytes whatever m says, sog8.c:79).
static uint64_t sog_sig(unsigned char **pattern, int p_size) {
	uint64_t h = fnv(&p_size, sizeof(p_size));
	for (int j = 0; j < p_size; j++)
		h = fnv(pattern[j], 8, h);
	return h;
}

void preproc_sog8(uint8_t *T8, uint32_t *scanner_hs, int *scanner_index, uint8_t *scanner_hs2, unsigned char **pattern, int m,
		unsigned char *text, int n, int p_size, int B) {
	(void) text, (void) n, (void) B;
	if (m != 8) {
		fprintf(stderr, "acwm preproc_sog8: patterns of 8 symbols only (sog/sog8.c:79)\n");
		exit(1);
	}
	fill_reference_sog8_tables((const uint8_t *const *) pattern, p_size, T8, scanner_hs, scanner_index, scanner_hs2);
	std::vector<uint8_t> flat = flatten_rows(pattern, 8, p_size);
	remember(T8, build_or_die(ACWM_ALGO_AC, flat.data(), 8, p_size, 256, "preproc_sog8"), sog_sig(pattern, p_size));
}

unsigned int search_sog8(uint8_t *T8, uint32_t *scanner_hs, int *scanner_index, uint8_t *scanner_hs2, unsigned char **pattern,
		int m, unsigned char *text, int n, int p_size, int B) {
	(void) scanner_hs, (void) scanner_index, (void) scanner_hs2, (void) B;
	if (m != 8) {
		fprintf(stderr, "acwm search_sog8: patterns of 8 symbols only (sog/sog8.c:79)\n");
		exit(1);
	}
	acwm_matcher *mt = lookup(T8, sog_sig(pattern, p_size));
	if (!mt) { // preproc_sog8 was not called with this table
		std::vector<uint8_t> flat = flatten_rows(pattern, 8, p_size);
		mt = build_or_die(ACWM_ALGO_AC, flat.data(), 8, p_size, 256, "search_sog8");
		remember(T8, mt, sog_sig(pattern, p_size));
	}
	return (unsigned) search_or_die(mt, text, n, "search_sog8");
}

void acwm_shim_forget(const void *table) { forget(table); }

// ------------------------------------------------------------------ GPU wrappers
#define ACWM_CUDA_AC(N)                                                                                         \
	void cuda_ac##N(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,         \
			unsigned int *state_supply, unsigned int *state_final) {                                            \
		cuda_ac_common(N, m, text, n, p_size, alphabet, state_transition, state_supply, state_final);           \
	}
ACWM_CUDA_AC(1)
ACWM_CUDA_AC(2)
ACWM_CUDA_AC(3)
ACWM_CUDA_AC(4)
ACWM_CUDA_AC(5)
#undef ACWM_CUDA_AC

#define ACWM_CUDA_WM(N)                                                                                         \
	int cuda_wm##N(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B,  \
			int *SHIFT, int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime) {              \
		(void) B, (void) PREFIX_value, (void) PREFIX_index, (void) PREFIX_size;                                 \
		return cuda_wm_common(pattern, m, text, n, p_size, alphabet, SHIFT, gpuTime);                           \
	}
ACWM_CUDA_WM(1)
ACWM_CUDA_WM(2)
ACWM_CUDA_WM(3)
ACWM_CUDA_WM(4)
ACWM_CUDA_WM(5)
#undef ACWM_CUDA_WM

} // extern "C"
