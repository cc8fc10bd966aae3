/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * CPU restatement ("port") of the reference's Aho-Corasick and Wu-Manber search
 * path, written from the behaviour of /root/reference/ac/ac.c and wu/wu.c, plus two
 * independent window-membership scanners.  Unlike the reference functions (which
 * return only a count) everything here also EMITS MATCH POSITIONS, using the
 * reference's own convention: the position of a match is `column`, the 0-based
 * index of the LAST byte of the occurrence (commented printfs at ac/ac.c:217,
 * wu/wu.c:93,195).
 *
 * Parity pinning: tests/test_oracle.py checks, on every fixture and on seeded
 * random inputs, that  port count == reference count (oracle/_ref, the unmodified
 * reference compiled from /root/reference) and that the reference-layout tables
 * exported here are byte-identical to the ones the reference fills.  The golden
 * vectors under tests/golden/ were generated with the reference present
 * (tests/golden/make_golden.py) and carry the reference's counts.
 *
 * All sizes are 64-bit here (the reference uses `int n`, smatcher.h:90,105).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EMIT(pos_buf, cap, count, value)                    \
	do {                                                    \
		if ((pos_buf) && (count) < (cap))                   \
			(pos_buf)[(count)] = (uint64_t) (value);        \
		(count)++;                                          \
	} while (0)

/* ============================ Aho-Corasick ============================ */

typedef struct {
	int alphabet, m;
	uint32_t n_states;   /* idcounter  (smatcher.h:50) */
	uint32_t n_distinct; /* patterncounter (smatcher.h:51) */
	int32_t *go;         /* [n_states*alphabet], -1 = no goto edge; row 0 absent edges = 0 */
	uint32_t *fail;      /* state_supply */
	uint8_t *fin;        /* state_final  */
} oracle_ac;

/* Trie insertion in pattern order, ids in creation order (ac/ac.c:127-196: new
 * state id = idcounter++ at :159, goto edge recorded at :162, final flag at :186,
 * identical patterns collapse onto one terminal state at :183), then failure
 * links breadth-first (ac/ac.c:79-124: depth-1 states fail to the root :91,
 * deeper states follow the parent's failure chain until a goto edge exists
 * :103-112; root self-loops on absent symbols :86-88).  The BFS queue is an
 * array, so this is linear where the reference's list_append (ac/list.h:57-74)
 * is quadratic -- same links, same ids. */
oracle_ac *oracle_ac_build(const uint8_t *patterns_flat, int m, int p, int alphabet) {
	oracle_ac *a = (oracle_ac *) calloc(1, sizeof(*a));
	size_t max_states = (size_t) m * (size_t) p + 1;
	a->alphabet = alphabet;
	a->m = m;
	a->go = (int32_t *) malloc(max_states * (size_t) alphabet * sizeof(int32_t));
	memset(a->go, -1, max_states * (size_t) alphabet * sizeof(int32_t));
	a->fail = (uint32_t *) calloc(max_states, sizeof(uint32_t));
	a->fin = (uint8_t *) calloc(max_states, 1);
	a->n_states = 1;
	for (int j = 0; j < p; j++) {
		const uint8_t *s = patterns_flat + (size_t) j * m;
		uint32_t st = 0;
		for (int i = 0; i < m; i++) {
			int32_t nx = a->go[(size_t) st * alphabet + s[i]];
			if (nx < 0) {
				nx = (int32_t) a->n_states++;
				a->go[(size_t) st * alphabet + s[i]] = nx;
			}
			st = (uint32_t) nx;
		}
		if (!a->fin[st]) {
			a->fin[st] = 1;
			a->n_distinct++;
		}
	}
	uint32_t *queue = (uint32_t *) malloc((size_t) a->n_states * sizeof(uint32_t));
	size_t qh = 0, qt = 0;
	for (int c = 0; c < alphabet; c++) {
		int32_t nx = a->go[c];
		if (nx < 0)
			a->go[c] = 0; /* root self-loop; also what ac_init leaves in state_transition row 0 (ac/ac.c:61-62) */
		else {
			a->fail[nx] = 0;
			queue[qt++] = (uint32_t) nx;
		}
	}
	while (qh < qt) {
		uint32_t cur = queue[qh++];
		for (int c = 0; c < alphabet; c++) {
			int32_t s = a->go[(size_t) cur * alphabet + c];
			if (s < 0)
				continue;
			queue[qt++] = (uint32_t) s;
			uint32_t st = a->fail[cur];
			while (a->go[(size_t) st * alphabet + c] < 0)
				st = a->fail[st];
			a->fail[s] = (uint32_t) a->go[(size_t) st * alphabet + c];
		}
	}
	free(queue);
	return a;
}

void oracle_ac_free(oracle_ac *a) {
	if (!a)
		return;
	free(a->go);
	free(a->fail);
	free(a->fin);
	free(a);
}

uint32_t oracle_ac_states(const oracle_ac *a) { return a->n_states; }
uint32_t oracle_ac_distinct(const oracle_ac *a) { return a->n_distinct; }

/* Writes the three flat arrays in the exact layout/content the reference's
 * preproc_ac leaves in its caller's buffers (sized (m*p+1)[*alphabet], caller
 * pre-fill -1 / 0 / 0 as main.c:410-420). */
void oracle_ac_export_ref_tables(const oracle_ac *a, int p, int32_t *state_transition, uint32_t *state_supply,
		uint32_t *state_final) {
	size_t max_states = (size_t) a->m * (size_t) p + 1;
	memset(state_transition, -1, max_states * (size_t) a->alphabet * sizeof(int32_t));
	memset(state_supply, 0, max_states * sizeof(uint32_t));
	memset(state_final, 0, max_states * sizeof(uint32_t));
	memcpy(state_transition, a->go, (size_t) a->n_states * (size_t) a->alphabet * sizeof(int32_t));
	for (uint32_t s = 0; s < a->n_states; s++) {
		state_supply[s] = a->fail[s];
		state_final[s] = a->fin[s];
	}
}

/* The scan loop of search_ac (ac/ac.c:198-222): follow failure links until a goto
 * edge exists, take it, count the column if the state is terminal. */
uint64_t oracle_ac_scan(const oracle_ac *a, const uint8_t *text, uint64_t n, uint64_t *positions, uint64_t cap) {
	uint64_t count = 0;
	uint32_t r = 0;
	const int sigma = a->alphabet;
	for (uint64_t column = 0; column < n; column++) {
		int32_t s;
		while ((s = a->go[(size_t) r * sigma + text[column]]) < 0)
			r = a->fail[r];
		r = (uint32_t) s;
		if (a->fin[r])
			EMIT(positions, cap, count, column);
	}
	return count;
}

uint64_t oracle_ac_search(const uint8_t *patterns_flat, int m, int p, int alphabet, const uint8_t *text, uint64_t n,
		uint64_t *positions, uint64_t cap) {
	oracle_ac *a = oracle_ac_build(patterns_flat, m, p, alphabet);
	uint64_t c = oracle_ac_scan(a, text, n, positions, cap);
	oracle_ac_free(a);
	return c;
}

/* ============================== Wu-Manber ============================== */

/* wu/wu.c:18-47 */
uint32_t oracle_wu_shiftsize(int alphabet) {
	switch (alphabet) {
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

typedef struct {
	int m, p;
	uint32_t shiftsize;
	int32_t *SHIFT;        /* [shiftsize] */
	int32_t *PREFIX_size;  /* [shiftsize] */
	uint32_t *bucket_start;/* [shiftsize+1] CSR over the dense rows the reference uses */
	int32_t *PREFIX_value; /* [sum sizes] */
	int32_t *PREFIX_index; /* [sum sizes] */
	const uint8_t *patterns;
} oracle_wu;

/* The reference's block hash: three symbols, 2-bit shifts (m_nBitsInShift = 2,
 * main.c:431; wu/wu.c:63-67, 120-124). */
static inline uint32_t wu_hash3(uint32_t a, uint32_t b, uint32_t c) { return (((a << 2) + b) << 2) + c; }
static inline uint32_t wu_hash2(uint32_t a, uint32_t b) { return (a << 2) + b; }

/* preproc_wu (wu/wu.c:109-149): for every pattern j and q = m..B: hash of
 * p[q-3..q-1]; SHIFT[hash] = min(SHIFT[hash], m-q); when m-q == 0 append
 * (prefix hash of p[0..1], j) to the bucket, in pattern order.  SHIFT starts
 * at m-B+1 (main.c:447), B = 3. */
oracle_wu *oracle_wu_build(const uint8_t *patterns_flat, int m, int p, int alphabet) {
	const int B = 3;
	uint32_t ss = oracle_wu_shiftsize(alphabet);
	if (!ss || m < B)
		return NULL;
	oracle_wu *w = (oracle_wu *) calloc(1, sizeof(*w));
	w->m = m;
	w->p = p;
	w->shiftsize = ss;
	w->patterns = patterns_flat;
	w->SHIFT = (int32_t *) malloc(ss * sizeof(int32_t));
	w->PREFIX_size = (int32_t *) calloc(ss, sizeof(int32_t));
	w->bucket_start = (uint32_t *) calloc((size_t) ss + 1, sizeof(uint32_t));
	for (uint32_t i = 0; i < ss; i++)
		w->SHIFT[i] = m - B + 1;
	for (int j = 0; j < p; j++) {
		const uint8_t *s = patterns_flat + (size_t) j * m;
		for (int q = m; q >= B; --q) {
			uint32_t h = wu_hash3(s[q - 3], s[q - 2], s[q - 1]);
			int32_t shiftlen = m - q;
			if (shiftlen < w->SHIFT[h])
				w->SHIFT[h] = shiftlen;
			if (shiftlen == 0)
				w->PREFIX_size[h]++;
		}
	}
	for (uint32_t i = 0; i < ss; i++)
		w->bucket_start[i + 1] = w->bucket_start[i] + (uint32_t) w->PREFIX_size[i];
	w->PREFIX_value = (int32_t *) malloc(((size_t) p + 1) * sizeof(int32_t));
	w->PREFIX_index = (int32_t *) malloc(((size_t) p + 1) * sizeof(int32_t));
	uint32_t *fill = (uint32_t *) calloc(ss, sizeof(uint32_t));
	for (int j = 0; j < p; j++) {
		const uint8_t *s = patterns_flat + (size_t) j * m;
		uint32_t h = wu_hash3(s[m - 3], s[m - 2], s[m - 1]);
		uint32_t at = w->bucket_start[h] + fill[h]++;
		w->PREFIX_value[at] = (int32_t) wu_hash2(s[0], s[1]);
		w->PREFIX_index[at] = j;
	}
	free(fill);
	return w;
}

void oracle_wu_free(oracle_wu *w) {
	if (!w)
		return;
	free(w->SHIFT);
	free(w->PREFIX_size);
	free(w->bucket_start);
	free(w->PREFIX_value);
	free(w->PREFIX_index);
	free(w);
}

/* Dense export in the reference layout (row stride p_size, main.c:436-440); cells
 * the reference never writes are left untouched. */
void oracle_wu_export_ref_tables(const oracle_wu *w, int32_t *SHIFT, int32_t *PREFIX_value, int32_t *PREFIX_index,
		int32_t *PREFIX_size) {
	memcpy(SHIFT, w->SHIFT, w->shiftsize * sizeof(int32_t));
	memcpy(PREFIX_size, w->PREFIX_size, w->shiftsize * sizeof(int32_t));
	for (uint32_t h = 0; h < w->shiftsize; h++)
		for (int32_t i = 0; i < w->PREFIX_size[h]; i++) {
			PREFIX_value[(size_t) h * w->p + i] = w->PREFIX_value[w->bucket_start[h] + i];
			PREFIX_index[(size_t) h * w->p + i] = w->PREFIX_index[w->bucket_start[h] + i];
		}
}

/* search_wu (wu/wu.c:49-107): skip loop on the 3-symbol block hash; on SHIFT == 0
 * walk the bucket, prefix-hash filter, memcmp, first hit wins (break), column++. */
uint64_t oracle_wu_scan(const oracle_wu *w, const uint8_t *text, uint64_t n, uint64_t *positions, uint64_t cap) {
	uint64_t count = 0;
	const int m = w->m;
	uint64_t column = (uint64_t) m - 1;
	while (column < n) {
		uint32_t h1 = wu_hash3(text[column - 2], text[column - 1], text[column]);
		int32_t shift = w->SHIFT[h1];
		if (shift == 0) {
			uint32_t h2 = wu_hash2(text[column - m + 1], text[column - m + 2]);
			uint32_t b0 = w->bucket_start[h1], b1 = w->bucket_start[h1 + 1];
			for (uint32_t i = b0; i < b1; i++) {
				if ((int32_t) h2 != w->PREFIX_value[i])
					continue;
				if (memcmp(w->patterns + (size_t) w->PREFIX_index[i] * m, text + column - m + 1, (size_t) m) == 0) {
					EMIT(positions, cap, count, column);
					break;
				}
			}
			column++;
		} else
			column += (uint64_t) shift;
	}
	return count;
}

uint64_t oracle_wu_search(const uint8_t *patterns_flat, int m, int p, int alphabet, const uint8_t *text, uint64_t n,
		uint64_t *positions, uint64_t cap) {
	oracle_wu *w = oracle_wu_build(patterns_flat, m, p, alphabet);
	if (!w)
		return ~(uint64_t) 0;
	uint64_t c = oracle_wu_scan(w, text, n, positions, cap);
	oracle_wu_free(w);
	return c;
}

/* ====================== window-membership scanners ====================== */
/* Result definition derived from the reference: with Pset the
 * distinct patterns, M = { (e, P) : P in Pset, e >= len(P)-1, text[e-len(P)+1..e] == P }.
 * For equal lengths at most one P matches per e and |M| is what search_ac/search_wu
 * return.  For mixed lengths (BASELINE config 4, not expressible in the reference)
 * the result is the union over length classes of the reference's result for that
 * class, so an end position appears once per distinct matching pattern.  Positions
 * are emitted in ascending e (ties: ascending length). */

/* Brute force, O(n * p * m): the hand-checkable checker for tiny inputs. */
uint64_t oracle_naive_search(const uint8_t *patterns, const uint64_t *offsets, const uint32_t *lens, int p,
		const uint8_t *text, uint64_t n, uint64_t *positions, uint64_t cap) {
	uint64_t count = 0;
	for (uint64_t e = 0; e < n; e++) {
		/* ascending length order among distinct patterns; duplicates reported once */
		uint32_t last_len = 0;
		for (;;) {
			uint32_t best = 0;
			int bestj = -1;
			for (int j = 0; j < p; j++)
				if (lens[j] > last_len && (best == 0 || lens[j] < best)) {
					best = lens[j];
					bestj = j;
				}
			if (bestj < 0)
				break;
			last_len = best;
			if ((uint64_t) best > e + 1)
				continue;
			for (int j = 0; j < p; j++)
				if (lens[j] == best && memcmp(patterns + offsets[j], text + e + 1 - best, best) == 0) {
					EMIT(positions, cap, count, e);
					break; /* identical-length patterns matching the same window are identical */
				}
		}
	}
	return count;
}

/* Linear-time scanner for big inputs: one Rabin-Karp rolling hash per distinct
 * length, open-addressing table of pattern hashes, memcmp on hash hit. */
typedef struct {
	uint64_t hash;
	int32_t pat; /* -1 = empty */
} set_slot;

static int cmp_u32(const void *a, const void *b) {
	uint32_t x = *(const uint32_t *) a, y = *(const uint32_t *) b;
	return x < y ? -1 : x > y;
}

uint64_t oracle_set_search(const uint8_t *patterns, const uint64_t *offsets, const uint32_t *lens, int p,
		const uint8_t *text, uint64_t n, uint64_t *positions, uint64_t cap, int sort_output) {
	const uint64_t MULT = 0x9E3779B97F4A7C15ull;
	uint64_t count = 0;
	uint32_t *dl = (uint32_t *) malloc((size_t) (p > 0 ? p : 1) * sizeof(uint32_t));
	for (int j = 0; j < p; j++)
		dl[j] = lens[j];
	qsort(dl, (size_t) p, sizeof(uint32_t), cmp_u32);
	int nd = 0;
	for (int j = 0; j < p; j++)
		if (j == 0 || dl[j] != dl[j - 1])
			dl[nd++] = dl[j];
	for (int d = 0; d < nd; d++) {
		const uint32_t L = dl[d];
		if (L == 0 || (uint64_t) L > n)
			continue;
		size_t cnt = 0;
		for (int j = 0; j < p; j++)
			cnt += lens[j] == L;
		size_t cap_slots = 16;
		while (cap_slots < 4 * cnt)
			cap_slots <<= 1;
		set_slot *tab = (set_slot *) malloc(cap_slots * sizeof(set_slot));
		for (size_t i = 0; i < cap_slots; i++)
			tab[i].pat = -1;
		for (int j = 0; j < p; j++) {
			if (lens[j] != L)
				continue;
			uint64_t h = 0;
			for (uint32_t i = 0; i < L; i++)
				h = h * MULT + (uint64_t) patterns[offsets[j] + i] + 1;
			size_t s = (size_t) ((h ^ (h >> 29)) & (cap_slots - 1));
			int dup = 0;
			while (tab[s].pat >= 0) {
				if (tab[s].hash == h && memcmp(patterns + offsets[tab[s].pat], patterns + offsets[j], L) == 0) {
					dup = 1;
					break;
				}
				s = (s + 1) & (cap_slots - 1);
			}
			if (!dup) {
				tab[s].hash = h;
				tab[s].pat = j;
			}
		}
		uint64_t top = 1; /* MULT^(L-1) */
		for (uint32_t i = 1; i < L; i++)
			top *= MULT;
		uint64_t h = 0;
		for (uint32_t i = 0; i < L; i++)
			h = h * MULT + (uint64_t) text[i] + 1;
		for (uint64_t e = (uint64_t) L - 1;; e++) {
			size_t s = (size_t) ((h ^ (h >> 29)) & (cap_slots - 1));
			while (tab[s].pat >= 0) {
				if (tab[s].hash == h && memcmp(patterns + offsets[tab[s].pat], text + e + 1 - L, L) == 0) {
					EMIT(positions, cap, count, e);
					break;
				}
				s = (s + 1) & (cap_slots - 1);
			}
			if (e + 1 >= n)
				break;
			h = (h - ((uint64_t) text[e + 1 - L] + 1) * top) * MULT + (uint64_t) text[e + 1] + 1;
		}
		free(tab);
	}
	free(dl);
	if (sort_output && positions && nd > 1) {
		/* merge the per-length runs: a stable sort by e keeps ascending length among ties */
		uint64_t w = count < cap ? count : cap;
		uint64_t *tmp = (uint64_t *) malloc((size_t) (w ? w : 1) * sizeof(uint64_t));
		for (uint64_t width = 1; width < w; width <<= 1) {
			for (uint64_t lo = 0; lo < w; lo += 2 * width) {
				uint64_t mid = lo + width < w ? lo + width : w, hi = lo + 2 * width < w ? lo + 2 * width : w;
				uint64_t i = lo, j = mid, k = lo;
				while (i < mid && j < hi)
					tmp[k++] = positions[j] < positions[i] ? positions[j++] : positions[i++];
				while (i < mid)
					tmp[k++] = positions[i++];
				while (j < hi)
					tmp[k++] = positions[j++];
			}
			memcpy(positions, tmp, (size_t) w * sizeof(uint64_t));
		}
		free(tmp);
	}
	return count;
}
