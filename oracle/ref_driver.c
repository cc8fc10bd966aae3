/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Thin driver around the UNMODIFIED reference CPU functions (compiled where they
 * lie under /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so).
 * It performs the caller-side duties the reference's main.c performs before it
 * calls the algorithms, and nothing else:
 *
 *   AC  (main.c:410-418): state_transition memset to -1, state_supply/state_final
 *       zeroed, sized (m*p_size+1) [* alphabet]; patterns handed over as
 *       unsigned char*[p_size].  Patterns are allocated with m+1 bytes because
 *       ac_addstring reads string[m] for a duplicate pattern (ac/ac.c:136-143).
 *   WM  (main.c:429-449): wu_determine_shiftsize(alphabet); m_nBitsInShift = 2;
 *       SHIFT[i] = m - B + 1; PREFIX_size[i] = 0, with B = 3 (main.c:335).
 *
 * The reference returns only match COUNTS (ac/ac.c:198-222, wu/wu.c:49-107); match
 * positions come from oracle_port.c and are accepted as golden only after their
 * count was checked against the counts returned here.
 *
 * The *_mt entry points shard the text exactly like the MPI ranks of main.c:467-477
 * (chunk = ceil(n/R); rank i scans [i*chunk, min((i+1)*chunk + m-1, n))) and run one
 * reference search per POSIX thread; they are the "reference on all host cores"
 * CPU baseline of bench.py.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "smatcher.h" /* /root/reference/smatcher.h via -I; defines the globals */

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static unsigned char **make_pattern_rows(const unsigned char *flat, int m, int p) {
	unsigned char **rows = (unsigned char **) malloc((size_t) p * sizeof(*rows));
	for (int j = 0; j < p; j++) {
		rows[j] = (unsigned char *) calloc((size_t) m + 1, 1); /* +1: ac/ac.c:136 over-read */
		memcpy(rows[j], flat + (size_t) j * m, (size_t) m);
	}
	return rows;
}

static void free_pattern_rows(unsigned char **rows, int p) {
	for (int j = 0; j < p; j++)
		free(rows[j]);
	free(rows);
}

/* ---- shard geometry, main.c:467-477 (integer form of the float arithmetic) ---- */
typedef struct {
	int algo; /* 0 = ac, 1 = wu (2-D patterns), 2 = wu2 (flat patterns) */
	const unsigned char *text;
	int n;
	/* ac */
	struct ac_table *table;
	/* wu */
	unsigned char **rows;
	const unsigned char *flat;
	int m, p;
	int *SHIFT, *PV, *PI, *PS;
	unsigned long long matches;
} shard_job;

static void *shard_main(void *arg) {
	shard_job *j = (shard_job *) arg;
	if (j->n <= 0) {
		j->matches = 0;
		return NULL;
	}
	if (j->algo == 0)
		j->matches = search_ac((unsigned char *) j->text, j->n, j->table);
	else if (j->algo == 1)
		j->matches = search_wu(j->rows, j->m, j->p, (unsigned char *) j->text, j->n, j->SHIFT, j->PV, j->PI, j->PS);
	else
		j->matches = search_wu2((unsigned char *) j->flat, j->m, j->p, (unsigned char *) j->text, j->n, j->SHIFT,
				j->PV, j->PI, j->PS);
	return NULL;
}

static unsigned long long run_sharded(shard_job proto, const unsigned char *text, uint64_t n, int m, int threads,
		double *search_s) {
	if (threads < 1)
		threads = 1;
	uint64_t chunk = (n + (uint64_t) threads - 1) / (uint64_t) threads;
	shard_job *jobs = (shard_job *) calloc((size_t) threads, sizeof(*jobs));
	pthread_t *tids = (pthread_t *) calloc((size_t) threads, sizeof(*tids));
	for (int i = 0; i < threads; i++) {
		uint64_t start = (uint64_t) i * chunk;
		uint64_t stop = (uint64_t) (i + 1) * chunk + (uint64_t) (m - 1);
		if (stop > n)
			stop = n;
		jobs[i] = proto;
		jobs[i].text = text + (start < n ? start : n);
		jobs[i].n = start < stop ? (int) (stop - start) : 0;
	}
	double t0 = now_s();
	if (threads == 1) {
		shard_main(&jobs[0]);
	} else {
		for (int i = 0; i < threads; i++)
			pthread_create(&tids[i], NULL, shard_main, &jobs[i]);
		for (int i = 0; i < threads; i++)
			pthread_join(tids[i], NULL);
	}
	double t1 = now_s();
	unsigned long long total = 0;
	for (int i = 0; i < threads; i++)
		total += jobs[i].matches;
	if (search_s)
		*search_s = t1 - t0;
	free(jobs);
	free(tids);
	return total;
}

/* ------------------------------- Aho-Corasick ------------------------------- */

/* Runs preproc_ac + search_ac (+ free_ac).  Optional outputs (may be NULL):
 *   tr_out   int[(m*p+1)*alphabet]   state_transition as the reference left it
 *   sup_out  unsigned[m*p+1]         state_supply
 *   fin_out  unsigned[m*p+1]         state_final
 *   meta_out unsigned[2]             {idcounter, patterncounter}  (smatcher.h:49-53)
 *   times    double[2]               {preproc seconds, search seconds}
 * threads > 1 shards the text like main.c:467-477. */
unsigned long long ref_ac_search(const unsigned char *patterns_flat, int m, int p, int alphabet,
		const unsigned char *text, uint64_t n, int threads, int *tr_out, unsigned *sup_out, unsigned *fin_out,
		unsigned *meta_out, double *times) {
	size_t states = (size_t) m * (size_t) p + 1;
	int *tr = tr_out ? tr_out : (int *) malloc(states * (size_t) alphabet * sizeof(int));
	unsigned *sup = sup_out ? sup_out : (unsigned *) malloc(states * sizeof(unsigned));
	unsigned *fin = fin_out ? fin_out : (unsigned *) malloc(states * sizeof(unsigned));
	memset(tr, -1, states * (size_t) alphabet * sizeof(int)); /* main.c:412 */
	memset(sup, 0, states * sizeof(unsigned)); /* main.c:416 */
	memset(fin, 0, states * sizeof(unsigned)); /* main.c:420 */
	unsigned char **rows = make_pattern_rows(patterns_flat, m, p);

	double t0 = now_s();
	struct ac_table *table = preproc_ac(rows, m, p, alphabet, tr, sup, fin);
	double t1 = now_s();

	shard_job proto;
	memset(&proto, 0, sizeof(proto));
	proto.algo = 0;
	proto.table = table;
	double ts = 0;
	unsigned long long matches = run_sharded(proto, text, n, m, threads, &ts);

	if (meta_out) {
		meta_out[0] = table->idcounter;
		meta_out[1] = table->patterncounter;
	}
	if (times) {
		times[0] = t1 - t0;
		times[1] = ts;
	}
	free_ac(table, alphabet);
	free_pattern_rows(rows, p);
	if (!tr_out)
		free(tr);
	if (!sup_out)
		free(sup);
	if (!fin_out)
		free(fin);
	return matches;
}

/* -------------------------------- Wu-Manber -------------------------------- */

/* Table size the reference uses for this alphabet (wu/wu.c:18-47); 0 if the
 * alphabet is one the reference would fail() on. */
unsigned ref_wu_shiftsize(int alphabet) {
	switch (alphabet) {
	case 2: case 4: case 8: case 20: case 128: case 256: case 512: case 1024:
		wu_determine_shiftsize(alphabet);
		return shiftsize;
	default:
		return 0;
	}
}

/* Runs preproc_wu + search_wu (flat == 0) or preproc_wu2 + search_wu2 (flat != 0).
 * Optional outputs (may be NULL): SHIFT_out/PS_out int[shiftsize],
 * PV_out/PI_out int[shiftsize*p] (dense, row stride p as in main.c:436-440),
 * times double[2] = {preproc seconds, search seconds}.
 * Returns ~0ull if the alphabet is unsupported (the reference would exit). */
unsigned long long ref_wu_search(const unsigned char *patterns_flat, int m, int p, int alphabet, int flat,
		const unsigned char *text, uint64_t n, int threads, int *SHIFT_out, int *PV_out, int *PI_out, int *PS_out,
		double *times) {
	const int B = 3; /* main.c:335 */
	unsigned ss = ref_wu_shiftsize(alphabet);
	if (ss == 0)
		return ~0ull;
	m_nBitsInShift = 2; /* main.c:431 */
	int *SHIFT = SHIFT_out ? SHIFT_out : (int *) malloc((size_t) ss * sizeof(int));
	int *PS = PS_out ? PS_out : (int *) malloc((size_t) ss * sizeof(int));
	int *PV = PV_out ? PV_out : (int *) malloc((size_t) ss * (size_t) p * sizeof(int));
	int *PI = PI_out ? PI_out : (int *) malloc((size_t) ss * (size_t) p * sizeof(int));
	for (unsigned i = 0; i < ss; i++) { /* main.c:444-449 */
		SHIFT[i] = m - B + 1;
		PS[i] = 0;
	}
	unsigned char **rows = make_pattern_rows(patterns_flat, m, p);

	double t0 = now_s();
	if (flat)
		preproc_wu2((unsigned char *) patterns_flat, m, p, alphabet, B, SHIFT, PV, PI, PS);
	else
		preproc_wu(rows, m, p, alphabet, B, SHIFT, PV, PI, PS);
	double t1 = now_s();

	shard_job proto;
	memset(&proto, 0, sizeof(proto));
	proto.algo = flat ? 2 : 1;
	proto.rows = rows;
	proto.flat = patterns_flat;
	proto.m = m;
	proto.p = p;
	proto.SHIFT = SHIFT;
	proto.PV = PV;
	proto.PI = PI;
	proto.PS = PS;
	double ts = 0;
	unsigned long long matches = run_sharded(proto, text, n, m, threads, &ts);
	if (times) {
		times[0] = t1 - t0;
		times[1] = ts;
	}
	free_pattern_rows(rows, p);
	if (!SHIFT_out)
		free(SHIFT);
	if (!PS_out)
		free(PS);
	if (!PV_out)
		free(PV);
	if (!PI_out)
		free(PI);
	return matches;
}
