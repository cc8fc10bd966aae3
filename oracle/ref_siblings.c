/* TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference sh/sh.c (Set-Horspool) and sbom/sbom.c (Set Backward
 * Oracle Matching), compiled where they lie under /root/reference by oracle/Makefile, so that the sibling shims of the
 * library (search_sh / search_sbom behind the same matcher, include/acwm.h) can be checked count for count.
 * Caller duties restated from main.c:157-232,407-427: state_transition pre-filled with -1, state_final / state_final_multi
 * zeroed, pointer_array allocated (sbom), bmBc from preBmBc -- a helper the reference calls (main.c:173) but does not ship;
 * restated here as the classic Set-Horspool bad-character table: bmBc[c] = min over patterns of the distance from the last
 * occurrence of c in pattern[0 .. m-2] to the pattern end, m when c does not occur. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "smatcher.h"

static unsigned char **rows_of(const unsigned char *flat, int m, int p) {
	unsigned char **rows = (unsigned char **) malloc((size_t) p * sizeof(*rows));
	for (int i = 0; i < p; i++) {
		rows[i] = (unsigned char *) malloc((size_t) m + 1);
		memcpy(rows[i], flat + (size_t) i * m, (size_t) m);
		rows[i][m] = 0;
	}
	return rows;
}
static void free_rows(unsigned char **rows, int p) {
	for (int i = 0; i < p; i++)
		free(rows[i]);
	free(rows);
}

static void pre_bm_bc(unsigned char **pattern, int m, int p, int alphabet, int *bmBc) {
	for (int c = 0; c < alphabet; c++)
		bmBc[c] = m;
	for (int j = 0; j < p; j++)
		for (int i = 0; i < m - 1; i++)
			if (m - 1 - i < bmBc[pattern[j][i]])
				bmBc[pattern[j][i]] = m - 1 - i;
}

unsigned long long ref_sh_search(const unsigned char *patterns_flat, int m, int p, int alphabet, const unsigned char *text, int n) {
	unsigned char **rows = rows_of(patterns_flat, m, p);
	const size_t states = (size_t) m * p + 1;
	int *state_transition = (int *) malloc(states * alphabet * sizeof(int));
	memset(state_transition, -1, states * alphabet * sizeof(int));
	unsigned int *state_final = (unsigned int *) calloc(states, sizeof(unsigned int));
	int *bmBc = (int *) malloc((size_t) alphabet * sizeof(int));
	pre_bm_bc(rows, m, p, alphabet, bmBc);
	struct ac_table *t = preproc_sh(rows, m, p, alphabet, state_transition, state_final);
	const unsigned long long c = search_sh(m, (unsigned char *) text, n, t, bmBc);
	free_sh(t, alphabet);
	free(bmBc);
	free(state_final);
	free(state_transition);
	free_rows(rows, p);
	return c;
}

unsigned long long ref_sbom_search(const unsigned char *patterns_flat, int m, int p, int alphabet, const unsigned char *text, int n) {
	unsigned char **rows = rows_of(patterns_flat, m, p);
	const size_t states = (size_t) m * p + 1;
	int *state_transition = (int *) malloc(states * alphabet * sizeof(int));
	memset(state_transition, -1, states * alphabet * sizeof(int));
	unsigned int *state_final_multi = (unsigned int *) calloc(states * 200, sizeof(unsigned int));
	pointer_array = malloc((size_t) p * m * sizeof(struct sbom_state));
	struct sbom_table *t = preproc_sbom(rows, m, p, alphabet, state_transition, state_final_multi);
	const unsigned long long c = search_sbom(rows, m, (unsigned char *) text, n, t);
	free_sbom(t, m);
	free(pointer_array);
	free(state_final_multi);
	free(state_transition);
	free_rows(rows, p);
	return c;
}

/* The flat tables the reference's preprocessing leaves in its caller's arrays (pre-initialised here as main.c:410-426
 * does: state_transition -1, the others 0); returns the number of states. */
unsigned ref_sh_tables(const unsigned char *patterns_flat, int m, int p, int alphabet, int *state_transition,
		unsigned int *state_final, unsigned *n_distinct) {
	unsigned char **rows = rows_of(patterns_flat, m, p);
	const size_t states = (size_t) m * p + 1;
	memset(state_transition, -1, states * alphabet * sizeof(int));
	memset(state_final, 0, states * sizeof(unsigned int));
	struct ac_table *t = preproc_sh(rows, m, p, alphabet, state_transition, state_final);
	const unsigned ns = t->idcounter;
	if (n_distinct)
		*n_distinct = t->patterncounter;
	free_sh(t, alphabet);
	free_rows(rows, p);
	return ns;
}

unsigned ref_sbom_tables(const unsigned char *patterns_flat, int m, int p, int alphabet, int *state_transition,
		unsigned int *state_final_multi) {
	unsigned char **rows = rows_of(patterns_flat, m, p);
	const size_t states = (size_t) m * p + 1;
	memset(state_transition, -1, states * alphabet * sizeof(int));
	memset(state_final_multi, 0, states * 200 * sizeof(unsigned int));
	pointer_array = malloc((size_t) p * m * sizeof(struct sbom_state));
	struct sbom_table *t = preproc_sbom(rows, m, p, alphabet, state_transition, state_final_multi);
	const unsigned ns = t->idcounter;
	free_sbom(t, m);
	free(pointer_array);
	free_rows(rows, p);
	return ns;
}
