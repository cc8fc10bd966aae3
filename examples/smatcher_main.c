/* smatcher_main.c -- a C driver that uses libacwm_b200.so exactly the way the reference's
 * main.c uses ac/ac.c, wu/wu.c and cuda/cuda_{ac,wm}.cu: same entry points, same
 * caller-side allocation and initialisation of the flat tables, same report lines.
 *
 *   smatcher_main <ac|wm> -m M -n N -p_size P -alphabet A [-text FILE -pattern FILE] [-data DIR [-c]] [-seed S]
 *                 [-gpus G]
 *
 * It is the integration example of INTEGRATION.md (what a maintainer of the reference
 * links instead of the reference's own objects) and the small CLI equivalent of
 * main.c:324-718 without MPI: the algorithm name restores the dispatch the reference
 * has commented out (main.c:519-531), the text/pattern files replace the missing
 * load_files helper (main.c:453) and, without files, a seeded generator plays the role
 * of create_multiple_pattern_with_hits (main.c:49): half of the patterns are windows
 * of the text.  Text and pattern bytes are symbol codes in [0, alphabet); a -text file that
 * holds a raw corpus (FASTA nucleotides / amino acids, English text) is mapped to codes by
 * the library's loader (acwm_load_text).  With -data DIR the corpus and the pattern file
 * are chosen the way the reference's select_data_file does (main.c:31-123: the text size
 * -n selects the corpus under DIR/text, patterns in DIR/pattern/<n>/<m>/<alphabet>/pattern);
 * -c draws the pattern set "with hits" from the text instead of reading it (main.c:49).
 * -gpus G adds the multi-rank flow of main.c:464-656 in one process: G shards with an (m-1)-byte halo
 * (MPI_Scatterv), one matcher and one host thread per shard, spread over the CUDA devices present
 * (acwm_search_host_sharded), the counts summed (MPI_Reduce) and printed per rank and in total.
 *
 * Build:  gcc -O2 -I include examples/smatcher_main.c -L cuda-aho-corasick-wu-manber_b200 \
 *             -lacwm_b200 -Wl,-rpath,'$ORIGIN/../cuda-aho-corasick-wu-manber_b200' -o examples/smatcher_main
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "acwm.h" /* declares the smatcher.h entry points + cuda_acN / cuda_wmN */

static void usage(void) {
	fprintf(stderr, "usage: smatcher_main <ac|wm|sh|sbom> -m M -n N -p_size P -alphabet A [-text FILE -pattern FILE] [-data DIR [-c]] [-seed S] [-gpus G]\n");
	exit(2);
}

static unsigned long long rng_state = 88172645463325252ull;
static unsigned long long rng_next(void) { /* xorshift64* */
	rng_state ^= rng_state >> 12;
	rng_state ^= rng_state << 25;
	rng_state ^= rng_state >> 27;
	return rng_state * 2685821657736338717ull;
}

static void read_exact(const char *path, unsigned char *dst, size_t bytes) {
	FILE *f = fopen(path, "rb");
	if (!f || fread(dst, 1, bytes, f) != bytes) {
		fprintf(stderr, "cannot read %zu bytes from %s\n", bytes, path);
		exit(1);
	}
	fclose(f);
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
	int m = 0, n = 0, p_size = 0, alphabet = 0, B = 3, i, j;
	const char *text_file = NULL, *pattern_file = NULL, *data_root = NULL;
	int create_data = 0, gpus = 0;
	unsigned long long seed = 0;
	static char sel_text[4096], sel_pattern[4096];
	if (argc < 2 || (strcmp(argv[1], "ac") && strcmp(argv[1], "wm") && strcmp(argv[1], "sh") && strcmp(argv[1], "sbom")))
		usage();
	const int use_sh = strcmp(argv[1], "sh") == 0, use_sbom = strcmp(argv[1], "sbom") == 0;
	const int use_ac = strcmp(argv[1], "ac") == 0 || use_sh || use_sbom; /* the siblings run on the automaton matcher */
	for (i = 2; i + 1 < argc; i++) {
		if (!strcmp(argv[i], "-m")) m = atoi(argv[i + 1]);
		if (!strcmp(argv[i], "-n")) n = atoi(argv[i + 1]);
		if (!strcmp(argv[i], "-p_size")) p_size = atoi(argv[i + 1]);
		if (!strcmp(argv[i], "-alphabet")) alphabet = atoi(argv[i + 1]);
		if (!strcmp(argv[i], "-text")) text_file = argv[i + 1];
		if (!strcmp(argv[i], "-pattern")) pattern_file = argv[i + 1];
		if (!strcmp(argv[i], "-data")) data_root = argv[i + 1];
		if (!strcmp(argv[i], "-gpus")) gpus = atoi(argv[i + 1]);
		if (!strcmp(argv[i], "-seed")) {
			seed = (unsigned long long) atoll(argv[i + 1]);
			rng_state ^= seed * 0x9E3779B97F4A7C15ull;
		}
	}
	for (i = 2; i < argc; i++)
		if (!strcmp(argv[i], "-c"))
			create_data = 1; /* main.c:362: regenerate the pattern set with hits */
	if (m <= 0 || n <= 0 || p_size <= 0 || alphabet <= 0)
		usage();
	if (data_root) { /* select_data_file, main.c:31-123 */
		if (acwm_select_data_file((uint32_t) m, (uint64_t) n, (uint32_t) alphabet, data_root, sel_pattern, sel_text,
					sizeof(sel_text)) != ACWM_OK) {
			fprintf(stderr, "%s\n", acwm_last_error());
			return 1;
		}
		text_file = sel_text;
		pattern_file = create_data ? NULL : sel_pattern;
	}

	/* ---- inputs: text[n], pattern[p_size][m] (+1 byte: the reference reads one past, ac/ac.c:136) and the flat copy */
	unsigned char *text = (unsigned char *) malloc((size_t) n);
	unsigned char **pattern = (unsigned char **) malloc((size_t) p_size * sizeof(unsigned char *));
	unsigned char *pattern2 = (unsigned char *) malloc((size_t) m * p_size);
	if (!text || !pattern || !pattern2)
		return 1;
	if (text_file) { /* load_files, main.c:453: the first n symbols of the corpus, as codes */
		uint8_t *loaded = NULL;
		uint64_t got = 0;
		if (acwm_load_text(text_file, (uint32_t) alphabet, (uint64_t) n, &loaded, &got) != ACWM_OK || got < (uint64_t) n) {
			fprintf(stderr, "cannot load %d symbols from %s: %s\n", n, text_file, loaded ? "corpus too short" : acwm_last_error());
			return 1;
		}
		memcpy(text, loaded, (size_t) n);
		acwm_free_text(loaded);
	} else
		for (i = 0; i < n; i++)
			text[i] = (unsigned char) (rng_next() % (unsigned) alphabet);
	if (pattern_file)
		read_exact(pattern_file, pattern2, (size_t) m * p_size);
	else if (data_root) /* create_multiple_pattern_with_hits, main.c:49 */
		acwm_patterns_with_hits(text, (uint64_t) n, (uint32_t) m, (uint32_t) p_size, (uint32_t) alphabet, seed, 50, pattern2);
	else
		for (j = 0; j < p_size; j++) {
			if (j % 2 == 0 && n >= m) { /* "with hits" */
				const size_t at = (size_t) (rng_next() % (unsigned long long) (n - m + 1));
				memcpy(pattern2 + (size_t) j * m, text + at, (size_t) m);
			} else
				for (i = 0; i < m; i++)
					pattern2[(size_t) j * m + i] = (unsigned char) (rng_next() % (unsigned) alphabet);
		}
	for (j = 0; j < p_size; j++) {
		pattern[j] = (unsigned char *) calloc((size_t) m + 1, 1);
		memcpy(pattern[j], pattern2 + (size_t) j * m, (size_t) m);
	}

	if (use_sh) {
		/* multish (main.c:158-195): the caller's tables (main.c:408-427), preprocess, search, free */
		const size_t states = (size_t) m * p_size + 1;
		int *state_transition = (int *) malloc(states * alphabet * sizeof(int));
		unsigned int *state_final = (unsigned int *) calloc(states, sizeof(unsigned int));
		int *bmBc = (int *) calloc((size_t) alphabet, sizeof(int)); /* preBmBc (main.c:173) steers the CPU skip loop only */
		memset(state_transition, -1, states * alphabet * sizeof(int));
		double t = now_s();
		struct ac_table *table = preproc_sh(pattern, m, p_size, alphabet, state_transition, state_final);
		const double t_pre = now_s() - t;
		t = now_s();
		const unsigned matches = search_sh(m, text, n, table, bmBc);
		printf("search_sh matches \t%u\t time \t%f\n", matches, now_s() - t);
		printf("preproc_sh states \t%u\t patterns \t%u\t time \t%f\n", table->idcounter, table->patterncounter, t_pre);
		printf("Total results: %u.\n", matches);
		free_sh(table, alphabet);
		free(state_transition);
		free(state_final);
		free(bmBc);
	} else if (use_sbom) {
		/* multisbom (main.c:197-233) */
		const size_t states = (size_t) m * p_size + 1;
		int *state_transition = (int *) malloc(states * alphabet * sizeof(int));
		unsigned int *state_final_multi = (unsigned int *) calloc(states * 200, sizeof(unsigned int));
		memset(state_transition, -1, states * alphabet * sizeof(int));
		double t = now_s();
		struct sbom_table *table = preproc_sbom(pattern, m, p_size, alphabet, state_transition, state_final_multi);
		const double t_pre = now_s() - t;
		t = now_s();
		const unsigned matches = search_sbom(pattern, m, text, n, table);
		printf("search_sbom matches \t%u\t time \t%f\n", matches, now_s() - t);
		printf("preproc_sbom states \t%u\t patterns \t%u\t time \t%f\n", table->idcounter, table->patterncounter, t_pre);
		printf("Total results: %u.\n", matches);
		free_sbom(table, m);
		free(state_transition);
		free(state_final_multi);
	} else if (use_ac) {
		/* ---- caller-side tables of main.c:408-420 */
		const size_t states = (size_t) m * p_size + 1;
		int *state_transition = (int *) malloc(states * alphabet * sizeof(int));
		unsigned int *state_supply = (unsigned int *) calloc(states, sizeof(unsigned int));
		unsigned int *state_final = (unsigned int *) calloc(states, sizeof(unsigned int));
		memset(state_transition, -1, states * alphabet * sizeof(int));
		/* multiac (main.c:125-157): preprocess, search, free */
		double t = now_s();
		struct ac_table *table = preproc_ac(pattern, m, p_size, alphabet, state_transition, state_supply, state_final);
		const double t_pre = now_s() - t;
		t = now_s();
		const unsigned matches = search_ac(text, n, table);
		printf("search_ac matches \t%u\t time \t%f\n", matches, now_s() - t);
		printf("preproc_ac states \t%u\t patterns \t%u\t time \t%f\n", table->idcounter, table->patterncounter, t_pre);
		free_ac(table, alphabet);
		/* the GPU wrapper works from the flat tables alone (main.c:582-593) */
		cuda_ac5(m, text, n, p_size, alphabet, state_transition, state_supply, state_final);
		printf("Total results: %llu.\n", acwm_shim_last_count());
		free(state_transition);
		free(state_supply);
		free(state_final);
	} else {
		/* ---- caller-side tables of main.c:429-449 */
		wu_determine_shiftsize(alphabet);
		m_nBitsInShift = 2;
		int *SHIFT = (int *) malloc(shiftsize * sizeof(int));
		int *PREFIX_value = (int *) malloc((size_t) shiftsize * p_size * sizeof(int));
		int *PREFIX_index = (int *) malloc((size_t) shiftsize * p_size * sizeof(int));
		int *PREFIX_size = (int *) malloc(shiftsize * sizeof(int));
		for (i = 0; i < (int) shiftsize; i++) {
			SHIFT[i] = m - B + 1;
			PREFIX_size[i] = 0;
		}
		/* multiwm2 (main.c:268-298) */
		double t = now_s();
		preproc_wu2(pattern2, m, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size);
		const double t_pre = now_s() - t;
		t = now_s();
		const unsigned matches = search_wu2(pattern2, m, p_size, text, n, SHIFT, PREFIX_value, PREFIX_index, PREFIX_size);
		printf("search_wm2 matches \t%u\t time \t%f\n", matches, now_s() - t);
		printf("preproc_wu2 time \t%f\n", t_pre);
		/* cuda_wm5 (main.c:646-647) */
		double gpuTime = 0;
		const int result = cuda_wm5(pattern2, m, text, n, p_size, alphabet, B, SHIFT, PREFIX_value, PREFIX_index,
				PREFIX_size, &gpuTime);
		printf("Kernel 5 matches \t%d\t time \t%f\n", result, gpuTime);
		printf("Total results: %d.\n", result);
		free(SHIFT);
		free(PREFIX_value);
		free(PREFIX_index);
		free(PREFIX_size);
	}
	if (gpus > 0) { /* the MPI ranks of main.c:464-656 as shards of one process */
		acwm_matcher **mts = (acwm_matcher **) calloc((size_t) gpus, sizeof(acwm_matcher *));
		uint64_t *per = (uint64_t *) calloc((size_t) gpus, sizeof(uint64_t));
		uint64_t total = 0;
		int rc = ACWM_OK, r;
		const int n_dev = acwm_device_count();
		for (r = 0; r < gpus && rc == ACWM_OK; r++) {
			rc = acwm_build(use_ac ? ACWM_ALGO_AC : ACWM_ALGO_WM, pattern2, NULL, (uint32_t) m, (uint32_t) p_size,
					(uint32_t) alphabet, NULL, &mts[r]);
			if (rc == ACWM_OK && n_dev > 0)
				rc = acwm_upload(mts[r], r % n_dev, 0);
		}
		if (rc == ACWM_OK) {
			const double t = now_s();
			rc = acwm_search_host_sharded(mts, (uint32_t) gpus, text, (uint64_t) n, &total, NULL, 0, NULL, per);
			const double dt = now_s() - t;
			if (rc == ACWM_OK) {
				for (r = 0; r < gpus; r++) {
					uint64_t start = 0, len = 0;
					acwm_shard_bounds((uint64_t) n, (uint32_t) gpus, (uint32_t) r, (uint32_t) m - 1, &start, &len);
					printf("rank %d device %d text [%llu, %llu) matches \t%llu\t gpuTime \t%f\n", r, n_dev ? r % n_dev : -1,
							(unsigned long long) start, (unsigned long long) (start + len), (unsigned long long) per[r],
							acwm_last_kernel_seconds(mts[r]));
				}
				printf("Total results: %llu.\n", (unsigned long long) total);
				printf("timeScatterExecuteGather: %f.\n", dt);
			}
		}
		if (rc != ACWM_OK)
			fprintf(stderr, "sharded search failed: %s\n", acwm_last_error());
		for (r = 0; r < gpus; r++)
			if (mts[r])
				acwm_free(mts[r]);
		free(mts);
		free(per);
		if (rc != ACWM_OK)
			return 1;
	}
	for (j = 0; j < p_size; j++)
		free(pattern[j]);
	free(pattern);
	free(pattern2);
	free(text);
	return 0;
}
