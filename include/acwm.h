/* acwm.h -- C ABI of libacwm_b200.so: B200-native Aho-Corasick / Wu-Manber
 * multi-pattern scan (count + every match position), the drop-in for the search
 * path of iassael/cuda-aho-corasick-wu-manber.
 *
 * Plain C, `extern "C"`, pointers + sizes only.  Two layers:
 *
 *   1. NATIVE handle API (acwm_*): build tables once on the host, upload once,
 *      scan host- or device-resident text of any size (64-bit), get the count and
 *      the sorted match positions.  Status codes, never exit().
 *
 *   2. REFERENCE-SHAPED SHIMS with the exact names and parameter lists of the
 *      reference's entry points, so the reference's main.c links against this
 *      library unchanged:
 *        preproc_ac / search_ac / free_ac            smatcher.h:89-91, ac/ac.c:198-252
 *        preproc_wu / preproc_wu2                    smatcher.h:101-102, wu/wu.c:109,211
 *        search_wu / search_wu2                      smatcher.h:105-106, wu/wu.c:49,151
 *        wu_determine_shiftsize                      smatcher.h:103, wu/wu.c:18
 *        cuda_ac1..cuda_ac5                          cuda/cuda_ac.cu:594,691,788,885,983
 *        cuda_wm1..cuda_wm5                          cuda/cuda_wm.cu:183,438,652,854,1060
 *      The shims run the scan on the GPU (there is no CPU search path in this
 *      library); like the reference (cuda/cuda.h:26-47) they print and exit(1) on
 *      a CUDA failure because their signatures have no error channel.
 *
 * Result definition (derived from ac/ac.c:207-219 and
 * wu/wu.c:163-206): with Pset the DISTINCT patterns,
 *     M = { (e, P) : P in Pset, len(P)-1 <= e < n, text[e-len(P)+1 .. e] == P }.
 * count = |M|; positions = the e of every element of M, ascending; e is the 0-based
 * index of the LAST byte of the occurrence (`column`, ac/ac.c:217, wu/wu.c:93).
 * With equal-length patterns (the only case the reference supports) at most one
 * pattern matches per e and count equals what search_ac / search_wu return.
 *
 * Text and pattern bytes are symbol codes in [0, alphabet), one byte per symbol,
 * exactly as in the reference (ac/ac.c:136,209; wu/wu.c:63-67): DNA is bytes 0..3.
 */
#ifndef ACWM_H
#define ACWM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACWM_VERSION 1

/* ------------------------------ status codes ------------------------------ */
enum {
	ACWM_OK = 0,
	ACWM_ERR_INVALID = 1,     /* bad argument (NULL, m < 1, symbol >= alphabet in a pattern, ...) */
	ACWM_ERR_UNSUPPORTED = 2, /* valid request this build does not serve (e.g. mixed-length AC) */
	ACWM_ERR_CUDA = 3,        /* CUDA runtime/driver failure; acwm_last_error() has the string */
	ACWM_ERR_NOMEM = 4,       /* host or device allocation failed */
	ACWM_ERR_OVERFLOW = 5,    /* more matches than the position capacity: count is exact, positions are not */
	ACWM_ERR_BAD_TEXT = 6     /* a text byte >= 4 was met on the 2-bit (alphabet <= 4) path */
};

enum { ACWM_ALGO_AC = 0, ACWM_ALGO_WM = 1 };

typedef struct acwm_matcher acwm_matcher; /* opaque: host tables + (after upload) device tables and buffers */

/* Build-time options; zero-initialise for defaults. */
typedef struct acwm_options {
	uint32_t smem_table_budget; /* bytes of shared memory the scan tables may take per SM; 0 = default */
	uint32_t force_stride;      /* AC: symbols per DFA lookup (1,2,3); WM: sampling stride (1,2,4,8,16); 0 = auto */
	uint32_t force_depth;       /* AC: truncate the automaton at this depth (candidates are verified); 0 = auto */
	uint32_t force_bytes_path;  /* 1 = use the byte-per-symbol kernels even when alphabet <= 4 */
	uint32_t force_threads;     /* threads per CTA of the scan kernel: 128..1024 (tuning); 0 = auto */
	uint32_t force_stages;      /* ring depth of the per-warp TMA tile pipeline: 1..4 (tuning); 0 = auto */
	uint32_t force_f2_bits;     /* log2 of the stage-2 bitmap size in bits, 13..19 (tuning); 0 = auto */
	uint32_t force_r_bits;      /* log2 of the number of WM offset-mask entries, 10..16 (tuning); 0 = auto */
	uint32_t force_smem_tables; /* 1 = never place a scan table in global memory (L2), whatever the set size */
	uint32_t force_front;       /* AC, 2-bit path: 1 = the automaton walks every symbol, 2 = sampled block filter in front and the
	                             * automaton walks candidate windows only; 0 = auto (cost model) */
	uint32_t force_ctas;        /* CTAs per SM of the scan kernel: 1, or 2 (half-size CTAs of consecutive scans share every SM
	                             * in overlap mode; 2-bit path only); 0 = auto */
} acwm_options;

/* What the builder chose; for reports and tests. */
typedef struct acwm_info {
	uint32_t algo, alphabet, n_patterns, n_distinct, m_min, m_max;
	uint32_t packed2bit;    /* 1: alphabet <= 4 path (text packed to 2 bits/symbol inside the kernel) */
	uint32_t stride;        /* AC: symbols per lookup; WM: sampling stride s */
	uint32_t depth;         /* AC: automaton depth D (== m_max: exact); WM: block length B (symbols) */
	uint32_t exact_front;   /* 1: the front-end alone is exact (no verification stage) */
	uint32_t n_states;      /* AC: trie states of the FULL automaton (idcounter, smatcher.h:50) */
	uint32_t n_rows;        /* AC: rows of the device DFA */
	uint32_t table_in_smem; /* 1: front-end table lives in shared memory, 0: global/L2 (access-policy window) */
	uint32_t smem_bytes;    /* dynamic shared memory per CTA of the scan kernel */
	uint64_t table_bytes;   /* bytes of device tables */
	uint32_t threads;       /* threads per CTA of the scan kernel */
	uint32_t stages;        /* tiles in flight per warp (TMA ring depth) */
	uint32_t ctas_per_sm;   /* 1: one CTA owns the SM; 2: two half-size CTAs share it (acwm_set_overlap) */
	uint32_t front_kind;    /* 0: the automaton / the block filter named by `algo` walks every symbol;
	                         * 1: AC behind a sampled block filter -- the automaton walks candidate windows only */
} acwm_info;

/* ------------------------------ native API ------------------------------ */

/* Compile the pattern set into scan tables (host only; no CUDA call).
 *   patterns: all pattern bytes back to back; lens == NULL -> p patterns of m bytes
 *   each (the reference's flat `pattern2` layout, main.c:455-461), else lens[j] bytes
 *   for pattern j (mixed lengths: WM only).  Duplicates are allowed and collapse.
 * Replaces preproc_ac (ac/ac.c:224) / preproc_wu (wu/wu.c:109). */
int acwm_build(int algo, const uint8_t *patterns, const uint32_t *lens, uint32_t m, uint32_t p, uint32_t alphabet,
		const acwm_options *opts, acwm_matcher **out);

/* Upload the tables to `device` (-1 = current) and allocate the per-matcher
 * device buffers.  Implicit on first search.  pos_capacity = how many match
 * positions the device staging buffers can hold (0 = count-only matcher). */
int acwm_upload(acwm_matcher *mt, int device, uint64_t pos_capacity);

/* Scan DEVICE-resident text (resident in HBM; no host<->device copy of the text).
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 * Results stay on the device until acwm_fetch().  d_text needs no alignment.
 * want_positions: 0 = count only, 1 = count + sorted positions.
 * report_from: matches whose end index is < report_from are not reported (0 = report
 * everything).  The sharding layer passes m_max-1 for every shard but the first so
 * that, with mixed-length patterns too, each match is reported by exactly one shard;
 * with equal-length patterns that is what a plain scan of the shard does anyway.
 * Replaces the kernel launch inside cuda_ac5 / cuda_wm5 (cuda/cuda_ac.cu:654,
 * cuda/cuda_wm.cu:276). */
int acwm_scan_device(acwm_matcher *mt, const uint8_t *d_text, uint64_t n, uint64_t report_from, int want_positions,
		void *stream);

/* Wait for the last scan on `stream` and fetch its results.  positions may be NULL
 * (count only); at most cap positions are written; *n_written receives how many.
 * Returns ACWM_ERR_OVERFLOW if count exceeded the capacity (count is still exact),
 * ACWM_ERR_BAD_TEXT if the text held a byte >= 4 on the 2-bit path. */
int acwm_fetch(acwm_matcher *mt, uint64_t *count, uint64_t *positions, uint64_t cap, uint64_t *n_written,
		void *stream);

/* Device pointers of the last scan's results, for callers that keep everything on
 * the GPU (the multi-GPU layer all-reduces *d_count over NCCL): d_count points to
 * one uint64 match count, d_positions to the sorted uint64 positions. */
int acwm_result_device_ptrs(acwm_matcher *mt, uint64_t **d_count, uint64_t **d_positions);

/* Scan HOST text end to end: chunked, pinned, double-buffered H2D overlapped with
 * the scan, count + sorted positions back in host memory.  The call a reference
 * user makes instead of cuda_ac5 / cuda_wm5; also what the search_* shims call. */
int acwm_search_host(acwm_matcher *mt, const uint8_t *text, uint64_t n, uint64_t *count, uint64_t *positions,
		uint64_t cap, uint64_t *n_written);

/* The host-side packer acwm_search_host puts in front of the H2D copy when the matcher was built for alphabet <= 4
 * and the text has >= 4 Mi symbols (a quarter of the bytes crosses PCIe; the scan kernel takes the packed tiles as
 * they are): packed[i] = text[4i] | text[4i+1] << 2 | text[4i+2] << 4 | text[4i+3] << 6 for ceil(n/4) bytes, on all
 * host cores; *bad_text = 1 if a byte >= 4 was met.  Exported for tests and for callers that keep packed corpora.
 * Environment: ACWM_HOST_PACK=0 makes acwm_search_host copy the text one byte per symbol instead. */
int acwm_pack_text_2bit(const uint8_t *text, uint64_t n, uint8_t *packed, int *bad_text);

/* Seconds of GPU time (CUDA events around the kernels only, as the reference times
 * its kernels: cuda/cuda_wm.cu:264-289) of the last acwm_search_host call. */
double acwm_last_kernel_seconds(const acwm_matcher *mt);
/* Bytes of text the last acwm_search_host call copied host -> device (n, or n/4 when the host packer ran). */
uint64_t acwm_last_h2d_bytes(const acwm_matcher *mt);

/* Back-to-back scans of device-resident text: with overlap on, acwm_scan_device launches its
 * kernel as a programmatic dependent launch (griddepcontrol): CTAs of the next scan take over
 * SMs as the CTAs of the previous one retire and run their read-only phase (table and tile
 * loads, filtering) until their first write, where they wait for the previous scan to complete.
 * Measured +8 % on back-to-back 128 MiB scans.  The caller guarantees that the TEXT of a scan is
 * not produced by the kernel that immediately precedes the scan on that stream (an earlier scan
 * of this matcher, a copy or an event wait are fine).  Off by default. */
int acwm_set_overlap(acwm_matcher *mt, int on);

/* Multi-GPU count exchange inside the scan kernel (replaces MPI_Reduce of the count, main.c:656,
 * and the NCCL all-reduce it would otherwise take): every rank owns a zero-initialised mailbox of
 * uint64[4][world] in peer-accessible device memory (NVLink; e.g. torch symmetric memory);
 * mailboxes[r] is rank r's mailbox AS MAPPED IN THIS PROCESS.  From then on every
 * acwm_scan_device stores the launch's count into all mailboxes over NVLink and sums the
 * previous scan's mailbox (the exchange runs one scan behind the scans, so no GPU waits for
 * another in lock step); acwm_fetch_global_count completes the exchange of the LAST scan and
 * returns its sum over the ranks.  All ranks must issue the same sequence of acwm_scan_device
 * calls (SPMD), like a collective.  world <= 1 or NULL switches it off. */
int acwm_set_peers(acwm_matcher *mt, uint32_t rank, uint32_t world, const uint64_t *mailboxes);
int acwm_fetch_global_count(acwm_matcher *mt, uint64_t *global_count, void *stream);

/* Bench instrumentation: with profiling on, acwm_scan_device brackets the scan kernel and
 * the finalize kernels with CUDA events on the caller's stream; acwm_profiled_seconds
 * waits for the last profiled scan and returns the two durations. */
int acwm_set_profiling(acwm_matcher *mt, int on);
int acwm_profiled_seconds(acwm_matcher *mt, double *scan_s, double *finalize_s);

/* Kernel timeline instrumentation: d_trace = device buffer of acwm_trace_words_per_cta() uint64 per CTA
 * of the scan grid (<= 256 CTAs), or NULL = off.  Every CTA of the following scans stores %globaltimer
 * stamps of its phases there (entry, prologue, tables ready, scan done, arrival, exit; per warp: first tile
 * ready, scan loop done, tiles scanned); scripts/trace.py reads them back. */
int acwm_set_trace(acwm_matcher *mt, unsigned long long *d_trace);
uint32_t acwm_trace_words_per_cta(void);

/* Kernels this matcher has launched so far (scan + finalize), for bench reports. */
unsigned long long acwm_launch_count(const acwm_matcher *mt);

int acwm_get_info(const acwm_matcher *mt, acwm_info *info);
void acwm_free(acwm_matcher *mt);

/* Text of the last error on this thread ("" if none). */
const char *acwm_last_error(void);

/* Shard geometry of the multi-GPU layer == the MPI rank geometry of main.c:467-477:
 * chunk = ceil(n/world); shard `rank` is text[start, start+len) with
 * start = rank*chunk, len = min((rank+1)*chunk + halo, n) - start, halo = m_max-1.
 * A plain scan of the shard reports local ends e_local in [m-1, len); the global
 * end is start + e_local, so every match is found by exactly one shard. */
void acwm_shard_bounds(uint64_t n, uint32_t world, uint32_t rank, uint32_t halo, uint64_t *start, uint64_t *len);

/* CUDA devices this process sees (0 if none / no driver). */
int acwm_device_count(void);

/* The whole multi-rank flow of main.c in one call and one process: MPI_Scatterv of the text with an (m-1)-byte halo
 * (main.c:464-488), the per-rank search (main.c:630-647) and MPI_Reduce(SUM) of the counts (main.c:656), plus the
 * gather of the positions the reference never had.  mts[r] is the matcher of shard r (build one per shard from the
 * same pattern set; upload each to the device it shall run on -- a matcher not uploaded yet goes to device
 * r mod acwm_device_count(); several shards may share a device).  Shard r = acwm_shard_bounds(n, world, r, m_max-1);
 * one host thread per shard runs acwm_search_host on its slice of `text` (pinned or pageable), so the copies and the
 * scans of all devices overlap.  *count = sum over the shards (== the unsharded count), positions = global match
 * ends, ascending (shards are in text order and each is sorted); shard_counts (may be NULL) receives the
 * world per-shard counts.  Errors as acwm_search_host; ACWM_ERR_OVERFLOW if cap positions do not hold them all. */
int acwm_search_host_sharded(acwm_matcher *const *mts, uint32_t world, const uint8_t *text, uint64_t n, uint64_t *count,
		uint64_t *positions, uint64_t cap, uint64_t *n_written, uint64_t *shard_counts);

/* The same multi-rank flow for DEVICE-RESIDENT text, in one process, without torch, NCCL or MPI: what replaces
 * MPI_Scatterv + MPI_Reduce (main.c:488,656) when the shards already sit in HBM.  mts[r] (uploaded, one per shard;
 * several may share a device) gets a zero-initialised mailbox in its device's memory (cudaMalloc), peer access is
 * enabled between the devices involved (cudaDeviceEnablePeerAccess) and the matchers are wired with acwm_set_peers:
 * from then on every acwm_scan_device_sharded exchanges the per-shard counts INSIDE the scan kernels over NVLink.
 * ACWM_ERR_UNSUPPORTED if two of the devices cannot access each other's memory. */
int acwm_peers_create(acwm_matcher *const *mts, uint32_t world);
int acwm_peers_destroy(acwm_matcher *const *mts, uint32_t world);
/* One scan of every shard: d_shards[r] (on mts[r]'s device) holds shard r of the text INCLUDING its halo, shard_lens[r]
 * bytes (acwm_shard_bounds); shard r > 0 reports ends from m_max-1 on, so every match is reported exactly once.
 * Asynchronous; matchers of one device run on one stream, in rank order.  Like a collective, every call scans all
 * shards.  acwm_fetch_sharded waits for the last one and returns the count summed over the shards as the kernels
 * exchanged it (every shard's kernel holds the same sum) and the per-shard counts; positions stay per shard
 * (acwm_fetch on mts[r], shard-local, + the shard start = global). */
int acwm_scan_device_sharded(acwm_matcher *const *mts, uint32_t world, const uint8_t *const *d_shards,
		const uint64_t *shard_lens, int want_positions);
int acwm_fetch_sharded(acwm_matcher *const *mts, uint32_t world, uint64_t *global_count, uint64_t *shard_counts);
/* Device buffers for callers that are plain C (no CUDA runtime of their own): a copy of n host bytes in the memory
 * of `device`, and its release. */
int acwm_text_to_device(int device, const uint8_t *text, uint64_t n, uint8_t **d_text);
void acwm_device_free(int device, void *d_ptr);

/* Raw views of the compiled tables (tests and diagnostics).  `which` is one of the
 * ACWM_BLOB_* ids; returns ACWM_ERR_INVALID if this matcher has no such table. */
enum {
	ACWM_BLOB_FRONT = 0,      /* AC: k-stride DFA (uint16 entries); WM: stage-1 block bitmap */
	ACWM_BLOB_FILTER2 = 1,    /* stage-2 suffix bitmap */
	ACWM_BLOB_BUCKET_START = 2,
	ACWM_BLOB_ENTRIES = 3,    /* acwm_ventry[] */
	ACWM_BLOB_PATTERNS = 4,   /* distinct pattern bytes, back to back */
	ACWM_BLOB_PARAMS = 5,     /* acwm_scan_params */
	ACWM_BLOB_SYMCLASS = 6,   /* bytes path AC: 256-entry symbol -> class map */
	ACWM_BLOB_RMASK = 7,      /* WM, stride > 1: offset masks of the candidate blocks (uint8, uint16 for stride 16) */
	ACWM_BLOB_VDFA = 8        /* filtered AC: full-depth one-symbol DFA, uint32 entries (next row << 1 | final) */
};
int acwm_table_blob(const acwm_matcher *mt, int which, const void **ptr, uint64_t *bytes);

/* ------------- data files (host only): the driver's side of the path, main.c:31-123,453 -------------
 * The reference's tables are indexed by the text bytes themselves (ac/ac.c:136,209; wu/wu.c:63-67), so
 * a corpus reaches the scan as symbol codes in [0, alphabet).  These replace the helpers the reference
 * calls but does not ship (load_files, create_multiple_pattern_with_hits: main.c:49,453). */

/* Raw byte -> symbol code for the corpora of main.c:38-110; 0xff = not a symbol (dropped by the loader).
 * 2 / 8: digits '0'..; 4: ACGT (either case, U = T); 20: the amino-acid letters ACDEFGHIKLMNPQRSTVWY;
 * 128: 7-bit ASCII; 256: identity. */
int acwm_symbol_map(uint32_t alphabet, uint8_t map[256]);
/* Encode n_raw corpus bytes into symbol codes (out may alias raw; at most n_raw codes).  Bytes that are
 * already all < alphabet pass through unchanged; FASTA '>' lines and non-symbols are dropped. */
int acwm_encode_symbols(const uint8_t *raw, uint64_t n_raw, uint32_t alphabet, uint8_t *out, uint64_t *n_out);
/* Load a corpus file as symbol codes (malloc'ed; release with acwm_free_text); max_symbols = 0: all of it,
 * else the first max_symbols (the reference's -n). */
int acwm_load_text(const char *path, uint32_t alphabet, uint64_t max_symbols, uint8_t **text, uint64_t *n);
void acwm_free_text(uint8_t *text);
/* p patterns of m symbols into patterns[p*m] (the flat pattern2 layout, main.c:455-461): hit_percent of
 * them are windows of the text at seeded offsets (they occur), the rest uniform symbols. */
int acwm_patterns_with_hits(const uint8_t *text, uint64_t n, uint32_t m, uint32_t p, uint32_t alphabet, uint64_t seed,
		uint32_t hit_percent, uint8_t *patterns);
/* select_data_file (main.c:31-123): the text size n selects the corpus under data_root (NULL =
 * "../data-cuda-multi"), the alphabet must fit it; pattern file = pattern/<n>/<m>/<alphabet>/pattern. */
int acwm_select_data_file(uint32_t m, uint64_t n, uint32_t alphabet, const char *data_root, char *pattern_path,
		char *text_path, size_t path_cap);

/* One pattern in the verification buckets. */
typedef struct acwm_ventry {
	uint32_t key;     /* packed last-B2 symbols of the pattern (what the text window must equal) */
	uint32_t len;     /* pattern length; bit 31 set = key equality alone proves the match */
	uint64_t offset;  /* byte offset of the pattern in the ACWM_BLOB_PATTERNS blob */
} acwm_ventry;

/* Scalar parameters the kernels run with (also what the test emulator reads). */
typedef struct acwm_scan_params {
	uint32_t algo, packed2bit, alphabet, m_min, m_max;
	uint32_t stride;        /* K (AC) or s (WM) */
	uint32_t depth;         /* D (AC) or B (WM) */
	uint32_t exact_front;
	uint32_t n_rows;        /* AC */
	uint32_t f1_sh1, f1_mult, f1_sh2, f1_words;   /* WM stage 1: idx = ((v >> sh1) * mult) >> sh2 */
	uint32_t b2;            /* stage-2 / bucket key length in symbols */
	uint32_t f2_mult, f2_sh, f2_words;            /* idx2 = (key * mult) >> sh */
	uint32_t hb_mult, hb_sh, n_buckets;           /* bucket = (key * mult) >> sh */
	uint32_t n_entries;
	uint32_t n_classes;     /* bytes path AC */
	uint32_t r_mult, r_sh, r_entries, r_entry_bytes; /* WM offset masks: ridx = (block * mult) >> sh; 0 entries = none */
	uint32_t r_in_smem, f2_in_smem; /* 1: the table is staged in shared memory, 0: read from global memory (L2-resident) */
	uint32_t front_kind;    /* AC: 0 = the automaton walks every symbol, 1 = WM-style block filter in front, the
	                         * automaton decides candidate windows only */
	uint32_t verify_kind;   /* what decides a candidate window: 0 = hash buckets + compare, 1 = walk of the verify DFA */
	uint32_t v_rows;        /* rows of the verify DFA (ACWM_BLOB_VDFA) */
	uint32_t ilp;           /* AC, exact K = 3 automaton in shared memory: 2 = every lane walks its chunk as two independent chains */
	uint32_t f1_k;          /* WM stage 1, hashed bitmaps: bits per entry (1, or 2 = blocked Bloom filter: both bits in the word
	                         * idx >> 5, the second at bit ((v >> sh1) * mult >> (sh2 - 5)) & 31) */
} acwm_scan_params;

/* ------------------- reference-shaped shims (smatcher.h) ------------------- */

struct ac_state; /* smatcher.h:41-47; never dereferenced by callers of this path */
struct ac_table { /* smatcher.h:49-53 -- same leading layout, so main.c compiles unchanged */
	unsigned int idcounter;      /* number of states of the full automaton */
	unsigned int patterncounter; /* number of distinct patterns */
	struct ac_state *zerostate;  /* here: opaque pointer to the acwm_matcher */
};

/* Fills the caller's flat tables exactly as the reference does (state ids in
 * creation order, -1 = no goto edge, root row 0: ac/ac.c:61-62,114,162,186) and
 * compiles + uploads the device tables. */
struct ac_table *preproc_ac(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
unsigned search_ac(unsigned char *text, int n, struct ac_table *table);
void free_ac(struct ac_table *table, int alphabet);

extern unsigned short m_nBitsInShift; /* smatcher.h:71 */
extern unsigned int shiftsize;        /* smatcher.h:73 */
void wu_determine_shiftsize(int alphabet);
/* Fill SHIFT / PREFIX_* exactly as the reference does (wu/wu.c:109-149) and compile
 * + upload the device tables (keyed by the SHIFT pointer for the search_* shims). */
void preproc_wu(unsigned char **pattern, int m, int p_size, int alphabet, int B, int *SHIFT, int *PREFIX_value,
		int *PREFIX_index, int *PREFIX_size);
void preproc_wu2(unsigned char *pattern, int m, int p_size, int alphabet, int B, int *SHIFT, int *PREFIX_value,
		int *PREFIX_index, int *PREFIX_size);
unsigned int search_wu(unsigned char **pattern, int m, int p_size, unsigned char *text, int n, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size);
unsigned int search_wu2(unsigned char *pattern, int m, int p_size, unsigned char *text, int n, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size);

/* GPU wrappers.  The reference's cuda_acN print "Kernel N matches \t%i\t time \t%f"
 * and return void (cuda/cuda_ac.cu:675); these print the same line.  The pattern
 * set is recovered from the flat goto table (every terminal state spells one
 * pattern).  cuda_wmN return the count and write the kernel seconds. */
void cuda_ac1(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
void cuda_ac2(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
void cuda_ac3(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
void cuda_ac4(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
void cuda_ac5(int m, unsigned char *text, int n, int p_size, int alphabet, int *state_transition,
		unsigned int *state_supply, unsigned int *state_final);
int cuda_wm1(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime);
int cuda_wm2(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime);
int cuda_wm3(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime);
int cuda_wm4(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime);
int cuda_wm5(unsigned char *pattern, int m, unsigned char *text, int n, int p_size, int alphabet, int B, int *SHIFT,
		int *PREFIX_value, int *PREFIX_index, int *PREFIX_size, double *gpuTime);

/* Count of the last cuda_acN call (the reference prints it and returns void). */
unsigned long long acwm_shim_last_count(void);

/* ------------------- sibling algorithms behind the same matcher (smatcher.h:93-99,109-110) -------------------
 * Set-Horspool (sh/sh.c:151-178), Set Backward Oracle Matching (sbom/sbom.c:152-196) and Shift-Or with q-grams
 * (sog/sog8.c:97-115) return what search_ac / search_wu return: the number of end positions whose m-symbol window is
 * one of the distinct patterns.  Same signatures as the reference; the searches run on the GPU through the matcher
 * compiled by the preproc call.  The caller's flat tables are filled as the reference fills them (state ids in
 * creation order over the REVERSED patterns, oracle edges and F(q) lists for sbom, 3-gram masks / sorted hashes /
 * two-level bitmap for sog8 -- see csrc/tables.cpp for the two documented deviations of sog8). */
struct sbom_state; /* smatcher.h:57-63; never dereferenced by callers of this path */
struct sbom_table { /* smatcher.h:65-69 */
	unsigned int idcounter;
	unsigned int patterncounter;
	struct sbom_state *zerostate; /* here: opaque pointer to the acwm_matcher */
};
struct ac_table *preproc_sh(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_final);
unsigned search_sh(int m, unsigned char *text, int n, struct ac_table *table, int *bmBc);
void free_sh(struct ac_table *table, int alphabet);
struct sbom_table *preproc_sbom(unsigned char **pattern, int m, int p_size, int alphabet, int *state_transition,
		unsigned int *state_final_multi);
unsigned search_sbom(unsigned char **pattern, int m, unsigned char *text, int n, struct sbom_table *table);
void free_sbom(struct sbom_table *table, int m);
/* T8: 2^24 bytes, scanner_hs / scanner_index: p_size entries, scanner_hs2: 8192 bytes (main.c:546 and the commented
 * allocation above it); m must be 8.  The matcher is remembered by T8; acwm_shim_forget(T8) releases it. */
void preproc_sog8(uint8_t *T8, uint32_t *scanner_hs, int *scanner_index, uint8_t *scanner_hs2, unsigned char **pattern, int m,
		unsigned char *text, int n, int p_size, int B);
unsigned int search_sog8(uint8_t *T8, uint32_t *scanner_hs, int *scanner_index, uint8_t *scanner_hs2, unsigned char **pattern,
		int m, unsigned char *text, int n, int p_size, int B);
/* Releases the matcher a table-keyed entry point (preproc_wu*, cuda_acN, cuda_wmN, preproc_sog8) remembers for `table`. */
void acwm_shim_forget(const void *table);

#ifdef __cplusplus
}
#endif
#endif /* ACWM_H */
