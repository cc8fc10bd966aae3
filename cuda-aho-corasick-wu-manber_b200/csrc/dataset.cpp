// Data-file side of the path (host only, no CUDA): what the reference's driver does before it
// calls preprocess / search -- pick the corpus for a run (select_data_file, main.c:31-123), load
// it as symbol codes (the missing load_files helper, main.c:453) and draw a pattern set "with
// hits" from it (the missing create_multiple_pattern_with_hits helper, main.c:49).
//
// The reference indexes its tables with the text bytes themselves (ac/ac.c:136,209;
// wu/wu.c:63-67), so a corpus has to reach the scan as codes in [0, alphabet).  Raw corpora
// (FASTA nucleotides / amino acids, English text) are mapped here; a file whose bytes are
// already all < alphabet is taken as it is.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <new>
#include <vector>

#include "../../include/acwm.h"

namespace acwm {
int set_error(int code, const std::string &msg);
}
using acwm::set_error;

namespace {

constexpr uint8_t kDrop = 0xff; // byte is not a symbol of this alphabet: skipped by the loader

// xorshift64* (the generator of examples/smatcher_main.c): the C driver and the tests draw the same sets
struct Rng {
	uint64_t s;
	explicit Rng(uint64_t seed) : s(88172645463325252ull ^ (seed * 0x9E3779B97F4A7C15ull)) {
		if (!s)
			s = 1;
	}
	uint64_t next() {
		s ^= s >> 12;
		s ^= s << 25;
		s ^= s >> 27;
		return s * 2685821657736338717ull;
	}
};

} // namespace

extern "C" {

int acwm_symbol_map(uint32_t alphabet, uint8_t map[256]) {
	if (!map)
		return set_error(ACWM_ERR_INVALID, "map == NULL");
	memset(map, kDrop, 256);
	switch (alphabet) {
	case 2: // random binary text (main.c:39-45, text2)
	case 8: // random octal text (text8)
		for (uint32_t c = 0; c < alphabet; c++)
			map['0' + c] = (uint8_t) c;
		break;
	case 4: { // nucleotides (E.coli, A_thaliana.fna: main.c:66,99)
		const char *up = "ACGT", *lo = "acgt";
		for (int c = 0; c < 4; c++)
			map[(uint8_t) up[c]] = map[(uint8_t) lo[c]] = (uint8_t) c;
		map['U'] = map['u'] = 3;
		break;
	}
	case 20: { // amino acids (swiss-prot, A_thaliana.faa: main.c:77,88)
		const char *aa = "ACDEFGHIKLMNPQRSTVWY";
		for (int c = 0; c < 20; c++) {
			map[(uint8_t) aa[c]] = (uint8_t) c;
			map[(uint8_t) (aa[c] + 32)] = (uint8_t) c;
		}
		break;
	}
	case 128: // English text (world192.txt: main.c:54): 7-bit ASCII as it is
		for (int c = 0; c < 128; c++)
			map[c] = (uint8_t) c;
		break;
	case 256:
		for (int c = 0; c < 256; c++)
			map[c] = (uint8_t) c;
		return ACWM_OK; // every byte is a symbol
	default:
		return set_error(ACWM_ERR_UNSUPPORTED, "no corpus mapping for this alphabet (2, 4, 8, 20, 128, 256)");
	}
	return ACWM_OK;
}

int acwm_encode_symbols(const uint8_t *raw, uint64_t n_raw, uint32_t alphabet, uint8_t *out, uint64_t *n_out) {
	if ((!raw && n_raw) || !out || !n_out)
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	bool coded = true; // already symbol codes?
	for (uint64_t i = 0; i < n_raw && coded; i++)
		coded = raw[i] < alphabet;
	if (coded || alphabet == 256) {
		if (out != raw)
			memmove(out, raw, n_raw);
		*n_out = n_raw;
		return ACWM_OK;
	}
	uint8_t map[256];
	const int rc = acwm_symbol_map(alphabet, map);
	if (rc != ACWM_OK)
		return rc;
	const bool fasta = alphabet == 4 || alphabet == 20;
	uint64_t w = 0;
	bool bol = true, header = false;
	for (uint64_t i = 0; i < n_raw; i++) {
		const uint8_t b = raw[i];
		if (fasta) { // '>' description lines are not sequence
			if (bol && b == '>')
				header = true;
			bol = b == '\n';
			if (header) {
				if (b == '\n')
					header = false;
				continue;
			}
		}
		const uint8_t c = map[b];
		if (c != kDrop)
			out[w++] = c;
	}
	*n_out = w;
	return ACWM_OK;
}

int acwm_load_text(const char *path, uint32_t alphabet, uint64_t max_symbols, uint8_t **text, uint64_t *n) {
	if (!path || !text || !n)
		return set_error(ACWM_ERR_INVALID, "NULL argument");
	*text = nullptr;
	*n = 0;
	FILE *f = fopen(path, "rb");
	if (!f)
		return set_error(ACWM_ERR_INVALID, std::string("cannot open ") + path);
	std::vector<uint8_t> raw;
	uint8_t buf[1 << 16];
	size_t got;
	try { // no exception leaves the C ABI
		while ((got = fread(buf, 1, sizeof(buf), f)) > 0)
			raw.insert(raw.end(), buf, buf + got);
	} catch (const std::bad_alloc &) {
		fclose(f);
		return set_error(ACWM_ERR_NOMEM, "host allocation failed while reading the corpus");
	}
	fclose(f);
	uint8_t *out = (uint8_t *) malloc(raw.size() + 1);
	if (!out)
		return set_error(ACWM_ERR_NOMEM, "host allocation failed");
	uint64_t w = 0;
	const int rc = acwm_encode_symbols(raw.data(), raw.size(), alphabet, out, &w);
	if (rc != ACWM_OK) {
		free(out);
		return rc;
	}
	if (max_symbols && w > max_symbols)
		w = max_symbols; // the reference reads the first n symbols of the corpus (n selects the corpus, main.c:38)
	*text = out;
	*n = w;
	return ACWM_OK;
}

void acwm_free_text(uint8_t *text) { free(text); }

int acwm_patterns_with_hits(const uint8_t *text, uint64_t n, uint32_t m, uint32_t p, uint32_t alphabet, uint64_t seed,
		uint32_t hit_percent, uint8_t *patterns) {
	if (!patterns || (!text && n) || m == 0 || alphabet == 0 || alphabet > 256 || hit_percent > 100)
		return set_error(ACWM_ERR_INVALID, "bad argument");
	Rng rng(seed);
	// hits are spread evenly over the set: pattern j is a window of the text when the running
	// share of hits falls behind hit_percent (50 % -> every other pattern, as smatcher_main does)
	uint64_t hits = 0;
	for (uint32_t j = 0; j < p; j++) {
		uint8_t *dst = patterns + (uint64_t) j * m;
		const bool hit = n >= m && hits * 100 < (uint64_t) (j + 1) * hit_percent;
		if (hit) {
			const uint64_t at = rng.next() % (n - m + 1);
			memcpy(dst, text + at, m);
			hits++;
		} else
			for (uint32_t i = 0; i < m; i++)
				dst[i] = (uint8_t) (rng.next() % alphabet);
	}
	return ACWM_OK;
}

int acwm_select_data_file(uint32_t m, uint64_t n, uint32_t alphabet, const char *data_root, char *pattern_path,
		char *text_path, size_t path_cap) {
	if (!pattern_path || !text_path || path_cap < 16)
		return set_error(ACWM_ERR_INVALID, "bad argument");
	const std::string root = data_root && *data_root ? data_root : "../data-cuda-multi"; // main.c:35
	struct Corpus {
		uint64_t n;
		uint32_t alphabet, alphabet2;
		const char *file, *file2, *what;
	};
	// main.c:38-110: the text size selects the corpus, the alphabet must fit it
	static const Corpus table[] = {
			{3999744, 2, 8, "text2", "text8", "random texts, you must use an alphabet size of 2 or 8"},
			{1903104, 128, 0, "world192.txt", nullptr, "english text, you must use an alphabet size of 128"},
			{4628736, 4, 0, "E.coli2", nullptr, "DNA sequences, you must use an alphabet size of 4"},
			{177649920, 20, 0, "swiss-prot", nullptr, "swiss-prot, you must use an alphabet size of 20"},
			{10821888, 20, 0, "A_thaliana.faa", nullptr, "A_thaliana.faa, you must use an alphabet size of 20"},
			{116234496, 4, 0, "A_thaliana.fna", nullptr, "A_thaliana.fna, you must use an alphabet size of 4"},
	};
	std::string text, pat = root + "/pattern/" + std::to_string(n) + "/" + std::to_string(m) + "/" +
			std::to_string(alphabet) + "/pattern";
	if (n == 100) { // the debug pair (main.c:111-118)
		if (alphabet != 2)
			return set_error(ACWM_ERR_INVALID, "The debug text uses a binary alphabet");
		text = root + "/text/debug";
		pat = root + "/pattern/debug";
	} else {
		const Corpus *c = nullptr;
		for (const Corpus &e : table)
			if (e.n == n)
				c = &e;
		if (!c)
			return set_error(ACWM_ERR_INVALID, "Please select an appropriate text size");
		if (alphabet != c->alphabet && !(c->alphabet2 && alphabet == c->alphabet2))
			return set_error(ACWM_ERR_INVALID, std::string("For ") + c->what);
		text = root + "/text/" + ((c->alphabet2 && alphabet == c->alphabet2) ? c->file2 : c->file);
	}
	if (text.size() + 1 > path_cap || pat.size() + 1 > path_cap)
		return set_error(ACWM_ERR_INVALID, "path buffer too small");
	memcpy(text_path, text.c_str(), text.size() + 1);
	memcpy(pattern_path, pat.c_str(), pat.size() + 1);
	return ACWM_OK;
}

} // extern "C"
