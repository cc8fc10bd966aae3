"""The data-file side of the path (csrc/dataset.cpp; host only): symbol maps, corpus loader, pattern sets
"with hits" and the reference's corpus selection (select_data_file, main.c:31-123)."""
import numpy as np
import pytest

import acwm_pkg

acwm = acwm_pkg.load()


def test_symbol_maps():
    dna = acwm.symbol_map(4)
    assert [dna[ord(c)] for c in "ACGTacgtU"] == [0, 1, 2, 3, 0, 1, 2, 3, 3]
    assert dna[ord("N")] == 0xFF and dna[ord("\n")] == 0xFF
    aa = acwm.symbol_map(20)
    assert sorted(aa[ord(c)] for c in "ACDEFGHIKLMNPQRSTVWY") == list(range(20))
    assert aa[ord("B")] == 0xFF and aa[ord("a")] == 0
    assert list(acwm.symbol_map(2)[[ord("0"), ord("1"), ord("2")]]) == [0, 1, 0xFF]
    assert list(acwm.symbol_map(8)[[ord("0"), ord("7"), ord("8")]]) == [0, 7, 0xFF]
    assert np.array_equal(acwm.symbol_map(128)[:128], np.arange(128, dtype=np.uint8))
    assert np.array_equal(acwm.symbol_map(256), np.arange(256, dtype=np.uint8))
    with pytest.raises(acwm.AcwmError):
        acwm.symbol_map(5)


def test_encode_fasta_and_coded_passthrough():
    fasta = b">seq1 some description ACGT\nACGTN\nacgt\n>seq2\nGGxCC\n"
    assert acwm.encode_symbols(fasta, 4).tolist() == [0, 1, 2, 3, 0, 1, 2, 3, 2, 2, 1, 1]
    prot = b">sp|P1\nMKV\nB*LA\n"
    aa = "ACDEFGHIKLMNPQRSTVWY"
    assert acwm.encode_symbols(prot, 20).tolist() == [aa.index(c) for c in "MKVLA"]
    coded = np.array([0, 3, 2, 1, 1, 0], np.uint8)  # already symbol codes: untouched
    assert np.array_equal(acwm.encode_symbols(coded, 4), coded)
    text = b"Hello, world\n"
    assert acwm.encode_symbols(text, 128).tobytes() == text
    raw = bytes(range(256))
    assert acwm.encode_symbols(raw, 256).tobytes() == raw


def test_load_text_and_patterns_with_hits(tmp_path):
    rng = np.random.default_rng(3)
    seq = "".join(rng.choice(list("ACGT"), 5000))
    path = tmp_path / "E.coli2"
    path.write_text(">x\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n")
    text = acwm.load_text(str(path), 4)
    assert text.size == 5000 and text.max() < 4 and "".join("ACGT"[c] for c in text[:50]) == seq[:50]
    assert acwm.load_text(str(path), 4, max_symbols=1234).size == 1234
    with pytest.raises(acwm.AcwmError):
        acwm.load_text(str(tmp_path / "missing"), 4)
    pats = acwm.patterns_with_hits(text, 8, 100, 4, seed=9, hit_percent=50)
    assert pats.shape == (100, 8) and pats.max() < 4
    windows = {text[i:i + 8].tobytes() for i in range(text.size - 7)}
    hits = sum(p.tobytes() in windows for p in pats)
    assert 50 <= hits <= 60  # the 50 planted ones (+ a random 8-mer that happens to occur)
    assert all(pats[j].tobytes() in windows for j in range(0, 100, 2))
    again = acwm.patterns_with_hits(text, 8, 100, 4, seed=9, hit_percent=50)
    assert np.array_equal(pats, again)
    assert acwm.patterns_with_hits(text, 8, 100, 4, seed=10).tobytes() != pats.tobytes()
    none = acwm.patterns_with_hits(text, 8, 20, 4, seed=1, hit_percent=0)
    assert none.shape == (20, 8)


def test_select_data_file_follows_the_reference_table():
    # main.c:38-110: the text size selects the corpus; main.c:35: the pattern path
    pp, tp = acwm.select_data_file(8, 4628736, 4)
    assert tp == "../data-cuda-multi/text/E.coli2" and pp == "../data-cuda-multi/pattern/4628736/8/4/pattern"
    assert acwm.select_data_file(8, 3999744, 2, "/d")[1] == "/d/text/text2"
    assert acwm.select_data_file(8, 3999744, 8, "/d")[1] == "/d/text/text8"
    assert acwm.select_data_file(16, 1903104, 128, "/d") == ("/d/pattern/1903104/16/128/pattern", "/d/text/world192.txt")
    assert acwm.select_data_file(8, 177649920, 20, "/d")[1] == "/d/text/swiss-prot"
    assert acwm.select_data_file(8, 10821888, 20, "/d")[1] == "/d/text/A_thaliana.faa"
    assert acwm.select_data_file(8, 116234496, 4, "/d")[1] == "/d/text/A_thaliana.fna"
    assert acwm.select_data_file(8, 100, 2, "/d") == ("/d/pattern/debug", "/d/text/debug")
    for bad in ((8, 3999744, 4), (8, 4628736, 20), (8, 100, 4), (8, 12345, 4)):  # the reference's fail() cases
        with pytest.raises(acwm.AcwmError):
            acwm.select_data_file(*bad)


def test_host_packer_matches_numpy():
    """acwm_pack_text_2bit (what acwm_search_host runs in front of the H2D copy of DNA texts) against numpy, on sizes
    around every boundary of its loops (8 / 32 symbols, 256 Ki-symbol work items), repeated: the pool is reused."""
    rng = np.random.default_rng(17)
    for n in (0, 1, 3, 4, 7, 8, 31, 32, 33, 1000, 262144, 262145, 3 * 262144 + 5, (5 << 20) + 3):
        for rep in range(2):
            t = rng.integers(0, 4, n, dtype=np.uint8)
            packed, bad = acwm.pack_text_2bit(t)
            pad = np.zeros((-n) % 4, np.uint8)
            q = np.concatenate([t, pad]).reshape(-1, 4).astype(np.uint32)
            want = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
            assert not bad and np.array_equal(packed, want), n
    t = rng.integers(0, 4, 700_000, dtype=np.uint8)
    for where in (0, 5, 262143, 699_999):
        u = t.copy()
        u[where] = 4
        assert acwm.pack_text_2bit(u)[1], where
