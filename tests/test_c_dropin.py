"""The C integration example (examples/smatcher_main.c, the driver INTEGRATION.md shows):
plain C, the reference's own call sequence (main.c:125-157, 268-298, 582-648), linked
against libacwm_b200.so.  CPU: it compiles and links with gcc against include/acwm.h.
GPU: it runs on files written here and prints the oracle's counts in the reference's
own report format."""
import os
import re
import subprocess

import numpy as np
import pytest

from cases import RANDOM_CASES, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "smatcher_main")


def _build(acwm):
    pkg = os.path.dirname(acwm.LIB_PATH)
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-I",
                    os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "smatcher_main.c"), "-L", pkg,
                    "-lacwm_b200", "-Wl,-rpath," + pkg, "-o", EXE], check=True)


def test_c_driver_compiles_and_links_as_plain_c(acwm):
    _build(acwm)
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True, check=True).stdout
    assert "libacwm_b200.so" in out and "not found" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_c_driver_prints_oracle_counts(acwm, oracle, tmp_path, idx):
    _build(acwm)
    case = RANDOM_CASES[idx]
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    want = oracle.set_search(pats, text)["count"]
    tf, pf = tmp_path / "text.bin", tmp_path / "pattern.bin"
    text.tofile(tf)
    np.ascontiguousarray(pats).tofile(pf)
    out = subprocess.run([EXE, "ac" if algo == acwm.AC else "wm", "-m", str(m), "-n", str(text.size), "-p_size",
                          str(p), "-alphabet", str(alphabet), "-text", str(tf), "-pattern", str(pf)],
                         capture_output=True, text=True, check=True, timeout=300).stdout
    first = "search_ac matches" if algo == acwm.AC else "search_wm2 matches"
    assert int(re.search(first + r" \t(\d+)\t", out).group(1)) == want
    assert int(re.search(r"Kernel 5 matches \t(\d+)\t", out).group(1)) == want
    assert int(re.search(r"Total results: (\d+)\.", out).group(1)) == want


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sh", "sbom"])
def test_c_driver_runs_the_sibling_algorithms(acwm, oracle, tmp_path, name):
    """argv[1] = sh / sbom (the dispatch main.c:519-531 keeps in comments): multish / multisbom through the shims."""
    _build(acwm)
    case = RANDOM_CASES[0]
    _, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    want = oracle.set_search(pats, text)["count"]
    tf, pf = tmp_path / "text.bin", tmp_path / "pattern.bin"
    text.tofile(tf)
    np.ascontiguousarray(pats).tofile(pf)
    out = subprocess.run([EXE, name, "-m", str(m), "-n", str(text.size), "-p_size", str(p), "-alphabet", str(alphabet),
                          "-text", str(tf), "-pattern", str(pf)], capture_output=True, text=True, check=True, timeout=300).stdout
    assert int(re.search(rf"search_{name} matches \t(\d+)\t", out).group(1)) == want
    assert int(re.search(r"Total results: (\d+)\.", out).group(1)) == want


@pytest.mark.gpu
def test_c_driver_selects_and_loads_a_corpus(acwm, oracle, tmp_path):
    """-data DIR -c: the corpus is chosen like the reference's select_data_file (n = 4628736 -> text/E.coli2), loaded
    from FASTA through the library's symbol map, the pattern set drawn with hits -- and the printed counts are the
    oracle's on the same symbols and patterns."""
    _build(acwm)
    n, m, p = 4628736, 8, 100
    rng = np.random.default_rng(12)
    seq = rng.integers(0, 4, n + 1000, dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", np.uint8)[seq]
    (tmp_path / "text").mkdir()
    with open(tmp_path / "text" / "E.coli2", "wb") as f:
        f.write(b">synthetic E. coli\n")
        for i in range(0, letters.size, 70):
            f.write(letters[i:i + 70].tobytes() + b"\n")
    out = subprocess.run([EXE, "wm", "-m", str(m), "-n", str(n), "-p_size", str(p), "-alphabet", "4", "-data", str(tmp_path),
                          "-c", "-seed", "5"], capture_output=True, text=True, check=True, timeout=300).stdout
    text = acwm.load_text(str(tmp_path / "text" / "E.coli2"), 4, n)
    assert np.array_equal(text, seq[:n])
    pats = acwm.patterns_with_hits(text, m, p, 4, seed=5, hit_percent=50)
    want = oracle.set_search(pats, text)["count"]
    assert int(re.search(r"search_wm2 matches \t(\d+)\t", out).group(1)) == want
    assert int(re.search(r"Total results: (\d+)\.", out).group(1)) == want


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_c_driver_multi_gpu_mode(acwm, oracle, tmp_path, idx):
    """-gpus G: the MPI flow of main.c:464-656 as shards of one process; per-rank counts are those of the
    reference's ranks and they sum to the unsharded count."""
    _build(acwm)
    case = RANDOM_CASES[idx]
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    ref = oracle.set_search(pats, text)
    tf, pf = tmp_path / "text.bin", tmp_path / "pattern.bin"
    text.tofile(tf)
    np.ascontiguousarray(pats).tofile(pf)
    G = 3
    out = subprocess.run([EXE, "ac" if algo == acwm.AC else "wm", "-m", str(m), "-n", str(text.size), "-p_size",
                          str(p), "-alphabet", str(alphabet), "-text", str(tf), "-pattern", str(pf), "-gpus", str(G)],
                         capture_output=True, text=True, check=True, timeout=300).stdout
    totals = [int(x) for x in re.findall(r"Total results: (\d+)\.", out)]
    assert totals == [ref["count"], ref["count"]], out  # the reference-shaped flow, then the sharded one
    ranks = re.findall(r"rank (\d+) device (\d+) text \[(\d+), (\d+)\) matches \t(\d+)\t", out)
    assert len(ranks) == G
    for r, (rank, dev, lo, hi, cnt) in enumerate(ranks):
        start, length = acwm.shard_bounds(text.size, G, r, m - 1)
        assert (int(rank), int(lo), int(hi)) == (r, start, start + length)
        first = start + m - 1
        inside = (ref["positions"] >= first) & (ref["positions"] < start + length)
        assert int(cnt) == int(inside.sum())
