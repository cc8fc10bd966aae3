"""CPU-side checks of the product: the C-ABI library loads and exports what
include/acwm.h declares, the table compiler's output (read back through
acwm_table_blob and run by the numpy emulator tests/emu.py) reproduces the oracle's
matches, the reference-shaped preproc_* shims fill the caller's flat tables exactly as
the reference does, and the error paths answer with status codes.  No CUDA call."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import RANDOM_CASES, edge_cases, make_case
from golden_util import load_golden

GOLD = load_golden()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(acwm):
    hdr = open(os.path.join(ROOT, "include", "acwm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b((?:acwm_[a-z_0-9]+|preproc_\w+|search_\w+|free_ac|wu_determine_shiftsize|"
                           r"cuda_ac[1-5]|cuda_wm[1-5]))\s*\(", hdr))
    names -= {"acwm_matcher"}
    assert len(names) >= 30
    L = acwm.lib()
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/acwm.h but not exported"
    for var in ("m_nBitsInShift", "shiftsize"):
        C.c_uint.in_dll(L, var)


def test_struct_sizes(acwm):
    assert C.sizeof(acwm.Options) == 44 and C.sizeof(acwm.Info) == 80 and C.sizeof(acwm.ScanParams) == 132
    assert acwm.VENTRY_DTYPE.itemsize == 16


@pytest.mark.parametrize("case", RANDOM_CASES, ids=lambda c: c[0])
def test_tables_reproduce_oracle(acwm, oracle, case):
    from emu import Emulator
    name, algo, alphabet, p, m, n, opts = case
    pats, text = make_case(case)
    text = text[: min(n, 120_000)]
    mt = acwm.Matcher(algo, pats, alphabet, **opts)
    count, pos = Emulator(mt).search(text)
    ref = oracle.set_search(pats, text)
    assert count == ref["count"]
    assert np.array_equal(np.sort(pos), ref["positions"])
    mt.close()


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("algo_name", ["AC", "WM"])
def test_tables_reproduce_golden(acwm, name, algo_name):
    from emu import Emulator
    g = GOLD[name]
    pats, text, alphabet = g["patterns"], g["text"], int(g["alphabet"])
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    mt = acwm.Matcher(algo, pats, alphabet)
    count, pos = Emulator(mt).search(text)
    assert count == int(g["ref_ac_count"])
    assert np.array_equal(np.sort(pos), g["positions"])
    if algo == acwm.AC and g.get("ref_states"):
        assert mt.info["n_states"] == int(g["ref_states"])
        assert mt.info["n_distinct"] == int(g["ref_distinct"])


def test_builder_choices_for_baseline_configs(acwm):
    """C1: the whole automaton fits shared memory with 3 symbols per lookup; C2: sampled WM."""
    dg = __import__("acwm_pkg").submodule("datagen")
    text = dg.text_host(1 << 20, 4, 1)
    c1 = acwm.Matcher(acwm.AC, dg.patterns_with_hits(text, 100, 8, 4, 2), 4).info
    assert c1["packed2bit"] == 1 and c1["exact_front"] == 1 and c1["stride"] == 3 and c1["depth"] == 8
    assert c1["table_in_smem"] == 1 and c1["smem_bytes"] <= 227 * 1024
    c2 = acwm.Matcher(acwm.WM, dg.patterns_with_hits(text, 1000, 16, 4, 3), 4).info
    assert c2["packed2bit"] == 1 and c2["stride"] == 8 and c2["depth"] == 9
    assert c2["smem_bytes"] <= 227 * 1024


def test_shim_preproc_ac_fills_reference_tables(acwm, oracle, have_ref):
    sm = __import__("acwm_pkg").submodule("smatcher")
    for case in (RANDOM_CASES[0], next(c for c in RANDOM_CASES if c[0] == "ac_protein_p100_m6")):
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        tr, sup, fin = sm.alloc_ac_tables(m, p, alphabet)
        table = sm.preproc_ac(pats, m, p, alphabet, tr, sup, fin)
        want = oracle.ref_ac(pats, alphabet, text[:1000], want_tables=True) if have_ref else \
            oracle.port_ac(pats, alphabet, text[:1000], want_tables=True)
        assert np.array_equal(tr, want["state_transition"])
        assert np.array_equal(sup, want["state_supply"])
        assert np.array_equal(fin, want["state_final"])
        assert table.contents.idcounter == want["n_states"]
        assert table.contents.patterncounter == want["n_distinct"]
        sm.free_ac(table, alphabet)


def test_shim_preproc_wu_fills_reference_tables(acwm, oracle, have_ref):
    sm = __import__("acwm_pkg").submodule("smatcher")
    for case in (RANDOM_CASES[1], next(c for c in RANDOM_CASES if c[0] == "wm_ascii_p1000_m8")):
        name, algo, alphabet, p, m, n, opts = case
        pats, text = make_case(case)
        for flat in (False, True):
            SHIFT, PV, PI, PS = sm.alloc_wu_tables(m, p, alphabet)
            if flat:
                sm.preproc_wu2(np.ascontiguousarray(pats).reshape(-1), m, p, alphabet, 3, SHIFT, PV, PI, PS)
            else:
                sm.preproc_wu(pats, m, p, alphabet, 3, SHIFT, PV, PI, PS)
            want = oracle.ref_wu(pats, alphabet, text[:1000], want_tables=True) if have_ref else \
                oracle.port_wu(pats, alphabet, text[:1000], want_tables=True)
            assert sm.wu_determine_shiftsize(alphabet) == want["shiftsize"]
            for k, got in (("SHIFT", SHIFT), ("PREFIX_size", PS), ("PREFIX_value", PV), ("PREFIX_index", PI)):
                assert np.array_equal(got, want[k]), k


def test_error_codes(acwm):
    pats = np.zeros((2, 4), np.uint8)
    with pytest.raises(acwm.AcwmError) as e:
        acwm.Matcher(acwm.AC, np.full((2, 4), 7, np.uint8), 4)          # symbol >= alphabet
    assert e.value.code == acwm.ERR_INVALID
    with pytest.raises(acwm.AcwmError) as e:
        acwm.Matcher(acwm.AC, [np.zeros(4, np.uint8), np.zeros(6, np.uint8)], 4)  # mixed-length AC
    assert e.value.code == acwm.ERR_UNSUPPORTED
    with pytest.raises(acwm.AcwmError) as e:
        acwm.Matcher(5, pats, 4)
    assert e.value.code == acwm.ERR_INVALID
    with pytest.raises(acwm.AcwmError) as e:
        acwm.Matcher(acwm.WM, pats, 300)
    assert e.value.code == acwm.ERR_INVALID
    with pytest.raises(acwm.AcwmError):
        acwm.Matcher(acwm.WM, np.zeros((0, 4), np.uint8), 4)


def test_no_cpu_fallback_without_gpu(acwm):
    """On a GPU-less box the search entry points must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mt = acwm.Matcher(acwm.AC, np.zeros((1, 4), np.uint8), 4)
    with pytest.raises(acwm.AcwmError) as e:
        mt.search_host(np.zeros(100, np.uint8))
    assert e.value.code == acwm.ERR_CUDA
    # the one-process multi-GPU flow too; its argument checks come first
    assert acwm.device_count() == 0
    mts = [acwm.Matcher(acwm.AC, np.zeros((1, 4), np.uint8), 4) for _ in range(2)]
    with pytest.raises(acwm.AcwmError) as e:
        acwm.search_host_sharded(mts, np.zeros(100, np.uint8))
    assert e.value.code == acwm.ERR_CUDA
    with pytest.raises(acwm.AcwmError) as e:
        acwm.search_host_sharded([mts[0], mts[0]], np.zeros(100, np.uint8))
    assert e.value.code == acwm.ERR_INVALID
    other = acwm.Matcher(acwm.AC, np.ones((1, 4), np.uint8), 4)
    with pytest.raises(acwm.AcwmError) as e:
        acwm.search_host_sharded([mts[0], other], np.zeros(100, np.uint8))
    assert e.value.code == acwm.ERR_INVALID


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cuda-aho-corasick-wu-manber_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "oracle/" not in src.replace("// ", ""), f
