"""bench.py's contract line, checked where no GPU is needed: the reference arm (`--impl reference`, the unmodified
reference compiled into oracle/_ref timed on the host cores) prints ONE JSON line with the keys the driver reads, the
same `config` object our arm prints, and -- under torchrun -- only rank 0 speaks."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--text-mib", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT, check=True).stdout
    d = _line(out)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "impl", "per_algo"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert set(d["per_algo"]) == {"AC", "WM"} and d["value"] == min(v["value"] for v in d["per_algo"].values())
    # the config object is the one our arm prints for the same flags
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == json.loads(json.dumps(bench.workload_config(bench.DEFAULT_WORKLOAD, 4 << 20)))


def test_reference_arm_other_ranks_stay_silent(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--text-mib", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
