"""BASELINE configs[4] as a test: pattern count 10..100 000 x length 8..64, Aho-Corasick and Wu-Manber, on every GPU
of the box -- the role of the reference's execute.sh:9-56 run matrix, with the check the reference leaves to the eye
(main.c:297 vs cuda_ac.cu:675: CPU count beside GPU count) made exact: match count AND every position against the
oracle.  The text is sharded over all devices present (at least two shards, so a one-GPU box still runs the halo /
report_from / count-exchange logic) through the C ABI's device-resident multi-GPU entry points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 8 << 20  # symbols per point: the oracle finishes it in about a second
PS = (10, 100, 1000, 10000, 100000)
MS = (8, 16, 32, 64)


@pytest.fixture(scope="module")
def rig(acwm):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    dg = __import__("acwm_pkg").submodule("datagen")
    text = dg.text_host(N, 4, 1)
    n_dev = acwm.device_count()
    world = max(2, n_dev)
    return torch, dg, text, n_dev, world


@pytest.mark.parametrize("m", MS)
@pytest.mark.parametrize("p", PS)
@pytest.mark.parametrize("algo_name", ["AC", "WM"])
def test_sweep_point_matches_oracle(acwm, oracle, have_ref, rig, algo_name, p, m):
    torch, dg, text, n_dev, world = rig
    pats = dg.patterns_with_hits(text, p, m, 4, 2)
    ref = oracle.set_search(pats, text)
    cap = max(1024, ref["count"])
    algo = acwm.AC if algo_name == "AC" else acwm.WM
    mts = [acwm.Matcher(algo, pats, 4).upload(device=r % n_dev, pos_capacity=cap) for r in range(world)]
    try:
        acwm.peers_create(mts)
        bounds = [acwm.shard_bounds(N, world, r, m - 1) for r in range(world)]
        shards = [torch.from_numpy(text[s:s + l]).to(f"cuda:{r % n_dev}") for r, (s, l) in enumerate(bounds)]
        acwm.scan_device_sharded(mts, shards)
        acwm.scan_device_sharded(mts, shards)  # the exchange runs one scan behind: the second scan collects the first
        g, per = acwm.fetch_sharded(mts)
        assert g == ref["count"] == int(per.sum()), (algo_name, p, m, mts[0].info)
        got = []
        for r, (start, _) in enumerate(bounds):
            torch.cuda.set_device(r % n_dev)
            c, pos, _ = mts[r].fetch(cap=cap)
            assert c == int(per[r]) == pos.size
            got.append(pos + np.uint64(start))
        got = np.concatenate(got)
        assert np.array_equal(got, ref["positions"]), (algo_name, p, m, mts[0].info)
        # where the unmodified reference's preprocessing finishes quickly, its count on a prefix agrees too
        if have_ref and p <= 1000:
            pre = 1 << 20
            rc = (oracle.ref_ac if algo_name == "AC" else oracle.ref_wu)(pats, 4, text[:pre])["count"]
            assert rc == int(np.count_nonzero(ref["positions"] < pre))
    finally:
        acwm.peers_destroy(mts)
        torch.cuda.set_device(0)
        for mt in mts:
            mt.close()
