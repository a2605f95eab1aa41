"""useHashtable(true): the reference's hashtable matcher (inference.hpp:204-225, ndb::Hashmatch hashmatch.hpp:48-272).
CPU tests pin the oracle's restatement against fixtures generated from the compiled reference
(scripts/make_golden_hashtable.py) and, where oracle/_ref is present, against the reference itself; GPU tests
compare the CUDA path (bucket-sorted records replayed one bucket per thread) with the oracle and the fixtures."""
import json
import os

import numpy as np
import pytest

from helpers import FORESTS, make_pair, supp_to_i32
from oraclelib import ROOT, digest, settings as osettings


@pytest.fixture(scope="module")
def ht_golden():
    with open(os.path.join(ROOT, "tests", "golden", "hashtable.json")) as f:
        return json.load(f)


def test_oracle_hashmatch_kats(oracle, ht_golden):
    assert len(ht_golden["kats"]) >= 50
    for k in ht_golden["kats"]:
        got = oracle.hashmatch(np.array(k["src"], np.uint64), np.array(k["tar"], np.uint64))
        assert got.tolist() == k["pairs"], (k["src"], k["tar"])


@pytest.mark.parametrize("idx", range(6))
def test_oracle_hashtable_pairs(oracle, ht_golden, idx):
    rec = ht_golden["pairs"][idx]
    L, R = make_pair(rec)
    f = oracle.read_forest(FORESTS[rec["forest"]])
    supp = oracle.pair_hashtable(L, R, f, osettings(rec["thr"], rec["disp_high"], rec["vt"], rec["epipolar"]))
    assert (len(supp), "%016x" % digest(supp)) == (rec["n_supports"], rec["digest"])


def test_oracle_hashmatch_vs_reference_random(oracle, reference):
    rng = np.random.default_rng(11)
    B = np.uint64(214673)
    for case in range(120):
        ns, nt = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        kind = case % 4
        if kind == 0:
            pool = rng.integers(0, 2 ** 40, size=40, dtype=np.uint64)
        elif kind == 1:
            pool = rng.integers(0, 30, size=25, dtype=np.uint64) * B + np.uint64(5)      # a single bucket
        elif kind == 2:
            pool = rng.integers(0, 2 ** 63, size=1000, dtype=np.uint64)
        else:
            pool = rng.integers(0, 3, size=3, dtype=np.uint64) * B + np.uint64(11)
        src, tar = rng.choice(pool, ns), rng.choice(pool, nt)
        assert np.array_equal(oracle.hashmatch(src, tar), reference.hashmatch(src, tar)), case


# ---- CUDA path --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g():
    import opengpc_b200
    return opengpc_b200


@pytest.fixture(scope="module")
def ctx(g):
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=4) as c:
        yield c


@pytest.mark.gpu
def test_hashtable_golden(g, ctx, ht_golden):
    for rec in ht_golden["pairs"]:
        L, R = make_pair(rec)
        ctx.set_forest(FORESTS[rec["forest"]])
        s = g.make_settings(thr=rec["thr"], disp_high=rec["disp_high"], vt=rec["vt"], epipolar=rec["epipolar"], use_hashtable=True)
        supp, _, _ = ctx.match_pair(L, R, s)
        assert (len(supp), "%016x" % digest(supp)) == (rec["n_supports"], rec["digest"]), rec


@pytest.mark.gpu
@pytest.mark.parametrize("forest", ["tau", "zero", "deep"])
def test_hashtable_vs_oracle(g, ctx, oracle, forest):
    from opengpc_b200.synth import sparsify, synth_pair
    of = oracle.read_forest(FORESTS[forest])
    ctx.set_forest(FORESTS[forest])
    for (w, h, seed, thr, dh, vt, epi, sparse) in [(1024, 436, 77, 5, 128, 0, True, False), (512, 200, 7, 10, 64, 1, False, False),
                                                   (640, 120, 9, 5, 128, 3, False, True), (256, 64, 3, 0, 1000, 100, True, False),
                                                   (256, 40, 4, 0, 128, 0, True, False)]:
        L, R = synth_pair(w, h, seed)
        if sparse:
            L, R = sparsify(L), sparsify(R)
        ref = oracle.pair_hashtable(L, R, of, osettings(thr, dh, vt, epi))
        supp, _, _ = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi, use_hashtable=True))
        assert np.array_equal(supp, ref), (forest, w, h, epi, len(supp), len(ref))


@pytest.mark.gpu
def test_hashtable_overfull_buckets(g, ctx, oracle):
    """Forests of one to four tests: a handful of distinct states, so every bucket overflows the 10-element cap
    and the walk's early-return / skip rules (hashmatch.hpp:171-195) decide the result."""
    from opengpc_b200.synth import synth_pair
    rng = np.random.default_rng(3)
    for n_tests in (1, 2, 4):
        tests = [tuple(int(v) for v in rng.integers(-6, 7, 4)) for _ in range(n_tests)]
        taus = [int(v) for v in rng.integers(-4, 5, n_tests)]
        ctx.set_forest(g.make_forest(tests, taus))
        of = oracle.make_forest(tests, taus)
        for (w, h, seed, epi) in [(256, 80, 1, True), (256, 80, 2, False), (1024, 100, 3, True)]:
            L, R = synth_pair(w, h, seed)
            ref = oracle.pair_hashtable(L, R, of, osettings(5, 128, 2, epi))
            supp, _, _ = ctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=2, epipolar=epi, use_hashtable=True))
            assert np.array_equal(supp, ref), (n_tests, w, h, epi, len(supp), len(ref))


@pytest.mark.gpu
def test_hashtable_batch_and_correspondences(g, ctx, oracle):
    from opengpc_b200.synth import synth_batch
    of = oracle.read_forest(FORESTS["tau"])
    ctx.set_forest(FORESTS["tau"])
    imgs = synth_batch(512, 160, 4, seed0=50)
    for epi in (True, False):
        s = g.make_settings(thr=5, disp_high=128, vt=1, epipolar=epi, use_hashtable=True)
        supp, offs, _ = ctx.match_batch(imgs, s)
        for p in range(4):
            ref = oracle.pair_hashtable(imgs[p, 0], imgs[p, 1], of, osettings(5, 128, 1, epi))
            assert np.array_equal(supp[offs[p]:offs[p + 1]], ref), (epi, p)
        il, ir = ctx.upload(imgs[0, 0]), ctx.upload(imgs[0, 1])
        corr = ctx.correspond_images(il, ir, s)
        got = np.stack([corr["xs"], corr["ys"], corr["xt"], corr["yt"]], 1)
        assert np.array_equal(got, oracle.correspondences_hashtable(imgs[0, 0], imgs[0, 1], of, osettings(5, 128, 1, epi)))
        supp1, _, _ = ctx.match_images(il, ir, s)
        assert np.array_equal(supp1, supp[offs[0]:offs[1]])
        il.release(); ir.release()


@pytest.mark.gpu
def test_hashmatch_explicit_keys(g, ctx, oracle, ht_golden):
    """gpc_hashmatch: the reference's known answers, then random key lists with heavy bucket collisions."""
    for k in ht_golden["kats"]:
        got = ctx.hashmatch(np.array(k["src"], np.uint64), np.array(k["tar"], np.uint64))
        assert got.tolist() == k["pairs"], (k["src"], k["tar"])
    rng = np.random.default_rng(21)
    B = np.uint64(214673)
    for case in range(40):
        ns, nt = int(rng.integers(1, 20000)), int(rng.integers(1, 20000))
        kind = case % 4
        if kind == 0:
            pool = rng.integers(0, 2 ** 40, size=3000, dtype=np.uint64)
        elif kind == 1:
            pool = rng.integers(0, 300, size=200, dtype=np.uint64) * B + rng.integers(0, 9, size=200, dtype=np.uint64)
        elif kind == 2:
            pool = rng.integers(0, 2 ** 63, size=100000, dtype=np.uint64)
        else:
            pool = rng.integers(0, 2 ** 20, size=5000, dtype=np.uint64)
        src, tar = rng.choice(pool, ns), rng.choice(pool, nt)
        assert np.array_equal(ctx.hashmatch(src, tar), oracle.hashmatch(src, tar)), case


@pytest.mark.gpu
def test_hashtable_many_pairs_per_launch(g, oracle):
    """20 Sintel-sized pairs in one chunk: more tiles per pair (185) than blocks per pair (120), so every sort /
    replay kernel walks several tiles per block, and the packed output spans the pairs of the chunk."""
    from opengpc_b200.synth import synth_batch
    n = 20
    imgs = np.tile(synth_batch(1024, 436, 4, seed0=900), (n // 4, 1, 1, 1))
    imgs[7] = synth_batch(1024, 436, 1, seed0=77)[0]
    of = oracle.read_forest(FORESTS["tau"])
    s = g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=True)
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=n) as c:
        c.set_forest(FORESTS["tau"])
        supp, offs, _ = c.match_batch(imgs, s)
    refs = {}
    for p in range(n):
        key = 7 if p == 7 else p % 4
        if key not in refs:
            refs[key] = oracle.pair_hashtable(imgs[p, 0], imgs[p, 1], of, osettings())
        assert np.array_equal(supp[offs[p]:offs[p + 1]], refs[key]), p
