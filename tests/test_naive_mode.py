"""GPC_RESULTS_NAIVE: the reference's SSE=OFF build (boxNaive, sobelNaive, gpcFilterNaive / gpcFilterTauNaive,
filter.hpp:157-293).  CPU tests pin the oracle's restatement against fixtures generated from the reference
compiled without -D_INTRINSICS_SSE (scripts/make_golden_naive.py) and, where oracle/_ref is present, against that
build itself; GPU tests compare the CUDA path with the oracle stage by stage and end to end."""
import json
import os
import zlib

import numpy as np
import pytest

from helpers import FORESTS, make_pair
from oraclelib import ROOT, Reference, digest, settings as osettings


@pytest.fixture(scope="module")
def nv_golden():
    with open(os.path.join(ROOT, "tests", "golden", "naive.json")) as f:
        return json.load(f)


def _written(gr):
    """grad positions the reference writes deterministically (row 0, (1,0), row h-1 and (h-2,w-1) excluded)."""
    g = gr.copy()
    h, w = g.shape
    g[0, :] = 0; g[1, 0] = 0; g[h - 1, :] = 0; g[h - 2, w - 1] = 0
    return g


@pytest.mark.parametrize("idx", range(6))
def test_oracle_naive_pairs(oracle, nv_golden, idx):
    rec = nv_golden["pairs"][idx]
    L, R = make_pair(rec)
    f = oracle.read_forest(FORESTS[rec["forest"]])
    s = osettings(rec["thr"], rec["disp_high"], rec["vt"], rec["epipolar"])
    supp, ncl, ncr = oracle.pair_naive(L, R, f, s)
    assert (ncl, ncr, len(supp), "%016x" % digest(supp)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"], rec["digest"])
    ht, _, _ = oracle.pair_naive(L, R, f, s, use_hashtable=True)
    assert (len(ht), "%016x" % digest(ht)) == (rec["n_supports_hashtable"], rec["digest_hashtable"])


def test_oracle_naive_stages(oracle, nv_golden):
    from opengpc_b200.synth import synth_pair
    for rec in nv_golden["stages"]:
        L, _ = synth_pair(rec["w"], rec["h"], rec["seed"])
        sm, gr, mk, st = oracle.stages_naive(L, oracle.read_forest(FORESTS[rec["forest"]]), rec["thr"])
        assert len(mk) == rec["n_mask"]
        assert zlib.crc32(sm.tobytes()) == rec["crc_smooth"] and zlib.crc32(_written(gr).tobytes()) == rec["crc_grad_written"]
        assert zlib.crc32(mk.astype("<i4").tobytes()) == rec["crc_mask"] and zlib.crc32(st.astype("<u4").tobytes()) == rec["crc_states"]


def test_oracle_naive_vs_reference_random(oracle):
    if not Reference.available(naive=True):
        pytest.skip("oracle/_ref/libgpc_ref_naive.so not built (needs /root/reference)")
    ref = Reference(naive=True)
    rng = np.random.default_rng(17)
    for it in range(6):
        w, h = int(rng.integers(4, 20)) * 16, int(rng.integers(30, 90))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if it % 2:
            img = (img // 16 * 16).astype(np.uint8)
        R = np.roll(img, -int(rng.integers(0, 9)), axis=1)
        thr = int(rng.choice([0, 3, 5, 10, 40, 255]))
        fname, epi = ["tau", "zero", "deep"][it % 3], bool(it % 2)
        if fname == "deep":
            fname = "tau"                                  # 32 tests fill the 32-bit code: allowed by the reference, same path
        f = oracle.read_forest(FORESTS[fname])
        sm, gr, mk = ref.preprocess(img, thr)
        osm, ogr, omk, ost = oracle.stages_naive(img, f, thr)
        assert np.array_equal(sm, osm) and np.array_equal(_written(gr), _written(ogr)) and np.array_equal(mk, omk)
        assert np.array_equal(ref.hash(img, thr, FORESTS[fname]), ost)
        rs, ncl, ncr, _ = ref.pair(img, R, FORESTS[fname], thr=thr, disp_high=50, vt=1, epipolar=epi)
        os_, ocl, ocr = oracle.pair_naive(img, R, f, osettings(thr, 50, 1, epi))
        assert (ncl, ncr) == (ocl, ocr) and np.array_equal(rs, os_), it


# ---- CUDA path --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g():
    import opengpc_b200
    return opengpc_b200


@pytest.fixture()
def nctx(g):
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=4) as c:
        c.set_result_mode(True)
        yield c


@pytest.mark.gpu
def test_naive_golden(g, nctx, nv_golden):
    for rec in nv_golden["pairs"]:
        L, R = make_pair(rec)
        nctx.set_forest(FORESTS[rec["forest"]])
        for ht in (False, True):
            s = g.make_settings(thr=rec["thr"], disp_high=rec["disp_high"], vt=rec["vt"], epipolar=rec["epipolar"], use_hashtable=ht)
            supp, ncl, ncr = nctx.match_pair(L, R, s)
            want = (rec["n_supports_hashtable"], rec["digest_hashtable"]) if ht else (rec["n_supports"], rec["digest"])
            assert (ncl, ncr) == (rec["n_cand_l"], rec["n_cand_r"])
            assert (len(supp), "%016x" % digest(supp)) == want, (rec, ht)


@pytest.mark.gpu
def test_naive_stages_vs_oracle(g, nctx, oracle):
    """smooth, grad (every position, including the linear-memory wrap at the first / last column), candidate list and
    states against the oracle on random and synthetic images."""
    from opengpc_b200.synth import synth_pair
    rng = np.random.default_rng(23)
    cases = [(rng.integers(0, 256, (int(rng.integers(30, 200)), 16 * int(rng.integers(2, 60))), dtype=np.uint8), int(t))
             for t in (0, 3, 5, 10, 40, 181, 255)]
    cases.append((synth_pair(1024, 436, 5)[0], 5))
    cases.append((np.full((40, 64), 200, np.uint8), 0))
    for fname in ("tau", "zero"):
        nctx.set_forest(FORESTS[fname])
        of = oracle.read_forest(FORESTS[fname])
        for img, thr in cases:
            osm, ogr, omk, ost = oracle.stages_naive(img, of, thr)
            sm, gr, mk = nctx.preprocess(img, thr)
            assert np.array_equal(sm, osm), (fname, img.shape, thr)
            assert np.array_equal(gr, ogr), (fname, img.shape, thr, np.argwhere(gr != ogr)[:5])
            assert np.array_equal(mk, omk)
            st, mk2 = nctx.hash(img, thr)
            assert np.array_equal(mk2, omk) and np.array_equal(st, ost), (fname, img.shape, thr)


@pytest.mark.gpu
def test_naive_pairs_vs_oracle(g, nctx, oracle):
    from opengpc_b200.synth import sparsify, synth_batch, synth_pair
    rng = np.random.default_rng(29)
    for it in range(10):
        nt = int(rng.choice([1, 2, 7, 8, 9, 10, 16, 17, 24, 25, 30, 31]))
        tests = [tuple(int(v) for v in rng.integers(-13, 14, 4)) for _ in range(nt)]
        tkind = it % 3
        taus = [0] * nt if tkind == 0 else [int(v) for v in (rng.integers(-300, 301, nt) if tkind == 1 else rng.integers(-12, 13, nt))]
        nctx.set_forest(g.make_forest(tests, taus))
        of = oracle.make_forest(tests, taus)
        w, h = 16 * int(rng.integers(4, 40)), int(rng.integers(30, 160))
        L = rng.integers(0, 256, (h // 3 + 1, w // 3 + 1), dtype=np.uint8).repeat(3, 0).repeat(3, 1)[:h, :w].copy()
        R = np.roll(L, -int(rng.integers(0, 12)), axis=1)
        R[rng.random((h, w)) < 0.03] ^= 0x11
        for epi, vt, ht in ((True, 0, False), (False, 2, False), (True, 0, True)):
            ref, ocl, ocr = oracle.pair_naive(L, R, of, osettings(5, 128, vt, epi), use_hashtable=ht)
            supp, ncl, ncr = nctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=vt, epipolar=epi, use_hashtable=ht))
            assert (ncl, ncr) == (ocl, ocr), (it, nt, tkind)
            assert np.array_equal(supp, ref), (it, nt, tkind, epi, ht, len(supp), len(ref))
    # a batch through the pipelined path
    nctx.set_forest(FORESTS["tau"])
    of = oracle.read_forest(FORESTS["tau"])
    imgs = synth_batch(512, 160, 4, seed0=70)
    supp, offs, ncand = nctx.match_batch(imgs, g.sparsematch_settings())
    for p in range(4):
        ref, ocl, ocr = oracle.pair_naive(imgs[p, 0], imgs[p, 1], of, osettings())
        assert (ncand[p, 0], ncand[p, 1]) == (ocl, ocr) and np.array_equal(supp[offs[p]:offs[p + 1]], ref), p


@pytest.mark.gpu
def test_result_mode_switching(g, oracle):
    """The mode is a property of the context: switching re-bakes the forest, SSE results come back unchanged, and a
    32-test forest is refused in the naive mode (bit 31 of a hash word is the candidate flag)."""
    from opengpc_b200 import capi
    from opengpc_b200.synth import synth_pair
    L, R = synth_pair(512, 128, 1234)
    of = oracle.read_forest(FORESTS["tau"])
    s = g.sparsematch_settings()
    with g.Context(device=0, max_w=512, max_h=128, max_batch=1) as c:
        c.set_forest(FORESTS["tau"])
        sse = c.match_pair(L, R, s)[0]
        assert np.array_equal(sse, oracle.pair(L, R, of, osettings())[0])
        c.set_result_mode(True)
        nv = c.match_pair(L, R, s)[0]
        assert np.array_equal(nv, oracle.pair_naive(L, R, of, osettings())[0]) and not np.array_equal(nv, sse)
        with pytest.raises(g.GpcError) as e:
            c.set_forest(FORESTS["deep"])
        assert e.value.status == capi.GPC_E_UNSUPPORTED
        assert np.array_equal(c.match_pair(L, R, s)[0], nv)          # the previous forest stays in place
        c.set_result_mode(False)
        assert np.array_equal(c.match_pair(L, R, s)[0], sse)
        c.set_forest(FORESTS["deep"])
        with pytest.raises(g.GpcError) as e:
            c.set_result_mode(True)
        assert e.value.status == capi.GPC_E_UNSUPPORTED


@pytest.mark.gpu
def test_naive_pyramid_and_resident_images(g, nctx, oracle):
    """The other entry points run the same pipeline: multi-level matching and resident images in the naive mode."""
    from opengpc_b200.synth import downsample2x, synth_pair
    of = oracle.read_forest(FORESTS["tau"])
    nctx.set_forest(FORESTS["tau"])
    L, R = synth_pair(1024, 436, 11)
    supp, offs, ncand = nctx.match_pyramid(L, R, 3, g.sparsematch_settings())
    dh, Ll, Rl = 128, L, R
    for l in range(3):
        ref, ocl, ocr = oracle.pair_naive(Ll, Rl, of, osettings(5, dh, 0, True))
        assert (ncand[l, 0], ncand[l, 1]) == (ocl, ocr) and np.array_equal(supp[offs[l]:offs[l + 1]], ref), l
        Ll, Rl, dh = downsample2x(Ll), downsample2x(Rl), dh // 2
    il, ir = nctx.upload(L), nctx.upload(R)
    sm, gr, mk = il.preprocess(5)
    osm, ogr, omk, _ = oracle.stages_naive(L, of, 5)
    assert np.array_equal(sm, osm) and np.array_equal(gr, ogr) and np.array_equal(mk, omk)
    got = nctx.match_images(il, ir, g.make_settings(thr=5, disp_high=128, vt=1, epipolar=False))[0]
    assert np.array_equal(got, oracle.pair_naive(L, R, of, osettings(5, 128, 1, False))[0])
    il.release(); ir.release()
