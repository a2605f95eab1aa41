"""GPU parity tests of the rows added after the first slice: global matching mode (radix sort +
segmented scan), the sort matcher cross-checked against the per-row matcher, explicit-key
findCorrespondences, evalFastMaskOnSubsetSSE on a caller's smoothed image, resident images."""
import numpy as np
import pytest

from helpers import FORESTS, make_pair, supp_to_i32
from oraclelib import settings as osettings

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import opengpc_b200
    return opengpc_b200


@pytest.fixture(scope="module")
def ctx(g):
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=3) as c:
        yield c


def test_global_mode_small_cases(g, ctx, small_cases):
    """Fixtures generated from the compiled reference with epipolarMode(false)."""
    from helpers import write_forest
    n = 0
    for name, c in small_cases.items():
        thr, epi, vt, dh = (int(v) for v in c["cfg"])
        if epi:
            continue
        ctx.set_forest(write_forest(c["forest"]))
        supp, ncl, ncr = ctx.match_pair(c["L"], c["R"], g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=False))
        assert (ncl, ncr) == (len(c["maskL"]), len(c["maskR"]))
        assert np.array_equal(supp_to_i32(supp), c["supp"].reshape(-1, 3)), name
        n += 1
    assert n >= 1


@pytest.mark.parametrize("forest", ["tau", "zero", "deep"])
def test_global_mode_vs_oracle(g, ctx, oracle, forest):
    """Library default settings (inference.hpp:74-89: global mode, vt=1, thr=10) and variations."""
    from opengpc_b200.synth import sparsify, synth_pair
    of = oracle.read_forest(FORESTS[forest])
    ctx.set_forest(FORESTS[forest])
    for (w, h, seed, thr, dh, vt, sparse) in [(1024, 436, 1234, 10, 128, 1, False), (512, 200, 7, 5, 64, 0, False),
                                              (640, 120, 9, 5, 128, 3, True), (256, 64, 3, 0, 1000, 100, False)]:
        L, R = synth_pair(w, h, seed)
        if sparse:
            L, R = sparsify(L), sparsify(R)
        ref, ocl, ocr = oracle.pair(L, R, of, osettings(thr, dh, vt, False))
        supp, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=False))
        assert (ncl, ncr) == (ocl, ocr)
        assert np.array_equal(supp, ref), (forest, w, h, len(supp), len(ref))


def test_golden_global(g, golden, oracle):
    recs = [r for r in golden["pairs"] if not r["epipolar"]]
    if not recs:
        pytest.skip("no global-mode golden records")
    for rec in recs:
        L, R = make_pair(rec)
        with g.Context(device=0, max_w=rec["w"], max_h=rec["h"], max_batch=1) as c:
            c.set_forest(FORESTS[rec["forest"]])
            supp, ncl, ncr = c.match_pair(L, R, g.make_settings(thr=5, disp_high=rec["disp_high"], vt=rec["vt"], epipolar=False))
        assert (ncl, ncr, len(supp)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), rec
        assert "%016x" % oracle.digest(supp) == rec["digest"], rec


def test_sort_matcher_equals_row_matcher(g, ctx, oracle):
    """Epipolar mode through the radix-sort matcher (64-bit keys y<<32|state) must give exactly the
    per-row matcher's result -- two independent implementations of inference.hpp:227-254."""
    from opengpc_b200.synth import sparsify, synth_batch
    imgs = synth_batch(1024, 436, 3, seed0=1234)
    imgs[1, 0], imgs[1, 1] = sparsify(imgs[1, 0]), sparsify(imgs[1, 1])
    s = g.sparsematch_settings()
    for forest in ("tau", "zero"):
        ctx.set_forest(FORESTS[forest])
        ctx.set_matcher(g.MATCHER_AUTO)
        a, oa, na = ctx.match_batch(imgs, s)
        for other in (g.MATCHER_SORT, g.MATCHER_ROWS_GENERAL):      # radix sort; every row through the general row kernel
            ctx.set_matcher(other)
            try:
                b, ob, nb = ctx.match_batch(imgs, s)
            finally:
                ctx.set_matcher(g.MATCHER_AUTO)
            assert np.array_equal(oa, ob) and np.array_equal(na, nb)
            assert np.array_equal(a, b), (forest, other)
    of = oracle.read_forest(FORESTS["zero"])
    ref, _, _ = oracle.pair(imgs[2, 0], imgs[2, 1], of, osettings())
    assert np.array_equal(b[ob[2]:ob[3]], ref)


def test_find_correspondences_largest_key(g, ctx, oracle):
    """A right key of 2^64 - 1 is legal (explicit 64-bit keys): the tail rules of inference.hpp:243-249 still apply to it."""
    top = (1 << 64) - 1
    for src, tar in (([1, top], [1, top]), ([top], [5, top, top]), ([3, top], [3, top, top, top]), ([top, 7], [7, top - 1, top])):
        want = oracle.find_correspondences(np.array(src, np.uint64), np.array(tar, np.uint64))
        got = ctx.find_correspondences(src, tar)
        assert got.tolist() == want.tolist(), (src, tar, got.tolist(), want.tolist())


def test_match_hash_images_ignores_border_flags(g, ctx, oracle):
    """Candidate flags outside the 13-pixel interior (which the reference's border lambda never produces,
    inference.hpp:318-330) are cleared on the staged copy: same result as without them, in every matcher."""
    from opengpc_b200.synth import synth_pair
    L, R = synth_pair(256, 64, 77)
    ctx.set_forest(FORESTS["tau"])
    hl, hr = ctx.hash(L, 5, want_image=True)[2], ctx.hash(R, 5, want_image=True)[2]
    s = g.sparsematch_settings()
    base = ctx.match_hash_images(hl, hr, s)
    dl, dr = hl.copy(), hr.copy()
    rng = np.random.default_rng(3)
    for img in (dl, dr):
        img[:13, :] = rng.integers(0, 1 << 31, (13, 256), dtype=np.uint32) | 0x80000000
        img[-13:, :] = 0x80000001
        img[:, :13] |= 0x80000000
        img[:, -13:] = rng.integers(0, 1 << 31, (64, 13), dtype=np.uint32) | 0x80000000
    for matcher in (g.MATCHER_AUTO, g.MATCHER_ROWS_GENERAL, g.MATCHER_SORT):
        ctx.set_matcher(matcher)
        try:
            assert np.array_equal(ctx.match_hash_images(dl, dr, s), base), matcher
            assert np.array_equal(ctx.match_hash_images(dl, dr, g.make_settings(thr=5, disp_high=128, vt=1, epipolar=False)),
                                  ctx.match_hash_images(hl, hr, g.make_settings(thr=5, disp_high=128, vt=1, epipolar=False))), matcher
        finally:
            ctx.set_matcher(g.MATCHER_AUTO)


def test_find_correspondences_kats_and_random(g, ctx, golden, oracle):
    for kat in golden["kats"]:
        got = ctx.find_correspondences(kat["src"], kat["tar"])
        assert got.tolist() == [list(p) for p in kat["pairs"]], kat
    rng = np.random.default_rng(5)
    for it in range(8):
        ns, nt = int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
        hi = int(rng.choice([8, 200, 5000, 1 << 20]))
        src = rng.integers(0, hi, ns).astype(np.uint64)
        tar = rng.integers(0, hi, nt).astype(np.uint64)
        if it % 2:
            src |= rng.integers(0, 300, ns).astype(np.uint64) << np.uint64(32)      # epipolar-style 64-bit keys
            tar |= rng.integers(0, 300, nt).astype(np.uint64) << np.uint64(32)
        if it == 3:
            tar[-1] = tar.max() + np.uint64(1)                                        # unique last key: never matches
            src[0] = tar[-1]
        if it == 5:
            tar[:2] = tar.max() + np.uint64(1)                                        # tail-dup-2
            src[0] = tar[0]
        want = oracle.find_correspondences(src, tar)
        got = ctx.find_correspondences(src, tar)
        assert np.array_equal(got, want), (it, len(got), len(want))
    assert len(ctx.find_correspondences([1, 2], [])) == 0
    assert len(ctx.find_correspondences([1], [1])) == 0


def test_hash_smooth_seam(g, ctx, oracle):
    """gpc_hash_smooth == gpcFilter / gpcFilterTau on a caller-provided smoothed image and index list."""
    rng = np.random.default_rng(3)
    for forest in ("tau", "zero", "deep"):
        of = oracle.read_forest(FORESTS[forest])
        ctx.set_forest(FORESTS[forest])
        smooth = rng.integers(0, 256, (90, 272), dtype=np.uint8)     # any image, not necessarily a box-filter output
        ys, xs = np.meshgrid(np.arange(13, 90 - 13), np.arange(13, 272 - 13), indexing="ij")
        idx = (ys * 272 + xs).reshape(-1)
        idx = idx[rng.random(len(idx)) < 0.4].astype(np.int32)
        want = oracle.hash(smooth, of, idx)
        got = ctx.hash_smooth(smooth, idx)
        assert np.array_equal(got, want), forest
        perm = rng.permutation(len(idx))                              # order of idx is the caller's
        assert np.array_equal(ctx.hash_smooth(smooth, idx[perm]), want[perm])


def test_resident_images(g, ctx, oracle):
    from opengpc_b200.synth import synth_pair
    L, R = synth_pair(768, 200, 77)
    of = oracle.read_forest(FORESTS["tau"])
    ctx.set_forest(FORESTS["tau"])
    il, ir = ctx.upload(L), ctx.upload(R)
    sm, gr, mk = il.preprocess(5)
    osm, ogr, omk, _ = oracle.stages(L, of, 5)
    assert np.array_equal(sm, osm) and np.array_equal(gr, ogr) and np.array_equal(mk, omk)
    for epi, vt in ((True, 0), (False, 1)):
        s = g.make_settings(thr=5, disp_high=128, vt=vt, epipolar=epi)
        ref, ocl, ocr = oracle.pair(L, R, of, osettings(5, 128, vt, epi))
        supp, ncl, ncr = ctx.match_images(il, ir, s)
        assert (ncl, ncr) == (ocl, ocr) and np.array_equal(supp, ref)
        corr = ctx.correspond_images(il, ir, s)
        want = oracle.correspondences(L, R, of, osettings(5, 128, vt, epi))
        got = np.stack([corr["xs"], corr["ys"], corr["xt"], corr["yt"]], 1)
        assert np.array_equal(got, want), (epi, len(got), len(want))
    il.release(); ir.release()
    # the cache of a resident image: every combination of cached / stale state gives the oracle's result, and the
    # cached paths launch fewer kernels (matcher only: no A1, no A2)
    ref5, _, _ = oracle.pair(L, R, of, osettings(5, 128, 0, True))
    ref9, _, _ = oracle.pair(L, R, of, osettings(9, 128, 0, True))
    of0 = oracle.read_forest(FORESTS["zero"])
    ref5z, _, _ = oracle.pair(L, R, of0, osettings(5, 128, 0, True))
    s5, s9 = g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True), g.make_settings(thr=9, disp_high=128, vt=0, epipolar=True)
    il, ir = ctx.upload(L), ctx.upload(R)
    n0 = ctx.launches
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5)                # nothing cached: both kernels
    cold = ctx.launches - n0
    _, _, mk = il.preprocess(5, images=False)
    ir.preprocess(5, images=False)
    assert np.array_equal(mk, omk)
    n0 = ctx.launches
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5)                # hash images cached: matcher only
    warm = ctx.launches - n0
    assert warm == cold - 2, (cold, warm)
    assert np.array_equal(ctx.match_images(il, ir, s9)[0], ref9)                # other threshold: recomputed from the raw image
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5)
    ctx.set_forest(FORESTS["zero"])                                             # other forest: kernel A2 only, then cached again
    n0 = ctx.launches
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5z)
    assert ctx.launches - n0 == cold - 1
    n0 = ctx.launches
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5z)
    assert ctx.launches - n0 == warm
    ctx.set_forest(FORESTS["tau"])
    sm2, gr2 = il.fetch(5)
    assert np.array_equal(sm2, osm) and np.array_equal(gr2, ogr)
    assert np.array_equal(ctx.match_images(il, ir, s5)[0], ref5)
    il.release(); ir.release()
    with g.Context(device=0, max_w=768, max_h=200, max_batch=1) as other:
        io = other.upload(L)
        with pytest.raises(g.GpcError):
            ctx.match_images(io, io, g.sparsematch_settings())        # image of another context
        io.release()


def test_row_lists_survive_mixed_slices(g, oracle):
    """The row matcher's kernels pass rows to each other through lists that live in the context (rows for the general
    kernel, rows for the block-wide ordering kernel).  A whole-batch launch followed by a chunked one on the same
    context uses overlapping slices of those buffers: the dense pairs here (left == right: hundreds of matches per row)
    fill the lists, and the chunked run must still see clean list headers."""
    from opengpc_b200.synth import synth_batch
    imgs = synth_batch(512, 96, 40, seed0=900)
    imgs[::2, 1] = imgs[::2, 0]                                     # every other pair: identical images
    s = g.sparsematch_settings()
    of = oracle.read_forest(FORESTS["tau"])
    refs = {p: oracle.pair(imgs[p, 0], imgs[p, 1], of, osettings())[0] for p in (0, 1, 38, 39)}
    assert len(refs[0]) > 20 * len(refs[1])
    with g.Context(device=0, max_w=512, max_h=96, max_batch=40) as c:
        c.set_forest(FORESTS["tau"])
        c.enable_kernel_timing(True)                                # timing on: one launch sequence over the whole batch
        a, oa, _ = c.match_batch(imgs, s)
        a = a.copy()
        c.enable_kernel_timing(False)                               # timing off: chunks on slices of the same buffers
        for rep in range(2):
            b, ob, _ = c.match_batch(imgs, s)
            assert np.array_equal(oa, ob) and np.array_equal(a, b), rep
        for p, ref in refs.items():
            assert np.array_equal(b[ob[p]:ob[p + 1]], ref), p


@pytest.mark.parametrize("mode", ["rows", "global", "hashtable"])
def test_pipelined_batch(g, oracle, monkeypatch, mode):
    """gpc_match_batch splits large batches into chunks that rotate over three streams (upload,
    kernels and download overlap).  Chunk size 2 forces that path on a 7-pair batch, including a
    pair without candidates and a too-small output buffer; per-row matcher, radix-sort matcher (global
    mode) and the hashtable matcher all run through it."""
    from opengpc_b200 import capi
    from opengpc_b200.synth import synth_batch
    monkeypatch.setenv("GPC_CHUNK_PAIRS", "2")
    imgs = synth_batch(320, 90, 7, seed0=300)
    imgs[3] = 50
    of = oracle.read_forest(FORESTS["tau"])
    if mode == "rows":
        s, os_ = g.sparsematch_settings(), osettings()
    elif mode == "global":
        s, os_ = g.make_settings(thr=5, disp_high=128, vt=1, epipolar=False), osettings(5, 128, 1, False)
    else:
        s, os_ = g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=True), osettings()
    with g.Context(device=0, max_w=320, max_h=90, max_batch=7) as c:
        c.set_forest(FORESTS["tau"])
        for rep in range(3):
            supp, offsets, ncand = c.match_batch(imgs, s)
            for p in range(7):
                if mode == "hashtable":
                    ref = oracle.pair_hashtable(imgs[p, 0], imgs[p, 1], of, os_)
                else:
                    ref, ocl, ocr = oracle.pair(imgs[p, 0], imgs[p, 1], of, os_)
                    assert (ncand[p, 0], ncand[p, 1]) == (ocl, ocr)
                assert np.array_equal(supp[offsets[p]:offsets[p + 1]], ref), (mode, rep, p)
        assert offsets[4] == offsets[3]
        small = np.empty(10, g.SUPPORT_DTYPE)
        with pytest.raises(g.GpcError) as e:
            c.match_batch(imgs, s, out=small)
        assert e.value.status == capi.GPC_E_CAPACITY
        supp2, offsets2, _ = c.match_batch(imgs, s)          # the context stays usable afterwards
        assert np.array_equal(supp2, supp) and np.array_equal(offsets2, offsets)


def test_pyramid_vs_golden_and_oracle(g, oracle):
    """BASELINE configs[3]: levels built on the device, the single-level path per level.  Golden values come
    from the unmodified reference run per level on identically down-sampled images (scripts/make_golden_pyramid.py)."""
    import json
    import os
    from oraclelib import ROOT
    from opengpc_b200.synth import downsample2x, synth_pair
    with open(os.path.join(ROOT, "tests", "golden", "pyramid.json")) as f:
        gold = json.load(f)
    of = oracle.read_forest(FORESTS["tau"])
    for case in gold["cases"]:
        w, h, nl = case["w"], case["h"], len(case["levels"])
        L, R = synth_pair(w, h, case["seed"])
        with g.Context(device=0, max_w=w, max_h=h, max_batch=1) as c:
            c.set_forest(FORESTS["tau"])
            supp, offs, ncand = c.match_pyramid(L, R, nl, g.sparsematch_settings())
        for rec in case["levels"]:
            l = rec["level"]
            lev = supp[offs[l]:offs[l + 1]]
            assert (ncand[l, 0], ncand[l, 1], len(lev)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), rec
            assert "%016x" % oracle.digest(lev) == rec["digest"], rec
            if w <= 1024:                                        # the oracle itself on the down-sampled images
                ref, _, _ = oracle.pair(L, R, of, osettings(5, rec["disp_high"], 0, True))
                assert np.array_equal(lev, ref)
            L, R = downsample2x(L), downsample2x(R)


def test_device_batch_global_mode(g, oracle):
    """gpc_match_batch_device with epipolarMode(false): the radix-sort matcher pair by pair, strided output."""
    import torch
    from opengpc_b200.synth import synth_batch
    imgs = synth_batch(256, 80, 3, seed0=41)
    of = oracle.read_forest(FORESTS["tau"])
    cap = 6000
    s = g.make_settings(thr=5, disp_high=100, vt=2, epipolar=False)
    with g.Context(device=0, max_w=256, max_h=80, max_batch=3) as ctx:
        ctx.set_forest(FORESTS["tau"])
        d_img = torch.from_numpy(imgs).cuda()
        d_out = torch.zeros((3, cap, 3), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(3, dtype=torch.int32, device="cuda")
        d_nc = torch.zeros((3, 2), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.match_batch_device(d_img.data_ptr(), 3, 256, 80, s, d_out.data_ptr(), cap, d_n.data_ptr(), d_nc.data_ptr())
        ctx.synchronize()
        n, out, nc = d_n.cpu().numpy(), d_out.cpu().numpy(), d_nc.cpu().numpy()
    for p in range(3):
        ref, ocl, ocr = oracle.pair(imgs[p, 0], imgs[p, 1], of, osettings(5, 100, 2, False))
        assert n[p] == len(ref) and (nc[p, 0], nc[p, 1]) == (ocl, ocr)
        assert np.array_equal(out[p, :n[p]].copy().view(g.SUPPORT_DTYPE).reshape(-1), ref)


def test_edge_shapes(g, oracle):
    """Minimum and ragged sizes: width 32 (one tile column, mostly border), heights with 0 / 1 / 2 candidate
    rows, odd heights, a row stride larger than the width, and a width that is not a tile multiple."""
    rng = np.random.default_rng(21)
    of = oracle.read_forest(FORESTS["tau"])
    with g.Context(device=0, max_w=1040, max_h=70, max_batch=1) as ctx:
        ctx.set_forest(FORESTS["tau"])
        for (w, h) in [(32, 40), (48, 26), (48, 27), (64, 28), (144, 29), (1040, 31), (400, 69), (16, 64)]:
            L = rng.integers(0, 256, (h, w), dtype=np.uint8)
            R = np.roll(L, -2, axis=1)
            for epi in (True, False):
                ref, ocl, ocr = oracle.pair(L, R, of, osettings(5, 128, 1 if not epi else 0, epi))
                supp, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=1 if not epi else 0, epipolar=epi))
                assert (ncl, ncr) == (ocl, ocr), (w, h, epi)
                assert np.array_equal(supp, ref), (w, h, epi, len(supp), len(ref))
            sm, gr, mk = ctx.preprocess(L, 5)
            osm, ogr, omk, _ = oracle.stages(L, of, 5)
            assert np.array_equal(sm, osm) and np.array_equal(gr, ogr) and np.array_equal(mk, omk), (w, h)
        # strided input rows (gpc_match_pair's stride argument)
        import ctypes as C
        from opengpc_b200 import capi
        w, h, stride = 96, 50, 128
        big_l = rng.integers(0, 256, (h, stride), dtype=np.uint8)
        big_r = np.roll(big_l, -3, axis=1)
        L, R = np.ascontiguousarray(big_l[:, :w]), np.ascontiguousarray(big_r[:, :w])
        ref, _, _ = oracle.pair(L, R, of, osettings())
        out = np.empty(w * h, g.SUPPORT_DTYPE)
        n = C.c_int(0)
        s = g.sparsematch_settings()
        rc = ctx.lib.gpc_match_pair(ctx._h, C.c_void_p(big_l.ctypes.data), C.c_void_p(big_r.ctypes.data), w, h, stride, C.byref(s),
                                    C.c_void_p(out.ctypes.data), C.c_int(len(out)), C.byref(n), None, None)
        assert rc == capi.GPC_OK and np.array_equal(out[:n.value], ref)


def test_forest_specialised_kernel(g, oracle, monkeypatch):
    """gpc_set_forest rebuilds the hashing kernel with the forest baked in (NVRTC).  The specialised
    and the generic precompiled kernel must both be bit-exact, for every forest type."""
    from opengpc_b200.synth import synth_pair
    L, R = synth_pair(640, 150, 8)
    rng = np.random.default_rng(4)
    forests = {n: (FORESTS[n], oracle.read_forest(FORESTS[n])) for n in ("tau", "zero", "deep")}
    tests = [tuple(int(v) for v in rng.integers(-13, 14, 4)) for _ in range(11)]
    taus = [0, -128, 127, 0, 3, 0, 0, -1, 0, 0, 55]
    with g.Context(device=0, max_w=640, max_h=150, max_batch=1) as c:
        for jit in ("1", "0"):
            monkeypatch.setenv("GPC_JIT", jit)
            for name, (path, of) in forests.items():
                c.set_forest(path)
                assert c.jit_status == ("specialised" if jit == "1" else "generic: disabled by GPC_JIT=0"), c.jit_status
                ref, ocl, ocr = oracle.pair(L, R, of, osettings())
                supp, ncl, ncr = c.match_pair(L, R, g.sparsematch_settings())
                assert (ncl, ncr) == (ocl, ocr) and np.array_equal(supp, ref), (jit, name)
            c.set_forest(g.make_forest(tests, taus))
            st, mk = c.hash(L, 5)
            _, _, omk, ost = oracle.stages(L, oracle.make_forest(tests, taus), 5)
            assert np.array_equal(mk, omk) and np.array_equal(st, ost), jit


def test_degenerate_states(g, oracle):
    """Rows where nearly every candidate carries the same state (periodic texture, short forests): the row
    matcher's buckets hold hundreds of equal entries and its ordering pass takes the skewed-state path."""
    w, h = 1024, 60
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    stripes = (((x // 3) % 2) * 200 + (y % 2) * 30).astype(np.uint8) * np.ones((h, 1), np.uint8)
    rng = np.random.default_rng(9)
    noisy = stripes.copy()
    noisy[rng.random((h, w)) < 0.02] ^= 0x40
    with g.Context(device=0, max_w=w, max_h=h, max_batch=1) as ctx:
        for tests, taus in (([(1, 0, -1, 0)], [0]), ([(2, 1, -3, 0), (0, 2, 1, -2), (5, 5, -5, -5)], [3, -2, 0]),
                            ([tuple(int(v) for v in rng.integers(-13, 14, 4)) for _ in range(12)], [0] * 12)):
            ctx.set_forest(g.make_forest(tests, taus))
            of = oracle.make_forest(tests, taus)
            for L, R in ((stripes, np.roll(stripes, -6, axis=1)), (noisy, np.roll(noisy, -4, axis=1)), (noisy, stripes)):
                for epi in (True, False):
                    ref, ocl, ocr = oracle.pair(L, R, of, osettings(5, 128, 0 if epi else 1, epi))
                    supp, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=0 if epi else 1, epipolar=epi))
                    assert (ncl, ncr) == (ocl, ocr)
                    assert np.array_equal(supp, ref), (len(tests), epi, len(supp), len(ref))


def test_overflow_lists_beyond_the_tail_slice(g, ctx, oracle):
    """Crafted state rows (gpc_match_hash_images): n states twice on the left and once on the right put 2n + n entries
    on the row's overflow lists.  Up to 256 entries the warp-per-row tail kernel joins them in its 3 KB slice; beyond
    that (each list still within its own 256-entry capacity) the row goes to the general kernel's list; beyond 256
    per list the fast kernel hands it over itself.  All three must agree with the general matcher, the sort matcher
    and, through them, the oracle's semantics: duplicated states never match, the unique ones around them do."""
    w, h = 1024, 40
    rng = np.random.default_rng(41)
    s = g.sparsematch_settings()
    for n_dup in (40, 90, 120, 200):
        hl = np.zeros((h, w), np.uint32)
        hr = np.zeros((h, w), np.uint32)
        for y in range(13, h - 13):
            states = rng.choice(1 << 30, size=n_dup + 300, replace=False).astype(np.uint32)
            dup, uniq = states[:n_dup], states[n_dup:]
            xs = rng.permutation(np.arange(13, w - 13))
            xl_dup, xl_uniq = xs[:2 * n_dup], xs[2 * n_dup:2 * n_dup + 300]
            hl[y, xl_dup] = np.repeat(dup, 2) | 0x80000000
            hl[y, xl_uniq] = uniq | 0x80000000
            xr = rng.permutation(np.arange(13, w - 13))
            hr[y, xr[:n_dup]] = dup | 0x80000000
            # the unique states sit within the disparity range of their partners
            xr_uniq = np.clip(xl_uniq - rng.integers(0, 60, 300), 13, w - 14)
            free = hr[y, xr_uniq] == 0
            _, first = np.unique(xr_uniq, return_index=True)
            keep = np.zeros(300, bool); keep[first] = True
            keep &= free
            hr[y, xr_uniq[keep]] = uniq[keep] | 0x80000000
        results = {}
        for matcher in (g.MATCHER_AUTO, g.MATCHER_ROWS_GENERAL, g.MATCHER_SORT):
            ctx.set_matcher(matcher)
            try:
                results[matcher] = ctx.match_hash_images(hl, hr, s)
            finally:
                ctx.set_matcher(g.MATCHER_AUTO)
        assert len(results[g.MATCHER_SORT]) > 100 * (h - 26) // 2, n_dup
        assert np.array_equal(results[g.MATCHER_AUTO], results[g.MATCHER_SORT]), n_dup
        assert np.array_equal(results[g.MATCHER_ROWS_GENERAL], results[g.MATCHER_SORT]), n_dup


def test_very_wide_rows(g, oracle):
    """Rows wider than 4096 pixels take the 2-quads-per-thread / 1024-thread shape of the row matcher; 8192 is
    the widest supported row (13-bit column fields)."""
    from opengpc_b200 import capi
    rng = np.random.default_rng(33)
    of = oracle.read_forest(FORESTS["tau"])
    with g.Context(device=0, max_w=8192, max_h=48, max_batch=1) as ctx:
        ctx.set_forest(FORESTS["tau"])
        for (w, h) in [(4112, 45), (6016, 40), (8192, 48)]:
            b = 5
            L = rng.integers(0, 256, (h // b + 1, w // b + 1), dtype=np.uint8).repeat(b, 0).repeat(b, 1)[:h, :w].copy()
            R = np.roll(L, -9, axis=1)
            R[rng.random((h, w)) < 0.02] ^= 0x11
            for epi in (True, False):
                ref, ocl, ocr = oracle.pair(L, R, of, osettings(5, 128, 0 if epi else 1, epi))
                supp, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=0 if epi else 1, epipolar=epi))
                assert (ncl, ncr) == (ocl, ocr), (w, h, epi)
                assert np.array_equal(supp, ref), (w, h, epi, len(supp), len(ref))
    with pytest.raises(g.GpcError) as e:
        g.Context(device=0, max_w=8208, max_h=32, max_batch=1).match_pair(np.zeros((32, 8208), np.uint8), np.zeros((32, 8208), np.uint8),
                                                                       g.sparsematch_settings())
    assert e.value.status == capi.GPC_E_DIMS


def test_contexts_on_two_host_threads(g, oracle):
    """One context per host thread (SURVEY.md 8b threading contract): two threads create their contexts, load
    different forests (two NVRTC builds racing for the cubin cache) and match concurrently; ctypes drops the GIL
    inside the library calls."""
    import threading
    from opengpc_b200.synth import synth_pair
    cases = [("tau", 1234, 512, 200), ("zero", 99, 768, 150)]
    want, got, errs = {}, {}, []
    for name, seed, w, h in cases:
        L, R = synth_pair(w, h, seed)
        want[name] = oracle.pair(L, R, oracle.read_forest(FORESTS[name]), osettings(5, 128, 0, True))[0]

    def work(name, seed, w, h):
        try:
            L, R = synth_pair(w, h, seed)
            with g.Context(device=0, max_w=w, max_h=h, max_batch=2) as c:
                c.set_forest(FORESTS[name])
                outs = []
                for _ in range(20):
                    outs.append(c.match_pair(L, R, g.sparsematch_settings())[0])
                got[name] = outs
        except Exception as e:                                  # surfaced in the main thread below
            errs.append((name, repr(e)))

    threads = [threading.Thread(target=work, args=c) for c in cases]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for name, _, _, _ in cases:
        assert all(np.array_equal(o, want[name]) for o in got[name]), name
