"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the golden
fixtures generated from the compiled reference and against the oracle on fresh seeded inputs."""
import numpy as np
import pytest

from helpers import FORESTS, make_pair, supp_to_i32, write_forest
from oraclelib import settings as osettings

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import opengpc_b200
    return opengpc_b200


@pytest.fixture(scope="module")
def ctx_small(g):
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=4) as c:
        yield c


def gsettings(g, thr, dh, vt, epi):
    return g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi)


def test_small_cases_all_stages(g, ctx_small, small_cases):
    """smooth / grad / mask / per-candidate states / ordered supports of every small fixture."""
    for name, c in small_cases.items():
        thr, epi, vt, dh = (int(v) for v in c["cfg"])
        ctx_small.set_forest(write_forest(c["forest"]))
        for side in "LR":
            sm, gr, mk = ctx_small.preprocess(c[side], thr)
            assert np.array_equal(sm, c["smooth" + side]), (name, side, "smooth")
            assert np.array_equal(gr[:, 2:], c["grad" + side][:, 2:]), (name, side, "grad")
            assert np.array_equal(mk, c["mask" + side]), (name, side, "mask")
            st, mk2 = ctx_small.hash(c[side], thr)
            assert np.array_equal(mk2, c["mask" + side]), (name, side, "mask(hash)")
            assert np.array_equal(st, c["states" + side]), (name, side, "states")
        if not epi:
            continue   # global matching: see test_global_mode
        supp, ncl, ncr = ctx_small.match_pair(c["L"], c["R"], gsettings(g, thr, dh, vt, epi))
        assert (ncl, ncr) == (len(c["maskL"]), len(c["maskR"]))
        assert np.array_equal(supp_to_i32(supp), c["supp"].reshape(-1, 3)), (name, "supports")


def _golden_records(golden, pred):
    return [r for r in golden["pairs"] if pred(r)]


def test_golden_sintel(g, ctx_small, golden, oracle):
    """configs[0] and configs[1]: 1024x436, tau and zero forests, dense and low-texture inputs."""
    recs = _golden_records(golden, lambda r: (r["w"], r["h"]) == (1024, 436) and r["epipolar"])
    assert len(recs) >= 5
    for rec in recs:
        L, R = make_pair(rec)
        ctx_small.set_forest(FORESTS[rec["forest"]])
        supp, ncl, ncr = ctx_small.match_pair(L, R, gsettings(g, 5, rec["disp_high"], rec["vt"], True))
        assert (ncl, ncr, len(supp)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), rec
        assert "%016x" % oracle.digest(supp) == rec["digest"], rec


@pytest.mark.parametrize("shape", [(1920, 1080), (3840, 2160), (960, 540), (480, 270)])
def test_golden_large(g, golden, oracle, shape):
    """configs[2..4] shapes at full size: counts + ordered-list digest."""
    recs = _golden_records(golden, lambda r: (r["w"], r["h"]) == shape and r["epipolar"])
    assert recs
    with g.Context(device=0, max_w=shape[0], max_h=shape[1], max_batch=1) as ctx:
        for rec in recs:
            L, R = make_pair(rec)
            ctx.set_forest(FORESTS[rec["forest"]])
            supp, ncl, ncr = ctx.match_pair(L, R, gsettings(g, 5, rec["disp_high"], rec["vt"], True))
            assert (ncl, ncr, len(supp)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), rec
            assert "%016x" % oracle.digest(supp) == rec["digest"], rec


def test_random_vs_oracle(g, ctx_small, oracle):
    """Fresh random images, thresholds, forests (<= 32 tests, tau in [-128,127]) against the oracle."""
    rng = np.random.default_rng(11)
    for it in range(12):
        w = int(rng.integers(2, 40)) * 16
        h = int(rng.integers(27, 120))
        kind = it % 4
        if kind == 0:
            L = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            L = (rng.integers(0, 256, (h // 4 + 1, w // 4 + 1), dtype=np.uint8).repeat(4, 0).repeat(4, 1)[:h, :w]).copy()
        elif kind == 2:
            L = np.full((h, w), 90, np.uint8)
            L[::7, ::5] = 200                     # sparse texture: rows with 0-2 candidates
        else:
            L = rng.integers(100, 140, (h, w), dtype=np.uint8)
        R = np.roll(L, -int(rng.integers(0, 6)), axis=1)
        thr = int(rng.choice([0, 2, 5, 10, 30, 181, 182, 255]))
        nt = int(rng.integers(1, 33))
        tests = [tuple(int(v) for v in rng.integers(-13, 14, 4)) for _ in range(nt)]
        taus = [int(v) for v in (rng.integers(-128, 128, nt) if it % 3 else np.zeros(nt, int))]
        dh = int(rng.choice([0, 3, 128]))
        of = oracle.make_forest(tests, taus)
        ctx_small.set_forest(g.make_forest(tests, taus))
        for img in (L, R):
            osm, ogr, omk, ost = oracle.stages(img, of, thr)
            sm, gr, mk = ctx_small.preprocess(img, thr)
            st, _ = ctx_small.hash(img, thr)
            assert np.array_equal(sm, osm) and np.array_equal(gr, ogr) and np.array_equal(mk, omk), (it, w, h, thr)
            assert np.array_equal(st, ost), (it, w, h, nt)
        ref, ocl, ocr = oracle.pair(L, R, of, osettings(thr, dh, 0, True))
        supp, ncl, ncr = ctx_small.match_pair(L, R, gsettings(g, thr, dh, 0, True))
        assert (ncl, ncr) == (ocl, ocr)
        assert np.array_equal(supp, ref), (it, len(supp), len(ref))


def _kat_images(src, tar, w=64, h=28, row=14, extra_r_row=None):
    """Hash images whose candidate row `row` carries the KAT keys (x = 13 + index)."""
    hl = np.zeros((h, w), np.uint32)
    hr = np.zeros((h, w), np.uint32)
    for i, k in enumerate(src):
        hl[row, 13 + i] = 0x80000000 | k
    for j, k in enumerate(tar):
        hr[row, 13 + j] = 0x80000000 | k
    if extra_r_row is not None:
        hr[extra_r_row, 20] = 0x80000000 | 0x123
    return hl, hr


def test_matcher_kats(g, ctx_small, golden, oracle):
    """findCorrespondences known answers incl. the tail rules (last right key never matches;
    two equal keys at the tail match with the first)."""
    s = gsettings(g, 5, 1000, 0, True)
    for kat in golden["kats"]:
        src, tar = kat["src"], kat["tar"]
        hl, hr = _kat_images(src, tar)
        supp = ctx_small.match_hash_images(hl, hr, s)
        want = [(13 + i, 14, float(i - j)) for i, j in kat["pairs"]]
        got = [(int(a), int(b), float(c)) for a, b, c in zip(supp["x"], supp["y"], supp["d"])]
        assert got == want, (kat, got)
        # same keys one row up with another right candidate below: the tail rules no longer apply
        hl2, hr2 = _kat_images(src, tar, row=13, extra_r_row=14)
        supp2 = ctx_small.match_hash_images(hl2, hr2, s)
        keys_s = np.array(src, np.uint64) | (np.uint64(13) << np.uint64(32))
        keys_t = np.concatenate([np.array(tar, np.uint64) | (np.uint64(13) << np.uint64(32)),
                                 np.array([0x123 | (14 << 32)], np.uint64)])
        pairs = oracle.find_correspondences(keys_s, keys_t)
        want2 = [(13 + i, 13, float(i - j)) for i, j in pairs.tolist()]
        got2 = [(int(a), int(b), float(c)) for a, b, c in zip(supp2["x"], supp2["y"], supp2["d"])]
        assert got2 == want2, (kat, got2, want2)


def test_batch_equals_single(g, ctx_small, oracle):
    from opengpc_b200.synth import synth_batch
    imgs = synth_batch(256, 96, 4, seed0=50)
    imgs[2] = 77                      # a pair without candidates in the middle of the batch
    ctx_small.set_forest(FORESTS["tau"])
    s = g.sparsematch_settings()
    supp, offsets, ncand = ctx_small.match_batch(imgs, s)
    of = oracle.read_forest(FORESTS["tau"])
    for p in range(4):
        ref, ocl, ocr = oracle.pair(imgs[p, 0], imgs[p, 1], of, osettings())
        got = supp[offsets[p]:offsets[p + 1]]
        assert (ncand[p, 0], ncand[p, 1]) == (ocl, ocr)
        assert np.array_equal(got, ref), p
    assert offsets[3] == offsets[2]


def test_device_resident_batch(g, oracle):
    """gpc_match_batch_device: device pointers in, device supports out, on the caller's stream."""
    import torch
    from opengpc_b200.synth import synth_batch
    imgs = synth_batch(512, 128, 3, seed0=9)
    of = oracle.read_forest(FORESTS["zero"])
    cap = 20000
    with g.Context(device=0, max_w=512, max_h=128, max_batch=3) as ctx:
        ctx.set_forest(FORESTS["zero"])
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            d_img = torch.from_numpy(imgs).cuda()
            d_out = torch.zeros((3, cap, 3), dtype=torch.int32, device="cuda")
            d_n = torch.zeros(3, dtype=torch.int32, device="cuda")
            d_nc = torch.zeros((3, 2), dtype=torch.int32, device="cuda")
            ctx.set_stream(stream.cuda_stream)
            ctx.match_batch_device(d_img.data_ptr(), 3, 512, 128, g.sparsematch_settings(), d_out.data_ptr(), cap,
                                   d_n.data_ptr(), d_nc.data_ptr())
        stream.synchronize()
        n = d_n.cpu().numpy()
        out = d_out.cpu().numpy()
        nc = d_nc.cpu().numpy()
    for p in range(3):
        ref, ocl, ocr = oracle.pair(imgs[p, 0], imgs[p, 1], of, osettings())
        assert n[p] == len(ref) and (nc[p, 0], nc[p, 1]) == (ocl, ocr)
        got = out[p, :n[p]].copy().view(g.SUPPORT_DTYPE).reshape(-1)
        assert np.array_equal(got, ref)


def test_error_codes(g, ctx_small):
    from opengpc_b200 import capi
    s = g.sparsematch_settings()
    img = np.zeros((64, 72), np.uint8)            # width not a multiple of 16
    with pytest.raises(g.GpcError) as e:
        ctx_small.match_pair(img, img, s)
    assert e.value.status == capi.GPC_E_WIDTH16
    img = np.zeros((64, 2048), np.uint8)          # exceeds the context capacity
    with pytest.raises(g.GpcError) as e:
        ctx_small.match_pair(img, img, s)
    assert e.value.status == capi.GPC_E_DIMS
    img = np.zeros((64, 64), np.uint8)
    with pytest.raises(g.GpcError) as e:
        ctx_small.match_pair(img, img, s, cap=0) if False else ctx_small.set_forest(g.make_forest([(14, 0, 0, 0)], [0]))
    assert e.value.status == capi.GPC_E_FOREST


def test_repeatability_stress(g, ctx_small, oracle):
    """The row matcher uses shared-memory atomics: the same batch, run many times, must give the
    same ordered result every time (guards against timing-dependent races)."""
    from opengpc_b200.synth import synth_batch
    imgs = synth_batch(1024, 436, 4, seed0=1234)
    ctx_small.set_forest(FORESTS["tau"])
    s = g.sparsematch_settings()
    of = oracle.read_forest(FORESTS["tau"])
    refs = [oracle.pair(imgs[p, 0], imgs[p, 1], of, osettings())[0] for p in range(4)]
    for rep in range(25):
        supp, offsets, _ = ctx_small.match_batch(imgs, s)
        for p in range(4):
            assert np.array_equal(supp[offsets[p]:offsets[p + 1]], refs[p]), (rep, p)


def test_gpu_vs_compiled_reference(g, reference):
    """The CUDA path against the UNMODIFIED reference itself (oracle/_ref/libgpc_ref.so travels to the GPU box prebuilt),
    not only against the restatement: supports, candidate counts, order -- both shipped forests, both matching modes."""
    from opengpc_b200.synth import sparsify, synth_pair
    cases = [(640, 200, 31, False), (512, 131, 32, False), (1024, 436, 1234, True)]
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=1) as ctx:
        for w, h, seed, sparse in cases:
            L, R = synth_pair(w, h, seed)
            if sparse:
                L, R = sparsify(L), sparsify(R)
            for forest in ("tau", "zero"):
                ctx.set_forest(FORESTS[forest])
                for epi, vt, thr in ((True, 0, 5), (False, 1, 10)):
                    want, rcl, rcr, _ = reference.pair(L, R, FORESTS[forest], thr=thr, disp_high=128, vt=vt, epipolar=epi)
                    got, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=128, vt=vt, epipolar=epi))
                    assert (ncl, ncr) == (rcl, rcr), (w, h, forest, epi)
                    assert np.array_equal(got, want), (w, h, forest, epi, len(got), len(want))
