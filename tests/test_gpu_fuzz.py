"""Randomised parity campaign (pytest -m gpu): random shapes, textures, thresholds, forests and settings
against the oracle, every case through the whole path in both matching modes, every third also through the
hashtable matcher and every fourth in the result mode of the reference's SSE=OFF build.  GPC_FUZZ_CASES scales it up."""
import os

import numpy as np
import pytest

from oraclelib import settings as osettings

pytestmark = pytest.mark.gpu


def _image(rng, w, h, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:                                                   # blocky texture, like the synthetic generator
        b = int(rng.integers(2, 7))
        return rng.integers(0, 256, (h // b + 1, w // b + 1), dtype=np.uint8).repeat(b, 0).repeat(b, 1)[:h, :w].copy()
    if kind == 2:                                                   # mostly flat with isolated features
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        n = int(rng.integers(1, 40))
        img[rng.integers(0, h, n), rng.integers(0, w, n)] = rng.integers(0, 256, n, dtype=np.uint8)
        return img
    if kind == 3:                                                   # values around the 127/128 wrap of the signed tau arithmetic
        return rng.integers(118, 140, (h, w), dtype=np.uint8)
    x = np.arange(w)[None, :] * int(rng.integers(1, 9))
    y = np.arange(h)[:, None] * int(rng.integers(0, 5))
    return ((x + y) % 256).astype(np.uint8) * np.ones((h, 1), np.uint8)   # periodic: many equal states


def test_fuzz_vs_oracle(oracle):
    import opengpc_b200 as g
    n_cases = int(os.environ.get("GPC_FUZZ_CASES", "60"))
    rng = np.random.default_rng(int(os.environ.get("GPC_FUZZ_SEED", "2026")))
    with g.Context(device=0, max_w=1408, max_h=160, max_batch=1) as ctx, \
            g.Context(device=0, max_w=1408, max_h=160, max_batch=1) as nctx:
        nctx.set_result_mode(True)                                  # the reference's SSE=OFF build
        for it in range(n_cases):
            w = 16 * int(rng.integers(1, 89))
            h = int(rng.integers(20, 161))
            L = _image(rng, w, h, int(rng.integers(0, 5)))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                R = np.roll(L, -int(rng.integers(0, 20)), axis=1)
            elif mode == 1:
                R = np.roll(L, (int(rng.integers(-2, 3)), -int(rng.integers(0, 9))), axis=(0, 1))
            elif mode == 2:
                R = _image(rng, w, h, int(rng.integers(0, 5)))
            else:
                R = L.copy()
                R[rng.random((h, w)) < 0.05] ^= 0x21
            nt = int(rng.choice([1, 2, 7, 8, 9, 10, 16, 17, 24, 25, 26, 30, 31, 32]))
            tests = [tuple(int(v) for v in rng.integers(-13, 14, 4)) for _ in range(nt)]
            tkind = int(rng.integers(0, 3))
            taus = [0] * nt if tkind == 0 else [int(v) for v in (rng.integers(-128, 128, nt) if tkind == 1 else rng.integers(-12, 13, nt))]
            thr = int(rng.choice([0, 1, 5, 10, 40, 181, 182, 255]))
            dh = int(rng.choice([0, 5, 128, 4000]))
            epi = bool(rng.integers(0, 2))
            vt = 0 if epi else int(rng.choice([0, 1, 3, 500]))
            of = oracle.make_forest(tests, taus)
            ctx.set_forest(g.make_forest(tests, taus))
            ref, ocl, ocr = oracle.pair(L, R, of, osettings(thr, dh, vt, epi))
            supp, ncl, ncr = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi))
            info = (it, w, h, nt, tkind, thr, dh, epi, vt, mode)
            assert (ncl, ncr) == (ocl, ocr), info
            assert np.array_equal(supp, ref), info + (len(supp), len(ref))
            if epi and it % 2 == 0:                                 # every row through the general row kernel
                ctx.set_matcher(g.MATCHER_ROWS_GENERAL)
                try:
                    supp_g, _, _ = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi))
                finally:
                    ctx.set_matcher(g.MATCHER_AUTO)
                assert np.array_equal(supp_g, ref), info + ("general rows", len(supp_g), len(ref))
            if it % 3 == 0:                                         # the same case through useHashtable(true)
                ref_h = oracle.pair_hashtable(L, R, of, osettings(thr, dh, vt, epi))
                supp_h, _, _ = ctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi, use_hashtable=True))
                assert np.array_equal(supp_h, ref_h), info + ("hashtable", len(supp_h), len(ref_h))
            if it % 4 == 1 and nt <= 31:                            # the same case in the naive result mode
                nctx.set_forest(g.make_forest(tests, taus))
                ref_n, ocl_n, ocr_n = oracle.pair_naive(L, R, of, osettings(thr, dh, vt, epi))
                supp_n, ncl_n, ncr_n = nctx.match_pair(L, R, g.make_settings(thr=thr, disp_high=dh, vt=vt, epipolar=epi))
                assert (ncl_n, ncr_n) == (ocl_n, ocr_n), info + ("naive",)
                assert np.array_equal(supp_n, ref_n), info + ("naive", len(supp_n), len(ref_n))
