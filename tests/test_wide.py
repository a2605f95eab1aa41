"""Forests of more than 32 tests (BASELINE configs[4] "extended mode") and 32-test forests in the SSE=OFF result mode.

Extended mode has no reference result (the reference drops every test after the 32nd, inference.hpp:426): PARITY UNPINNED,
checked against the scalar restatement in oraclelib.pair_wide only -- except for its first word, which IS the
reference's truncated state.  The naive 32-test case is pinned: oracle vs the reference's SSE=OFF build here, the GPU
against fixtures generated from that build (tests/golden/naive32.json, scripts/make_golden_naive32.py)."""
import json
import os

import numpy as np
import pytest

from helpers import FORESTS
from oraclelib import ROOT, Reference, digest, pair_wide, settings as osettings, wide_words

GOLD = os.path.join(ROOT, "tests", "golden", "naive32.json")


def _tests32(seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(-13, 14, (32, 5)).astype(np.int32)
    t[:, 4] = rng.integers(-10, 11, 32) if seed % 2 else 0
    return t


def _forest_file(path, tests):
    with open(path, "w") as f:
        f.write(f"1\n0 l {len(tests)}\n")
        for j, r in enumerate(tests):
            f.write(f"{j} {r[0]} {r[1]} {r[2]} {r[3]} {r[4]}\n")


def test_oracle_naive32_vs_reference(oracle, tmp_path):
    """gpcFilter[Tau]Naive accepts 32 tests (filter.hpp:245-293): the 31-bit word restatement equals the SSE=OFF build."""
    if not Reference.available(naive=True):
        pytest.skip("oracle/_ref naive build not present")
    from opengpc_b200.synth import synth_pair
    ref = Reference(naive=True)
    for seed in (1, 2):
        tests = _tests32(seed)
        path = str(tmp_path / f"f{seed}.txt")
        _forest_file(path, tests)
        L, R = synth_pair(320, 96, 40 + seed)
        for epi, vt in ((True, 0), (False, 1)):
            want = ref.pair(L, R, path, thr=5, disp_high=128, vt=vt, epipolar=epi)[0]
            got, _, _ = pair_wide(oracle, L, R, tests, osettings(5, 128, vt, epi), naive=True)
            assert np.array_equal(got, want), (seed, epi, len(got), len(want))


def test_oracle_wide_first_word_is_truncated_state(oracle):
    from opengpc_b200 import read_forest_tests
    from opengpc_b200.synth import synth_pair
    L, _ = synth_pair(256, 80, 5)
    tests = read_forest_tests(FORESTS["deep"])
    assert len(tests) == 192
    mk, words = wide_words(oracle, L, tests, 5, naive=False)
    _, _, mk32, st32 = oracle.stages(L, oracle.read_forest(FORESTS["deep"]), 5)
    assert words.shape[0] == 6 and np.array_equal(mk, mk32) and np.array_equal(words[0], st32)


@pytest.mark.gpu
def test_gpu_wide_forest_vs_restatement(oracle):
    import opengpc_b200 as g
    from opengpc_b200.synth import synth_pair
    tests = g.read_forest_tests(FORESTS["deep"])
    rng = np.random.default_rng(11)
    with g.Context(device=0, max_w=640, max_h=200, max_batch=1) as ctx:
        for w, h, nt, seed in ((640, 200, 192, 7), (256, 80, 70, 8), (320, 64, 33, 9), (320, 64, 32, 10), (256, 60, 5, 11)):
            sub = tests[:nt] if nt <= 192 else tests
            L, R = synth_pair(w, h, seed)
            ctx.set_wide_forest(sub)
            words = ctx.hash_wide(L, 5)
            mk, ow = wide_words(oracle, L, sub, 5, naive=False)
            assert words.shape[0] == ow.shape[0]
            for k in range(len(ow)):
                assert np.array_equal(words[k].reshape(-1)[mk] & 0x7fffffff, ow[k]), (nt, k)
                assert np.array_equal(np.flatnonzero(words[k].reshape(-1) >> 31), mk)
            if nt >= 32:                                            # word 0 = the reference's truncated 32-test state
                ctx.set_forest(g.make_forest([tuple(r[:4]) for r in sub[:32]], [int(r[4]) for r in sub[:32]]))
                st, mk2 = ctx.hash(L, 5)
                assert np.array_equal(words[0].reshape(-1)[mk2] & 0x7fffffff, st)
            for epi, vt, dh in ((True, 0, 128), (False, 1, 64)):
                want, ocl, ocr = pair_wide(oracle, L, R, sub, osettings(5, dh, vt, epi))
                got, ncl, ncr = ctx.match_pair_wide(L, R, g.make_settings(thr=5, disp_high=dh, vt=vt, epipolar=epi))
                assert (ncl, ncr) == (ocl, ocr)
                assert np.array_equal(got, want), (nt, epi, len(got), len(want))
        # duplicated rows / flat regions: many equal tuples, tail rules on the last key
        L = np.tile(rng.integers(0, 256, (1, 256), dtype=np.uint8), (60, 1))
        L[20:40] = rng.integers(0, 256, (20, 256), dtype=np.uint8)
        R = np.roll(L, -3, axis=1)
        ctx.set_wide_forest(tests[:100])
        for epi, vt in ((True, 0), (False, 2)):
            want, _, _ = pair_wide(oracle, L, R, tests[:100], osettings(1, 128, vt, epi))
            got, _, _ = ctx.match_pair_wide(L, R, g.make_settings(thr=1, disp_high=128, vt=vt, epipolar=epi))
            assert np.array_equal(got, want), (epi, len(got), len(want))


@pytest.mark.gpu
def test_gpu_naive32_vs_golden(oracle):
    """32-test forests in GPC_RESULTS_NAIVE: fixtures from the reference's SSE=OFF build + the restatement."""
    import opengpc_b200 as g
    from opengpc_b200.synth import synth_pair
    with open(GOLD) as f:
        gold = json.load(f)
    with g.Context(device=0, max_w=512, max_h=160, max_batch=1) as ctx:
        ctx.set_result_mode(True)
        with pytest.raises(g.GpcError):
            ctx.set_forest(g.make_forest([tuple(r[:4]) for r in _tests32(1)], [int(r[4]) for r in _tests32(1)]))   # one hash word cannot hold it
        for rec in gold["cases"]:
            tests = np.array(rec["tests"], np.int32)
            L, R = synth_pair(rec["w"], rec["h"], rec["seed"])
            if rec["seed"] % 2 == 0:                               # scripts/make_golden_naive32.py: pair_images
                rng = np.random.default_rng(1000 + rec["seed"])
                R = np.roll(L, -7, axis=1).copy()
                R[rng.random((rec["h"], rec["w"])) < 0.02] ^= 0x15
            ctx.set_wide_forest(tests)
            s = g.make_settings(thr=rec["thr"], disp_high=rec["disp_high"], vt=rec["vt"], epipolar=rec["epipolar"])
            got, ncl, ncr = ctx.match_pair_wide(L, R, s)
            assert (ncl, ncr, len(got)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), rec["seed"]
            assert "%016x" % digest(got) == rec["digest"], rec["seed"]
            want, _, _ = pair_wide(oracle, L, R, tests, osettings(rec["thr"], rec["disp_high"], rec["vt"], rec["epipolar"]), naive=True)
            assert np.array_equal(got, want)
