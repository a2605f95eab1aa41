#!/usr/bin/env python
"""Workload for the compute-sanitizer runs (scripts/gpu_sanitize.sh): every kernel of the path on small inputs,
checked against the golden digests so that a tool-induced timing change that exposes a race shows up as a wrong
result as well as a tool report.

usage: python scripts/sanitize_cases.py [smoke|stress|hd]
  smoke   __graft_entry__.smoke()'s pair (512x128) through the fast row matcher, the general row kernel, the
          radix-sort matcher in epipolar and global mode, the hashtable matcher and the naive result mode
  stress  test_repeatability_stress: a batch of 4 Sintel-sized pairs, 3 repetitions, digest per pair
  hd      one 1920x1080 pair (512-thread row kernels), digest against tests/golden/golden.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import opengpc_b200 as g  # noqa: E402
from opengpc_b200.synth import synth_batch, synth_pair  # noqa: E402

FORESTS = {"tau": os.path.join(ROOT, "forests", "defaultTauForest.txt"), "zero": os.path.join(ROOT, "forests", "defaultZeroForest.txt")}


def digest(supp):
    vals = np.stack([supp["x"].astype(np.int64), supp["y"].astype(np.int64), supp["d"].astype(np.int64)], axis=1).reshape(-1) & 0xFFFFFFFF
    h = 1469598103934665603
    for v in vals.tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def golden(w, h, forest, seed, epipolar=True):
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        for p in json.load(f)["pairs"]:
            if (p["w"], p["h"], p["forest"], p["seed"], p["epipolar"], p["sparse"]) == (w, h, forest, seed, epipolar, False):
                return p
    raise KeyError((w, h, forest, seed))


def smoke():
    L, R = synth_pair(512, 128, 1234)
    with g.Context(device=0, max_w=512, max_h=128, max_batch=1) as ctx:
        ctx.set_forest(FORESTS["tau"])
        s = g.sparsematch_settings()
        base, ncl, ncr = ctx.match_pair(L, R, s)
        for m in (g.MATCHER_ROWS_GENERAL, g.MATCHER_SORT):
            ctx.set_matcher(m)
            other, _, _ = ctx.match_pair(L, R, s)
            assert np.array_equal(other, base), m
        ctx.set_matcher(g.MATCHER_AUTO)
        glob, _, _ = ctx.match_pair(L, R, g.make_settings(thr=10, disp_high=128, vt=1, epipolar=False))
        ht, _, _ = ctx.match_pair(L, R, g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=True))
        ctx.set_result_mode(True)
        ctx.set_forest(FORESTS["tau"])
        naive, _, _ = ctx.match_pair(L, R, s)
    print(f"smoke: {len(base)} supports ({ncl}/{ncr} candidates), global {len(glob)}, hashtable {len(ht)}, naive {len(naive)}")


def stress():
    imgs = synth_batch(1024, 436, 4, seed0=1234)
    want = [golden(1024, 436, "tau", 1234), golden(1024, 436, "tau", 1235)]
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=4) as ctx:
        ctx.set_forest(FORESTS["tau"])
        first = None
        for rep in range(3):
            supp, off, _ = ctx.match_batch(imgs, g.sparsematch_settings())
            d = [digest(supp[off[p]:off[p + 1]]) for p in range(4)]
            assert d[0] == want[0]["digest"] and d[1] == want[1]["digest"], (rep, d)
            first = first or d
            assert d == first, (rep, d, first)
    print("stress: 3 x 4 pairs, digests", first)


def hd():
    L, R = synth_pair(1920, 1080, 1234)
    want = golden(1920, 1080, "tau", 1234)
    with g.Context(device=0, max_w=1920, max_h=1080, max_batch=1) as ctx:
        ctx.set_forest(FORESTS["tau"])
        supp, _, _ = ctx.match_pair(L, R, g.sparsematch_settings())
    assert len(supp) == want["n_supports"] and digest(supp) == want["digest"], (len(supp), digest(supp))
    print(f"hd: {len(supp)} supports, digest {digest(supp)}")


if __name__ == "__main__":
    {"smoke": smoke, "stress": stress, "hd": hd}[sys.argv[1] if len(sys.argv) > 1 else "smoke"]()
