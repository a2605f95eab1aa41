#!/usr/bin/env python
"""Generate tests/golden/naive32.json from the reference's SSE=OFF build (oracle/_ref/libgpc_ref_naive.so): 32-test
forests, which gpcFilter[Tau]Naive accepts (filter.hpp:245-293) and which need the multi-word state path here.
Run in the build container (needs /root/reference):  python scripts/make_golden_naive32.py"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from oraclelib import Reference, digest  # noqa: E402
from opengpc_b200.synth import synth_pair  # noqa: E402


def pair_images(w, h, seed):
    """Odd seeds: the noisy synthetic pair (few matches under 32 random tests); even seeds: a shifted copy with sparse
    pixel flips (many matches, repeated states)."""
    L, R = synth_pair(w, h, seed)
    if seed % 2 == 0:
        rng = np.random.default_rng(1000 + seed)
        R = np.roll(L, -7, axis=1).copy()
        R[rng.random((h, w)) < 0.02] ^= 0x15
    return L, R


def main():
    ref = Reference(naive=True)
    cases = []
    for seed, (w, h), epi, vt, thr in ((21, (512, 160), True, 0, 5), (22, (320, 96), False, 1, 10), (23, (256, 64), True, 0, 5), (24, (512, 100), False, 2, 5)):
        rng = np.random.default_rng(seed)
        tests = rng.integers(-13, 14, (32, 5)).astype(np.int32)
        tests[:, 4] = rng.integers(-10, 11, 32) if seed % 2 else 0
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write("1\n0 l 32\n")
            for j, r in enumerate(tests):
                f.write(f"{j} {r[0]} {r[1]} {r[2]} {r[3]} {r[4]}\n")
            path = f.name
        L, R = pair_images(w, h, seed)
        supp, ncl, ncr, _ = ref.pair(L, R, path, thr=thr, disp_high=128, vt=vt, epipolar=epi)
        os.unlink(path)
        cases.append({"seed": seed, "w": w, "h": h, "epipolar": epi, "vt": vt, "thr": thr, "disp_high": 128, "tests": tests.tolist(),
                      "n_cand_l": ncl, "n_cand_r": ncr, "n_supports": len(supp), "digest": "%016x" % digest(supp)})
        print(cases[-1]["seed"], ncl, ncr, len(supp), cases[-1]["digest"])
    with open(os.path.join(ROOT, "tests", "golden", "naive32.json"), "w") as f:
        json.dump({"generator": "scripts/make_golden_naive32.py (reference SSE=OFF build)", "cases": cases}, f)


if __name__ == "__main__":
    main()
