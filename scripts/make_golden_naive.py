#!/usr/bin/env python
"""Golden values for the reference's SSE=OFF result mode (the *Naive functions of filter.hpp), from the UNMODIFIED
reference compiled WITHOUT -D_INTRINSICS_SSE (oracle/_ref/libgpc_ref_naive.so).  Needs /root/reference.
Writes tests/golden/naive.json."""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oraclelib import FOREST_TAU, FOREST_ZERO, Reference, digest   # noqa: E402
from opengpc_b200.synth import sparsify, synth_pair   # noqa: E402


def main():
    ref = Reference(naive=True)
    out = {"build": "g++ -std=c++11 -O3 -funroll-loops (no -D_INTRINSICS_SSE), oracle/Makefile", "pairs": [], "stages": []}
    for (w, h, seed, forest, ep, vt, dh, thr, sparse) in [(1024, 436, 1234, "tau", True, 0, 128, 5, False), (1024, 436, 1234, "zero", True, 0, 128, 5, False),
                                                          (1024, 436, 1235, "tau", False, 1, 128, 10, False), (512, 200, 7, "zero", False, 0, 64, 5, False),
                                                          (640, 120, 9, "tau", True, 0, 128, 5, True), (256, 64, 3, "tau", False, 100, 1000, 0, False)]:
        L, R = synth_pair(w, h, seed)
        if sparse:
            L, R = sparsify(L), sparsify(R)
        fp = FOREST_TAU if forest == "tau" else FOREST_ZERO
        supp, ncl, ncr, _ = ref.pair(L, R, fp, thr=thr, disp_high=dh, vt=vt, epipolar=ep)
        ht = ref.pair_hashtable(L, R, fp, thr=thr, disp_high=dh, vt=vt, epipolar=ep)
        rec = {"w": w, "h": h, "seed": seed, "forest": forest, "epipolar": ep, "vt": vt, "disp_high": dh, "thr": thr, "sparse": sparse,
               "n_cand_l": ncl, "n_cand_r": ncr, "n_supports": int(len(supp)), "digest": "%016x" % digest(supp),
               "n_supports_hashtable": int(len(ht)), "digest_hashtable": "%016x" % digest(ht)}
        print(rec)
        out["pairs"].append(rec)
    for (w, h, seed, forest, thr) in [(512, 128, 1234, "tau", 5), (256, 64, 3, "zero", 10)]:
        L, _ = synth_pair(w, h, seed)
        fp = FOREST_TAU if forest == "tau" else FOREST_ZERO
        sm, gr, mk = ref.preprocess(L, thr)
        st = ref.hash(L, thr, fp)
        gr = gr.copy()
        gr[0, :] = 0; gr[1, 0] = 0; gr[h - 1, :] = 0; gr[h - 2, w - 1] = 0     # positions the reference leaves unwritten / reads past the image for
        out["stages"].append({"w": w, "h": h, "seed": seed, "forest": forest, "thr": thr, "n_mask": int(len(mk)),
                              "crc_smooth": zlib.crc32(sm.tobytes()), "crc_grad_written": zlib.crc32(gr.tobytes()),
                              "crc_mask": zlib.crc32(mk.astype("<i4").tobytes()), "crc_states": zlib.crc32(st.astype("<u4").tobytes())})
        print(out["stages"][-1])
    with open(os.path.join(ROOT, "tests", "golden", "naive.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
