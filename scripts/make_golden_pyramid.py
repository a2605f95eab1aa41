#!/usr/bin/env python
"""Golden values for the multi-level configuration (BASELINE.json configs[3]), from the
UNMODIFIED reference (oracle/_ref) run once per level on identically down-sampled CPU images.
Needs /root/reference (to build oracle/_ref).  Writes tests/golden/pyramid.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oraclelib import FOREST_TAU, Reference, digest   # noqa: E402
from opengpc_b200.synth import downsample2x, synth_pair   # noqa: E402


def main():
    ref = Reference()
    out = {"definition": "level l+1 = 2x2 floor-mean of the raw level-l images; reference path per level; disp_high halved per level",
           "cases": []}
    for (w, h, seed, levels) in [(3840, 2160, 1234, 4), (1024, 436, 1234, 3), (512, 250, 5, 2)]:
        L, R = synth_pair(w, h, seed)
        dh = 128
        recs = []
        for l in range(levels):
            supp, ncl, ncr, _ = ref.pair(L, R, FOREST_TAU, thr=5, disp_high=dh, vt=0, epipolar=True)
            recs.append({"level": l, "w": int(L.shape[1]), "h": int(L.shape[0]), "disp_high": dh, "n_cand_l": ncl, "n_cand_r": ncr,
                         "n_supports": int(len(supp)), "digest": "%016x" % digest(supp)})
            print(w, h, recs[-1])
            L, R = downsample2x(L), downsample2x(R)
            dh //= 2
        out["cases"].append({"w": w, "h": h, "seed": seed, "forest": "tau", "levels": recs})
    with open(os.path.join(ROOT, "tests", "golden", "pyramid.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
