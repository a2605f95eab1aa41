#!/usr/bin/env python
"""Golden values for InferenceSettings::useHashtable(true) (inference.hpp:204-225, hashmatch.hpp:48-272), from the
UNMODIFIED reference (oracle/_ref): key-list known answers of ndb::Hashmatch driven as depthPriorFast drives it, and
whole-pair support digests.  Needs /root/reference (to build oracle/_ref).  Writes tests/golden/hashtable.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oraclelib import FOREST_TAU, FOREST_ZERO, Reference, digest   # noqa: E402
from opengpc_b200.synth import sparsify, synth_pair   # noqa: E402

B = 214673   # bucket count, inference.hpp:212


def main():
    ref = Reference()
    rng = np.random.default_rng(2024)
    kats = []
    hand = [([1, 2, 3], [1, 2, 3]), ([1, 2, 3], [1, 2, 3, 9]), ([5, 5], [5]), ([5], [5, 5]), ([5], [5, 5, 5]), ([7], [7]),
            ([], [1]), ([1], []), ([10, 10, 10], [10, 11]),                                   # the header's 10s10s10s10t11t case
            ([1, 1 + B, 1 + 2 * B], [1 + B, 1, 1 + 3 * B]),                                    # one bucket, distinct keys
            ([3] * 12, [3]), ([k * B + 4 for k in range(12)], [k * B + 4 for k in range(12)]),  # more than 10 per bucket
            ([4, 4 + B], [4 + B, 4 + B, 4]), ([8, 8 + B, 8 + 2 * B], [8, 8 + B, 8 + 2 * B])]
    for s, t in hand:
        kats.append((np.array(s, np.uint64), np.array(t, np.uint64)))
    for case in range(40):
        ns, nt = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        kind = case % 4
        if kind == 0:
            pool = rng.integers(0, 2 ** 40, size=12, dtype=np.uint64)
        elif kind == 1:
            pool = rng.integers(0, 16, size=10, dtype=np.uint64) * np.uint64(B) + np.uint64(7)
        elif kind == 2:
            pool = rng.integers(0, 2 ** 63, size=200, dtype=np.uint64)
        else:
            pool = np.concatenate([rng.integers(0, 4, size=4, dtype=np.uint64) * np.uint64(B),
                                   rng.integers(0, 2 ** 32, size=10, dtype=np.uint64)])
        kats.append((rng.choice(pool, ns), rng.choice(pool, nt)))
    out = {"buckets": B, "kats": [], "pairs": []}
    for s, t in kats:
        p = ref.hashmatch(s, t)
        out["kats"].append({"src": [int(v) for v in s], "tar": [int(v) for v in t], "pairs": p.tolist()})
    for (w, h, seed, forest, ep, vt, dh, thr, sparse) in [(1024, 436, 1234, "tau", True, 0, 128, 5, False), (1024, 436, 1234, "zero", True, 0, 128, 5, False),
                                                          (1024, 436, 1235, "tau", False, 1, 128, 10, False), (512, 200, 7, "zero", False, 0, 64, 5, False),
                                                          (640, 120, 9, "tau", True, 0, 128, 5, True), (256, 64, 3, "tau", False, 100, 1000, 0, False)]:
        L, R = synth_pair(w, h, seed)
        if sparse:
            L, R = sparsify(L), sparsify(R)
        supp = ref.pair_hashtable(L, R, FOREST_TAU if forest == "tau" else FOREST_ZERO, thr=thr, disp_high=dh, vt=vt, epipolar=ep)
        rec = {"w": w, "h": h, "seed": seed, "forest": forest, "epipolar": ep, "vt": vt, "disp_high": dh, "thr": thr, "sparse": sparse,
               "n_supports": int(len(supp)), "digest": "%016x" % digest(supp)}
        print(rec)
        out["pairs"].append(rec)
    with open(os.path.join(ROOT, "tests", "golden", "hashtable.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
