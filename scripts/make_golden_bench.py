#!/usr/bin/env python
"""Generate tests/golden/bench_pairs.json from the UNMODIFIED reference compiled into oracle/_ref:
count + FNV digest of the ordered support list of the synthetic pairs bench.py times (seeds 1234..1241,
sparsematch settings) for every forest / shape of BASELINE.json's configs.  bench.py checks the batch it
timed against these values in plain Python (it never imports oracle/ in its product arm).

Run in the build container (needs /root/reference):  python scripts/make_golden_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from oraclelib import Reference, digest  # noqa: E402
from opengpc_b200.synth import sparsify, synth_pair  # noqa: E402

FORESTS = {"tau": "defaultTauForest.txt", "zero": "defaultZeroForest.txt", "deep": "deepRandomForest16x12.txt"}
CASES = [("1024x436", "tau", False), ("1024x436", "zero", False), ("1024x436", "zero", True), ("1920x1080", "tau", False), ("1920x1080", "deep", False)]


def main():
    ref = Reference()
    out = {"settings": {"thr": 5, "vt": 0, "disp_high": 128, "epipolar": True}, "seeds": list(range(1234, 1242)), "cases": {}}
    for shape, forest, sparse in CASES:
        w, h = (int(v) for v in shape.split("x"))
        recs = []
        for seed in out["seeds"]:
            L, R = synth_pair(w, h, seed)
            if sparse:
                L, R = sparsify(L), sparsify(R)
            supp, ncl, ncr, _ = ref.pair(L, R, os.path.join(ROOT, "forests", FORESTS[forest]))
            recs.append({"seed": seed, "n_cand_l": ncl, "n_cand_r": ncr, "n_supports": len(supp), "digest": "%016x" % digest(supp)})
            print(shape, forest, sparse, recs[-1], flush=True)
        out["cases"][f"{shape}/{forest}" + ("/sparse" if sparse else "")] = recs
    with open(os.path.join(ROOT, "tests", "golden", "bench_pairs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
