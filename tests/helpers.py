"""Shared helpers for the parity tests."""
import os
import tempfile

import numpy as np

from oraclelib import FOREST_TAU, FOREST_ZERO, ROOT

FOREST_DEEP = os.path.join(ROOT, "forests", "deepRandomForest16x12.txt")
FORESTS = {"tau": FOREST_TAU, "zero": FOREST_ZERO, "deep": FOREST_DEEP}


def write_forest(text_bytes):
    """Materialise a forest text held in a fixture; returns the path."""
    fd, path = tempfile.mkstemp(suffix=".txt")
    with os.fdopen(fd, "wb") as f:
        f.write(bytes(text_bytes))
    return path


def supp_to_i32(supp):
    return np.stack([supp["x"], supp["y"], supp["d"].astype(np.int32)], 1).astype(np.int32)


def make_pair(rec):
    from opengpc_b200.synth import sparsify, synth_pair
    L, R = synth_pair(rec["w"], rec["h"], rec["seed"])
    if rec.get("sparse"):
        L, R = sparsify(L), sparsify(R)
    return L, R
