#!/usr/bin/env python
"""Build tuning variants of libgpc_b200.so (kernel A tile geometry) into opengpc_b200/variants/.
usage: python scripts/build_variants.py W,H,T [W,H,T ...]     e.g. 256,64,512 128,32,256
Run a variant with GPC_B200_LIB=opengpc_b200/variants/libgpc_W_H_T.so python bench.py ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengpc_b200.build import CSRC, NVCC_FLAGS, SOURCES, _nvcc   # noqa: E402

out_dir = os.path.join(ROOT, "opengpc_b200", "variants")
os.makedirs(out_dir, exist_ok=True)


def build(spec):
    w, h, t = spec.split(",")[:3]
    extra = spec.split(",")[3:]
    import re
    lib = os.path.join(out_dir, "libgpc_" + re.sub(r"[^A-Za-z0-9_]", "", spec.replace(",", "_")) + ".so")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-DGPC_TILE_W={w}", f"-DGPC_TILE_H={h}", f"-DGPC_THREADS_A={t}"] + \
          [f"-D{e}" for e in extra] + ["-I", os.path.join(ROOT, "include"), "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return spec, r.returncode, r.stderr[-800:]


with ThreadPoolExecutor(4) as ex:
    for spec, rc, err in ex.map(build, sys.argv[1:]):
        print(spec, "ok" if rc == 0 else "FAILED\n" + err)
