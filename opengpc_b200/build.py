"""Build libgpc_b200.so (hand-written sm_100a CUDA + the C ABI of include/gpc_b200.h) in-tree.

nvcc cross-compiles without a GPU; the built library is git-ignored but travels to the GPU box
with the working-tree snapshot.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libgpc_b200.so")
SOURCES = ["gpc_capi.cu", "smooth_sobel.cu", "hash_tiles.cu", "match_rows.cu", "match_global.cu", "pyramid.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall", "--shared", "-cudart", "static"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "gpc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False, extra_flags=()):
    """Compile every CUDA source for sm_100a into opengpc_b200/libgpc_b200.so."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(" ".join(cmd))
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgpc_b200.so")
    return LIB


if __name__ == "__main__":
    build_native(force=True, verbose=True, extra_flags=("-Xptxas", "-v"))
