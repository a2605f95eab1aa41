#!/usr/bin/env python
"""profiles/rNN_sass_excerpt.txt: per hot kernel of libgpc_b200.so the instruction count and the mnemonics that show what
the hardware runs (TMA tensor copies + mbarriers, DPX clamp, byte-SIMD differences, dot products, the absence of wide
multiplies in the inner loops, shared-memory atomics).  The forest-specialised kernel A2 is rebuilt offline the way
jit.cu builds it (scripts/jit_sass.py).   usage: python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opengpc_b200", "libgpc_b200.so")
KERNELS = ["smooth_sobel_tma_kernelILb0", "hash_tiles_kernelILi0", "hash_tiles_kernelILi1", "match_rows_fast_kernelILi1ELi256",
           "match_rows_tail_kernelILi256ELi128", "emit_supports_kernel"]
WATCH = ["UTMALDG", "SYNCS", "VIADDMNMX", "VABSDIFF4", "IDP.4A", "LEA.HI", "IMAD.HI", "IMAD.WIDE", "ATOMS.OR", "ATOMS.CAS", "ATOMS.ADD",
         "LOP3", "PRMT", "LDS", "LDG", "STG", "REDG"]
INST = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")


def excerpt(title, sass):
    lines = [l for l in sass.splitlines() if INST.match(l)]
    print(f"## {title}: {len(lines)} instructions")
    for w in WATCH:
        hits = [l for l in lines if INST.match(l).group(1).startswith(w)]
        if hits:
            print(f"   {w:10s} x{len(hits):<4d} e.g. {hits[0].strip()[:110]}")
    print()


def main():
    ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]
    print(f"# SASS excerpt of opengpc_b200/libgpc_b200.so (cuobjdump -sass, {ver.strip()}, sm_100a)")
    print("# Per kernel: instruction count and the mnemonics that show what the hardware runs.  No IMAD.HI and (outside address")
    print("# arithmetic) no IMAD.WIDE in the loops of kernels A1 / A2: see DESIGN.md section 3 for what they cost.\n")
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    for k in KERNELS:
        for f in funcs:
            if k in f:
                sass = subprocess.run(["cuobjdump", "-sass", "-fun", f, LIB], capture_output=True, text=True).stdout
                excerpt(f, sass)
                break
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import jit_sass
    for forest in ("defaultZeroForest.txt", "defaultTauForest.txt"):
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "gpc_jit_forest.h"), "w") as fh:
                fh.write(jit_sass.header(jit_sass.read_tests(os.path.join(ROOT, "forests", forest))))
            cubin = os.path.join(tmp, "a2.cubin")
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-lineinfo",
                            "-DGPC_JIT_HEADER=\"gpc_jit_forest.h\"", "-I", tmp, "-I", jit_sass.CSRC, "-o", cubin,
                            os.path.join(jit_sass.CSRC, "hash_tiles.cu")], check=True, capture_output=True)
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", "gpc_hash_tiles_jit", cubin], capture_output=True, text=True).stdout
        excerpt(f"gpc_hash_tiles_jit specialised for forests/{forest} (what gpc_set_forest builds with NVRTC)", sass)


if __name__ == "__main__":
    main()
