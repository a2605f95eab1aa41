#!/usr/bin/env python
"""Offline view of the forest-specialised kernel A2 (what jit.cu builds with NVRTC at gpc_set_forest time).

usage: python scripts/jit_sass.py forests/defaultZeroForest.txt [-DNAME ...] [--sass out.sass]
Generates the same header jit.cu generates (kJitTests, kJitTau, jit_imm_a / jit_imm_b / jit_mtau2), compiles
csrc/hash_tiles.cu with nvcc for sm_100a (no GPU needed) and prints registers, SASS instruction histogram and the
per-pipe counts of the main loop.  Development tool only; nothing in the product path uses it."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "opengpc_b200", "csrc")
TILE_W, TILE_H, RADIUS = 128, 64, 13


def read_tests(path):
    tok = open(path).read().split()
    n_ferns, i, tests = int(tok[0]), 1, []
    for _ in range(n_ferns):
        nt = int(tok[i + 2]); i += 3
        for _ in range(nt):
            tests.append(tuple(int(v) for v in tok[i + 1:i + 6])); i += 6
    return tests[:32]


def header(tests, tile_w=TILE_W, tile_h=TILE_H):
    pitch, rows = tile_w + 32, tile_h + 2 * RADIUS
    copy_bytes = (rows * pitch + 127) // 128 * 128

    def imm(dx, dy):
        o = dy * pitch + dx
        k = o % 4
        return k * copy_bytes + (o - k)

    tau = any(t[4] != 0 for t in tests)

    def sw(name, typ, vals):
        return ("__device__ __forceinline__ constexpr %s %s(int t) { switch (t) {" % (typ, name) +
                "".join(" case %d: return %s;" % (i, v) for i, v in enumerate(vals)) + " default: return 0; } }\n")

    def mt(t):
        tau8 = ((t[4] + 128) & 255) - 128
        m = (-tau8) & 0xffff
        return "0x%08xu" % ((m | (m << 16)) if tau else 0)

    return ("constexpr int kJitTests = %d;\nconstexpr bool kJitTau = %s;\n" % (len(tests), "true" if tau else "false") +
            sw("jit_imm_a", "int", [imm(t[0], t[1]) for t in tests]) + sw("jit_imm_b", "int", [imm(t[2], t[3]) for t in tests]) +
            sw("jit_mtau2", "unsigned", [mt(t) for t in tests]))


def main():
    forest = sys.argv[1]
    defs = [a for a in sys.argv[2:] if a.startswith("-D")]
    sass_out = sys.argv[sys.argv.index("--sass") + 1] if "--sass" in sys.argv else None
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "gpc_jit_forest.h"), "w") as f:
            f.write(header(read_tests(forest)))
        cubin = os.path.join(tmp, "a2.cubin")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-lineinfo", "-Xptxas", "-v",
               "-DGPC_JIT_HEADER=\"gpc_jit_forest.h\"", "-I", tmp, "-I", CSRC] + defs + ["-o", cubin, os.path.join(CSRC, "hash_tiles.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit(r.stderr)
        prev = ""
        for line in r.stderr.splitlines():
            if "gpc_hash_tiles_jit" in line or ("registers" in line and "jit" in prev):
                print(line.strip())
            prev = line
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "gpc_hash_tiles_jit", cubin], capture_output=True, text=True).stdout
    if sass_out:
        open(sass_out, "w").write(sass)
    ops = [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass, re.M)]
    hist = collections.Counter(o.split(".")[0] for o in ops)
    print("instructions:", len(ops))
    print(", ".join("%s %d" % kv for kv in hist.most_common(24)))


if __name__ == "__main__":
    main()
