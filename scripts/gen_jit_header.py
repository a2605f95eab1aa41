#!/usr/bin/env python
"""Experiment helper: write a header that bakes one forest into kernel A2 (see GPC_JIT_HEADER in hash_tiles.cu).
usage: python scripts/gen_jit_header.py <forest.txt> <out.h> [tile_h=64] [tile_w=128]"""
import sys
sys.path.insert(0, ".")
import opengpc_b200 as g

f = g.read_forest(sys.argv[1])
th = int(sys.argv[3]) if len(sys.argv) > 3 else 64
tw = int(sys.argv[4]) if len(sys.argv) > 4 else 128
pitch = tw + 32
copy = ((th + 26) * pitch + 127) // 128 * 128


def imm(dx, dy):
    o = dy * pitch + dx
    k = o % 4
    return k * copy + (o - k)


def fn(name, typ, vals):
    body = " ".join(f"case {t}: return {v};" for t, v in enumerate(vals))
    return f"__device__ __forceinline__ constexpr {typ} {name}(int t) {{ switch (t) {{ {body} default: return 0; }} }}\n"


T = f.n_tests
mt = []
for t in range(T):
    tau8 = (f.tau[t] + 128) % 256 - 128
    m = (-tau8) & 0xffff
    mt.append(f"0x{(m | (m << 16)) if f.type == 1 else 0:08x}u")
with open(sys.argv[2], "w") as out:
    out.write(f"// generated from {sys.argv[1]}\nconstexpr int kJitTests = {T};\n")
    out.write(fn("jit_imm_a", "int", [imm(f.ix[t], f.iy[t]) for t in range(T)]))
    out.write(fn("jit_imm_b", "int", [imm(f.jx[t], f.jy[t]) for t in range(T)]))
    out.write(fn("jit_mtau2", "unsigned", mt))
