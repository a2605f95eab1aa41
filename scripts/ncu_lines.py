#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
blocks, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "func": "", "hdr": None, "lines": []}
        blocks.append(cur)
    elif r[0] == "Function Name" and cur is not None:
        cur["func"] = r[1]
    elif r[0] == "Line No":
        if cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r[0].isdigit():
        cur["lines"].append(r)
for b in blocks:
    hdr = b["hdr"]
    if not b["lines"]:
        continue
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    def num(v):
        try:
            return float(v or 0)
        except ValueError:
            return 0.0
    b["lines"] = [l for l in b["lines"] if len(l) > max(ci, cs)]
    tot_i = sum(num(l[ci]) for l in b["lines"]) or 1.0
    tot_s = sum(num(l[cs]) for l in b["lines"]) or 1.0
    print(f"\n==== {b['func'][:70]} | {b['file'].split('/')[-1]}: warp-inst {tot_i:.4g}, samples {tot_s:.0f}")
    agg = {}
    for l in b["lines"]:
        for i in stall_cols:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + float(l[i] or 0)
            except ValueError:
                pass
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
    print("   stalls:", ", ".join(f"{k[6:]} {100 * v / tot_s:.0f}%" for k, v in top))
    for l in b["lines"]:
        pi, ps = 100 * num(l[ci]) / tot_i, 100 * num(l[cs]) / tot_s
        if pi >= min_pct or ps >= min_pct:
            print(f"{l[0]:>5} inst {pi:5.1f}%  samp {ps:5.1f}%  | {l[1].strip()[:105]}")
