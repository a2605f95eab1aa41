#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` reports of the small bench (--batch 32: 64 images, 32 pairs):
dram__bytes_read.sum + dram__bytes_write.sum per launch, divided by the launch's pixel count (kernel A: image pixels, kernel B: pixels of one image per pair).
usage: python scripts/ncu_traffic.py <round tag> kernelA.ncu-rep kernelB.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
PIX = 1024 * 436
UNITS = {"smooth_sobel": 64 * PIX, "hash_tiles": 64 * PIX, "match_rows": 32 * PIX}      # image pixels / pair pixels (one side) per launch at --batch 32
out = {"source": f"ncu --set full, {sys.argv[1]} (profiles/), small bench --batch 32", "dram_bytes_per_pixel": {}}
for rep in sys.argv[2:]:
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = next((k for k in UNITS if k in name), None)
        if key is None:
            continue
        tot = sum(float(r[col[m]].replace(",", "")) * MULT[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out["dram_bytes_per_pixel"][key] = tot / UNITS[key]
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
