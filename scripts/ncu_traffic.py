#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` reports captured at the benchmarked batch (scripts/gpu_profile.sh):
dram__bytes_read.sum + dram__bytes_write.sum per launch, divided by the launch's pixel count (kernels A1 / A2: pixels of all
images of the step, matcher kernels: pixels of one image per pair).  bench.py multiplies back by the pixels of its step.
The bench's "match_rows" bucket is the sum of the row matcher's kernels (fast + tail); "emit_supports" is kernel C.
usage: python scripts/ncu_traffic.py <round tag> <pairs per captured launch> report.ncu-rep [more reports]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
PIX = 1024 * 436
B = int(sys.argv[2])
BUCKET = {"smooth_sobel": ("smooth_sobel", 2 * B * PIX), "hash_tiles": ("hash_tiles", 2 * B * PIX), "match_rows_fast": ("match_rows", B * PIX),
          "match_rows_tail": ("match_rows", B * PIX), "emit_supports": ("emit_supports", B * PIX)}
out = {"source": f"ncu --set full, {sys.argv[1]} (profiles/), launches of {B} pairs (the benchmarked step of {2 * B} pairs runs as two such launches on two streams)",
       "dram_bytes_per_pixel": {},
       "per_kernel_dram_bytes_per_launch": {}}
for rep in sys.argv[3:]:
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = next((k for k in BUCKET if k in name), None)
        if key is None or key in seen:
            continue
        seen.add(key)
        tot = sum(float(r[col[m]].replace(",", "")) * MULT[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        bucket, pix = BUCKET[key]
        out["per_kernel_dram_bytes_per_launch"][key] = tot
        out["dram_bytes_per_pixel"][bucket] = out["dram_bytes_per_pixel"].get(bucket, 0.0) + tot / pix
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
