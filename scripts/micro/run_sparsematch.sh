#!/bin/bash
# Writes a synthetic Sintel-shaped pair as PNGs and runs the CLI driver on it (GPU box).
python - <<'PY'
import sys; sys.path.insert(0, ".")
from PIL import Image
from opengpc_b200.synth import synth_pair
L, R = synth_pair(1024, 436, 1234)
Image.fromarray(L).save("/tmp/l.png"); Image.fromarray(R).save("/tmp/r.png")
PY
make -C samples > /dev/null && samples/sparsematch forests/defaultTauForest.txt /tmp/l.png /tmp/r.png /tmp/disp.png 20
