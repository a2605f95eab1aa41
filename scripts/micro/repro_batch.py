import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import opengpc_b200 as g
from opengpc_b200.synth import synth_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = synth_batch(1024, 436, 8, seed0=1234)
imgs = np.ascontiguousarray(np.tile(base, (B // 8, 1, 1, 1)))
with g.Context(device=0, max_w=1024, max_h=436, max_batch=B) as ctx:
    ctx.set_forest("/root/repo/forests/defaultZeroForest.txt")
    s = g.sparsematch_settings()
    for rep in range(4):
        supp, off, nc = ctx.match_batch(imgs, s)
        print(rep, len(supp), off[1], flush=True)
