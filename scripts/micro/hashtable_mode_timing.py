"""Throughput of useHashtable(true) through gpc_match_batch (pinned host buffers)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import opengpc_b200 as g
from opengpc_b200.synth import synth_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
imgs = pin(np.tile(synth_batch(1024, 436, 4, seed0=1234), (B // 4, 1, 1, 1)))
with g.Context(device=0, max_w=1024, max_h=436, max_batch=B) as c:
    c.set_forest("forests/defaultTauForest.txt")
    s = g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=True)
    out = pin(np.empty(B * 60000 * 3, np.int32)).view(g.SUPPORT_DTYPE)
    for _ in range(2):
        supp, offs, _ = c.match_batch(imgs, s, out=out)
    t0 = time.perf_counter()
    for _ in range(5):
        supp, offs, _ = c.match_batch(imgs, s, out=out)
    dt = (time.perf_counter() - t0) / 5
    print(f"hashtable mode: {B / dt:.0f} pairs/s ({1e3 * dt / B:.3f} ms per pair), supports per pair {len(supp) / B:.0f}")
