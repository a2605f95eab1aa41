import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import opengpc_b200 as g
from opengpc_b200.synth import synth_pair
L, R = synth_pair(256, 64, 1)
with g.Context(device=0, max_w=256, max_h=64, max_batch=1) as c:
    c.set_forest('forests/defaultTauForest.txt')
    try:
        sm, gr, mk = c.preprocess(L, 5)
        print('preprocess ok', len(mk))
        s, a, b = c.match_pair(L, R, g.sparsematch_settings())
        print('pair ok', len(s), a, b)
    except Exception as e:
        print('ERR', e)
