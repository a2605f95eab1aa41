"""gpc_pool -- the library-level multi-GPU driver: one context + host thread per device, chunks dealt round-robin,
supports gathered in pair order directly in the caller's buffer.  Runs on one GPU too (a device may be listed more than
once: several contexts share it), and on every visible GPU when the box has more (gpurun --gpus 2)."""
import numpy as np
import pytest

from helpers import FORESTS

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
@pytest.mark.parametrize("mode", ["rows", "global"])
def test_pool_equals_single_context(devices, mode):
    import opengpc_b200 as g
    from opengpc_b200.synth import sparsify, synth_batch
    imgs = synth_batch(512, 160, 23, seed0=300)
    imgs[5, 0], imgs[5, 1] = sparsify(imgs[5, 0]), sparsify(imgs[5, 1])
    imgs[11] = 128                                                   # a flat pair: no candidates at all
    s = g.sparsematch_settings() if mode == "rows" else g.make_settings(thr=10, disp_high=128, vt=1, epipolar=False)
    with g.Context(device=0, max_w=512, max_h=160, max_batch=23) as ctx:
        ctx.set_forest(FORESTS["tau"])
        want, woff, wnc = ctx.match_batch(imgs, s)
        want = want.copy()
    with g.Pool(devices, max_w=512, max_h=160, max_batch_per_device=23) as pool:
        pool.set_forest(FORESTS["tau"])
        for rep in range(3):
            got, off, nc = pool.match_batch(imgs, s)
            assert np.array_equal(off, woff) and np.array_equal(nc, wnc), (devices, rep)
            assert np.array_equal(got, want), (devices, rep)
        # capacity too small: reported with the need, nothing written past the end
        with pytest.raises(g.GpcError) as e:
            pool.match_batch(imgs, s, cap=int(woff[-1]) - 1)
        assert str(int(woff[-1])) in str(e.value)
        small, off1, _ = pool.match_batch(imgs[:1], s)              # fewer pairs than devices
        assert np.array_equal(small, want[:woff[1]])


def test_pool_every_visible_gpu():
    """Pairs sharded over every GPU of the box (weak scaling unit = the pair, no collective): same bytes as one GPU."""
    import opengpc_b200 as g
    from opengpc_b200.synth import synth_batch
    n = _n_gpus()
    if n < 2:
        pytest.skip("one GPU visible: covered by the duplicated-device cases above")
    imgs = synth_batch(1024, 436, 8 * n + 3, seed0=1234)
    s = g.sparsematch_settings()
    with g.Context(device=0, max_w=1024, max_h=436, max_batch=len(imgs)) as ctx:
        ctx.set_forest(FORESTS["zero"])
        want, woff, _ = ctx.match_batch(imgs, s)
        want = want.copy()
    with g.Pool(list(range(n)), max_w=1024, max_h=436, max_batch_per_device=len(imgs)) as pool:
        pool.set_forest(FORESTS["zero"])
        got, off, _ = pool.match_batch(imgs, s)
    assert np.array_equal(off, woff) and np.array_equal(got, want)
