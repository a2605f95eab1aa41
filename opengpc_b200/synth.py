"""Synthetic stereo pairs of SURVEY.md appendix C (numpy, vectorised).

Input generation only -- not part of the measured path.  Consumes raw std::mt19937 outputs
(numpy's legacy RandomState(seed) uses the same init_genrand seeding and yields the same
32-bit words through .bytes()), so the images are identical to the survey's C++ generator.
"""
import numpy as np


def _raw_u32(rs, n):
    return np.frombuffer(rs.bytes(4 * n), dtype="<u4")


def synth_pair(w, h, seed=1234):
    """Return (L, R) uint8 [h, w] with ground-truth disparity d(y) = 5 + 40*y//h."""
    rs = np.random.RandomState(seed)
    cw, ch = 2 * w // 4 + 2, h // 4 + 2
    coarse = (_raw_u32(rs, cw * ch) & 255).astype(np.int32).reshape(ch, cw)
    noise = (_raw_u32(rs, h * 2 * w) & 255).astype(np.int32).reshape(h, 2 * w)
    v = coarse[np.arange(h)[:, None] // 4, np.arange(2 * w)[None, :] // 4]
    tex = ((v * 3 + noise) // 4).astype(np.int32)
    n = (_raw_u32(rs, h * w) % 3).astype(np.int32).reshape(h, w) - 1
    ys = np.arange(h)[:, None]
    xs = np.arange(w)[None, :]
    d = 5 + (ys * 40) // h
    L = tex[ys, xs + 200].astype(np.uint8)
    R = np.clip(tex[ys, xs + d + 200] + n, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(L), np.ascontiguousarray(R)


def synth_batch(w, h, n_pairs, seed0=1234, sparse=False):
    """uint8 [n_pairs, 2, h, w]; pair i uses seed0 + i."""
    out = np.empty((n_pairs, 2, h, w), np.uint8)
    for i in range(n_pairs):
        L, R = synth_pair(w, h, seed0 + i)
        if sparse:
            L, R = sparsify(L), sparsify(R)
        out[i, 0], out[i, 1] = L, R
    return out


def sparsify(img, tile=64):
    """Low-texture variant: keep the texture on a 25 % checkerboard of `tile`-pixel tiles and
    flatten the rest to mid-grey, so that many rows carry 0-2 candidates."""
    h, w = img.shape
    ty = (np.arange(h)[:, None] // tile)
    tx = (np.arange(w)[None, :] // tile)
    keep = ((ty % 2) == 0) & ((tx % 2) == 0)
    out = np.where(keep, img, 128).astype(np.uint8)
    return np.ascontiguousarray(out)


def downsample2x(img):
    """Pyramid level l+1 from level l: 2x2 mean with floor, (a + b + c + d) // 4."""
    h, w = img.shape
    v = img[:h // 2 * 2, :w // 2 * 2].astype(np.uint16)
    return np.ascontiguousarray(((v[0::2, 0::2] + v[0::2, 1::2] + v[1::2, 0::2] + v[1::2, 1::2]) >> 2).astype(np.uint8))
