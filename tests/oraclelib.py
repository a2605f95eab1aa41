"""ctypes bindings for the TEST-ONLY checkers under oracle/ (see oracle/gpc_oracle.h).

`Oracle`    -> oracle/_build/libgpc_oracle.so  (plain-C restatement, always buildable)
`Reference` -> oracle/_ref/libgpc_ref.so       (unmodified reference headers; built only where
                                                /root/reference exists, travels prebuilt)
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libgpc_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgpc_ref.so")
REF_NAIVE_SO = os.path.join(ORACLE_DIR, "_ref", "libgpc_ref_naive.so")      # the reference's SSE=OFF build
FOREST_TAU = os.path.join(ROOT, "forests", "defaultTauForest.txt")
FOREST_ZERO = os.path.join(ROOT, "forests", "defaultZeroForest.txt")

SUPPORT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("d", "<f4")])


class GpcoForest(C.Structure):
    _fields_ = [("n_tests", C.c_int32), ("type", C.c_int32), ("n_discarded", C.c_int32),
                ("ix", C.c_int32 * 32), ("iy", C.c_int32 * 32), ("jx", C.c_int32 * 32),
                ("jy", C.c_int32 * 32), ("tau", C.c_int32 * 32)]


class GpcoSettings(C.Structure):
    _fields_ = [("gradient_threshold", C.c_int32), ("disp_high", C.c_int32),
                ("vertical_tolerance", C.c_int32), ("epipolar_mode", C.c_int32)]


def build_oracle(force=False):
    """Compile the checkers (gcc; the reference build only where /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "gpc_oracle.c")):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/lib/gpc"):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def settings(thr=5, disp_high=128, vt=0, epipolar=True):
    return GpcoSettings(int(thr), int(disp_high), int(vt), int(bool(epipolar)))


def digest(supp):
    """FNV-1a-64 variant over the ordered support list."""
    h = 1469598103934665603
    vals = np.stack([supp["x"].astype(np.int64), supp["y"].astype(np.int64),
                     supp["d"].astype(np.int64)], axis=1).reshape(-1) & 0xFFFFFFFF
    for v in vals.tolist():
        h ^= v
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


class Oracle:
    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.gpco_digest.restype = C.c_uint64
        L.gpco_candidates.restype = C.c_int
        L.gpco_find_correspondences.restype = C.c_int
        L.gpco_match.restype = C.c_int
        L.gpco_read_forest.restype = C.c_int
        L.gpco_pair.restype = C.c_int

    def synth(self, w, h, seed):
        Lm = np.empty((h, w), np.uint8)
        Rm = np.empty((h, w), np.uint8)
        self.lib.gpco_synth(_p(Lm), _p(Rm), C.c_int(w), C.c_int(h), C.c_uint32(seed))
        return Lm, Rm

    def digest(self, supp):
        supp = np.ascontiguousarray(supp)
        return int(self.lib.gpco_digest(_p(supp), C.c_int(len(supp))))

    def read_forest(self, path):
        f = GpcoForest()
        rc = self.lib.gpco_read_forest(path.encode(), C.byref(f))
        if rc != 0:
            raise FileNotFoundError(path)
        return f

    @staticmethod
    def make_forest(tests, taus, type_=None, n_discarded=0):
        """tests: iterable of (ix, iy, jx, jy); taus: iterable of int."""
        f = GpcoForest()
        tests = list(tests)[:32]
        taus = list(taus)[:32]
        f.n_tests = len(tests)
        for t, (ix, iy, jx, jy) in enumerate(tests):
            f.ix[t], f.iy[t], f.jx[t], f.jy[t] = ix, iy, jx, jy
            f.tau[t] = taus[t]
        f.type = int(any(taus)) if type_ is None else type_
        f.n_discarded = n_discarded
        return f

    def box(self, img):
        h, w = img.shape
        out = np.empty((h, w), np.uint8)
        self.lib.gpco_box(_p(np.ascontiguousarray(img)), _p(out), C.c_int(w), C.c_int(h))
        return out

    def sobel(self, img, thr):
        h, w = img.shape
        out = np.empty((h, w), np.uint8)
        self.lib.gpco_sobel(_p(np.ascontiguousarray(img)), _p(out), C.c_int(w), C.c_int(h), C.c_int(thr))
        return out

    def candidates(self, grad):
        h, w = grad.shape
        mask = np.empty(h * w + 1, np.int32)
        n = self.lib.gpco_candidates(_p(np.ascontiguousarray(grad)), C.c_int(w), C.c_int(h), _p(mask))
        return mask[:n].copy()

    def hash(self, smooth, forest, mask):
        h, w = smooth.shape
        mask = np.ascontiguousarray(mask, np.int32)
        st = np.empty(max(len(mask), 1), np.uint32)
        self.lib.gpco_hash(_p(np.ascontiguousarray(smooth)), C.c_int(w), C.c_int(h), C.byref(forest),
                           _p(mask), C.c_int(len(mask)), _p(st))
        return st[:len(mask)].copy()

    def find_correspondences(self, src, tar):
        src = np.ascontiguousarray(src, np.uint64)
        tar = np.ascontiguousarray(tar, np.uint64)
        out = np.empty(2 * max(min(len(src), len(tar)), 1), np.int32)
        m = self.lib.gpco_find_correspondences(_p(src), C.c_int(len(src)), _p(tar), C.c_int(len(tar)), _p(out))
        return out[:2 * m].reshape(-1, 2).copy()

    def hashmatch(self, src, tar):
        """useHashtable(true) matcher on bare keys: (src index, tar index) pairs in the reference's order."""
        src = np.ascontiguousarray(src, np.uint64)
        tar = np.ascontiguousarray(tar, np.uint64)
        out = np.empty(2 * max(min(len(src), len(tar)), 1), np.int32)
        m = self.lib.gpco_hashmatch(_p(src), C.c_int(len(src)), _p(tar), C.c_int(len(tar)), _p(out))
        return out[:2 * m].reshape(-1, 2).copy()

    def pair_hashtable(self, Lm, Rm, forest, s):
        h, w = Lm.shape
        supp = np.empty(max((w - 26) * (h - 26), 1), SUPPORT_DTYPE)
        n = self.lib.gpco_pair_hashtable(_p(np.ascontiguousarray(Lm)), _p(np.ascontiguousarray(Rm)), C.c_int(w),
                                         C.c_int(h), C.byref(forest), C.byref(s), _p(supp), None, None)
        return supp[:n].copy()

    def correspondences_hashtable(self, Lm, Rm, forest, s):
        h, w = Lm.shape
        thr = s.gradient_threshold
        _, _, mkl, stl = self.stages(Lm, forest, thr)
        _, _, mkr, str_ = self.stages(Rm, forest, thr)
        cap = max(min(len(mkl), len(mkr)), 1)
        corr = np.empty((cap, 4), np.int32)
        supp = np.empty(cap, SUPPORT_DTYPE)
        nc = C.c_int(0)
        self.lib.gpco_match_hashtable(_p(mkl), _p(stl), C.c_int(len(mkl)), _p(mkr), _p(str_), C.c_int(len(mkr)),
                                      C.c_int(w), C.byref(s), _p(corr), C.byref(nc), _p(supp))
        return corr[:nc.value].copy()

    def match(self, mask_l, st_l, mask_r, st_r, w, s):
        mask_l = np.ascontiguousarray(mask_l, np.int32)
        mask_r = np.ascontiguousarray(mask_r, np.int32)
        st_l = np.ascontiguousarray(st_l, np.uint32)
        st_r = np.ascontiguousarray(st_r, np.uint32)
        cap = max(min(len(mask_l), len(mask_r)), 1)
        supp = np.empty(cap, SUPPORT_DTYPE)
        n = self.lib.gpco_match(_p(mask_l), _p(st_l), C.c_int(len(mask_l)), _p(mask_r), _p(st_r),
                                C.c_int(len(mask_r)), C.c_int(w), C.byref(s), None, None, _p(supp))
        return supp[:n].copy()

    def correspondences(self, Lm, Rm, forest, s):
        """stereoMatch: every unique-unique correspondence (xs, ys, xt, yt) before the filter."""
        h, w = Lm.shape
        thr = s.gradient_threshold
        _, _, mkl, stl = self.stages(Lm, forest, thr)
        _, _, mkr, str_ = self.stages(Rm, forest, thr)
        cap = max(min(len(mkl), len(mkr)), 1)
        corr = np.empty((cap, 4), np.int32)
        supp = np.empty(cap, SUPPORT_DTYPE)
        nc = C.c_int(0)
        self.lib.gpco_match(_p(mkl), _p(stl), C.c_int(len(mkl)), _p(mkr), _p(str_), C.c_int(len(mkr)), C.c_int(w),
                            C.byref(s), _p(corr), C.byref(nc), _p(supp))
        return corr[:nc.value].copy()

    def pair(self, Lm, Rm, forest, s):
        h, w = Lm.shape
        supp = np.empty(max((w - 26) * (h - 26), 1), SUPPORT_DTYPE)
        ncl, ncr = C.c_int(0), C.c_int(0)
        n = self.lib.gpco_pair(_p(np.ascontiguousarray(Lm)), _p(np.ascontiguousarray(Rm)), C.c_int(w),
                               C.c_int(h), C.byref(forest), C.byref(s), _p(supp), C.byref(ncl), C.byref(ncr))
        return supp[:n].copy(), ncl.value, ncr.value

    def stages(self, img, forest, thr):
        """All per-image intermediates: smooth, grad, mask, states."""
        sm = self.box(img)
        gr = self.sobel(img, thr)
        mk = self.candidates(gr)
        st = self.hash(sm, forest, mk)
        return sm, gr, mk, st

    # ---- the reference's SSE=OFF build (the *Naive functions of filter.hpp) ---------------------
    def stages_naive(self, img, forest, thr):
        h, w = img.shape
        img = np.ascontiguousarray(img)
        sm = np.empty((h, w), np.uint8)
        gr = np.empty((h, w), np.uint8)
        self.lib.gpco_box_naive(_p(img), _p(sm), C.c_int(w), C.c_int(h))
        self.lib.gpco_sobel_naive(_p(img), _p(gr), C.c_int(w), C.c_int(h), C.c_int(thr))
        mk = self.candidates(gr)
        st = np.empty(max(len(mk), 1), np.uint32)
        self.lib.gpco_hash_naive(_p(sm), C.c_int(w), C.c_int(h), C.byref(forest), _p(mk), C.c_int(len(mk)), _p(st))
        return sm, gr, mk, st[:len(mk)].copy()

    def pair_naive(self, Lm, Rm, forest, s, use_hashtable=False):
        h, w = Lm.shape
        supp = np.empty(max((w - 26) * (h - 26), 1), SUPPORT_DTYPE)
        ncl, ncr = C.c_int(0), C.c_int(0)
        n = self.lib.gpco_pair_naive(_p(np.ascontiguousarray(Lm)), _p(np.ascontiguousarray(Rm)), C.c_int(w), C.c_int(h),
                                     C.byref(forest), C.byref(s), C.c_int(int(use_hashtable)), _p(supp), C.byref(ncl), C.byref(ncr))
        return supp[:n].copy(), ncl.value, ncr.value


class Reference:
    """The compiled, unmodified reference (None-like if the prebuilt library is absent)."""

    @staticmethod
    def available(naive=False):
        so = REF_NAIVE_SO if naive else REF_SO
        if not os.path.exists(so) and os.path.isdir("/root/reference/lib/gpc"):
            build_oracle(force=True)
        return os.path.exists(so)

    def __init__(self, naive=False):
        """naive=True: the reference compiled without -D_INTRINSICS_SSE (its SSE=OFF result mode)."""
        if not self.available(naive):
            raise RuntimeError("oracle/_ref reference build missing (needs /root/reference)")
        self.lib = C.CDLL(REF_NAIVE_SO if naive else REF_SO)
        assert self.lib.ref_is_sse_build() == (0 if naive else 1)
        L = self.lib
        for name in ("ref_read_forest", "ref_preprocess", "ref_hash", "ref_find_correspondences", "ref_pair",
                     "ref_sizeof_descriptor", "ref_sizeof_support", "ref_sizeof_correspondence"):
            getattr(L, name).restype = C.c_int
        L.ref_time_pairs.restype = C.c_double

    def read_forest(self, path, w, h):
        off = np.zeros(64, np.int32)
        tau = np.zeros(32, np.int32)
        typ, ntau = C.c_int(0), C.c_int(0)
        T = self.lib.ref_read_forest(path.encode(), C.c_int(w), C.c_int(h), _p(off), _p(tau),
                                     C.byref(typ), C.byref(ntau))
        return off[:2 * T].copy(), tau[:ntau.value].copy(), typ.value

    def preprocess(self, img, thr):
        h, w = img.shape
        sm = np.empty((h, w), np.uint8)
        gr = np.empty((h, w), np.uint8)
        mk = np.empty(h * w, np.int32)
        n = self.lib.ref_preprocess(_p(np.ascontiguousarray(img)), C.c_int(w), C.c_int(h), C.c_int(thr),
                                    _p(sm), _p(gr), _p(mk))
        return sm, gr, mk[:n].copy()

    def hash(self, img, thr, forest_path, num_threads=1):
        h, w = img.shape
        st = np.empty(h * w, np.uint32)
        n = self.lib.ref_hash(_p(np.ascontiguousarray(img)), C.c_int(w), C.c_int(h), C.c_int(thr),
                              forest_path.encode(), C.c_int(num_threads), _p(st))
        return st[:n].copy()

    def find_correspondences(self, src, tar):
        src = np.ascontiguousarray(src, np.uint64)
        tar = np.ascontiguousarray(tar, np.uint64)
        out = np.empty(2 * max(min(len(src), len(tar)), 1), np.int32)
        m = self.lib.ref_find_correspondences(_p(src), C.c_int(len(src)), _p(tar), C.c_int(len(tar)), _p(out))
        return out[:2 * m].reshape(-1, 2).copy()

    def hashmatch(self, src, tar):
        src = np.ascontiguousarray(src, np.uint64)
        tar = np.ascontiguousarray(tar, np.uint64)
        out = np.empty(2 * max(min(len(src), len(tar)), 1), np.int32)
        m = self.lib.ref_hashmatch(_p(src), C.c_int(len(src)), _p(tar), C.c_int(len(tar)), _p(out))
        return out[:2 * m].reshape(-1, 2).copy()

    def pair_hashtable(self, Lm, Rm, forest_path, thr=5, disp_high=128, vt=0, epipolar=True):
        h, w = Lm.shape
        cap = max((w - 26) * (h - 26), 1)
        supp = np.empty(cap, SUPPORT_DTYPE)
        n = self.lib.ref_pair_hashtable(_p(np.ascontiguousarray(Lm)), _p(np.ascontiguousarray(Rm)), C.c_int(w),
                                        C.c_int(h), forest_path.encode(), C.c_int(thr), C.c_int(disp_high),
                                        C.c_int(vt), C.c_int(int(epipolar)), _p(supp), C.c_int(cap))
        assert n >= 0
        return supp[:n].copy()

    def pair(self, Lm, Rm, forest_path, thr=5, disp_high=128, vt=0, epipolar=True, num_threads=1):
        h, w = Lm.shape
        cap = max((w - 26) * (h - 26), 1)
        supp = np.empty(cap, SUPPORT_DTYPE)
        ncl, ncr = C.c_int(0), C.c_int(0)
        t_pre, t_match = C.c_double(0), C.c_double(0)
        n = self.lib.ref_pair(_p(np.ascontiguousarray(Lm)), _p(np.ascontiguousarray(Rm)), C.c_int(w), C.c_int(h),
                              forest_path.encode(), C.c_int(thr), C.c_int(disp_high), C.c_int(vt),
                              C.c_int(int(epipolar)), C.c_int(num_threads), _p(supp), C.c_int(cap),
                              C.byref(ncl), C.byref(ncr), C.byref(t_pre), C.byref(t_match))
        assert n >= 0
        return supp[:n].copy(), ncl.value, ncr.value, (t_pre.value, t_match.value)

    def time_pairs(self, images, forest_path, threads, iters, thr=5, disp_high=128, vt=0, epipolar=True):
        """images: uint8 [n_pairs, 2, h, w].  Returns (wall seconds, total supports)."""
        images = np.ascontiguousarray(images, np.uint8)
        n_pairs, _, h, w = images.shape
        tot = C.c_longlong(0)
        sec = self.lib.ref_time_pairs(_p(images), C.c_int(n_pairs), C.c_int(w), C.c_int(h), forest_path.encode(),
                                      C.c_int(thr), C.c_int(disp_high), C.c_int(vt), C.c_int(int(epipolar)),
                                      C.c_int(threads), C.c_int(iters), C.byref(tot))
        return sec, tot.value


# ---- wide states (forests of more than 32 tests; 32-test forests of the SSE=OFF build) ---------------------------
def wide_words(o, img, tests, thr, naive):
    """Scalar restatement of the word planes of include/gpc_b200.h's "extended mode": (mask, words[n_words][n_cand]).
    SSE results: word k = tests 32k .. 32k+31 hashed as a forest of their own (gpco_hash).  Naive results: the T-bit
    state of gpcFilter[Tau]Naive cut into 31-bit words (word k = state bits 31k .. 31k+30)."""
    tests = np.asarray(tests, np.int32).reshape(-1, 5)
    T = len(tests)
    type_ = int(np.any(tests[:, 4] != 0))
    per = 31 if naive else 32
    K = (T + per - 1) // per
    if naive:
        sm, gr, mk, _ = o.stages_naive(img, o.make_forest([], [], type_=0), thr)
    else:
        sm, gr, mk, _ = o.stages(img, o.make_forest([], [], type_=0), thr)
    h, w = img.shape
    words = []
    for k in range(K):
        tk = min(per, T - per * k)
        t0 = T - per * k - tk if naive else per * k
        sub = tests[t0:t0 + tk]
        f = o.make_forest([tuple(int(v) for v in r[:4]) for r in sub], [int(r[4]) for r in sub], type_=type_)
        st = np.empty(max(len(mk), 1), np.uint32)
        fn = o.lib.gpco_hash_naive if naive else o.lib.gpco_hash
        fn(_p(np.ascontiguousarray(sm)), C.c_int(w), C.c_int(h), C.byref(f), _p(np.ascontiguousarray(mk, np.int32)), C.c_int(len(mk)), _p(st))
        words.append(st[:len(mk)].copy())
    return mk, np.stack(words) if words else np.zeros((0, len(mk)), np.uint32)


def pair_wide(o, Lm, Rm, tests, s, naive=False):
    """Supports of the extended mode: tuples compared from the last word down, dense ranks, then the ordinary
    findCorrespondences + rectifiedMatch filter restatement (gpco_match) on the ranks."""
    mkl, wl = wide_words(o, Lm, tests, s.gradient_threshold, naive)
    mkr, wr = wide_words(o, Rm, tests, s.gradient_threshold, naive)
    both = np.concatenate([wl, wr], axis=1)[::-1].T            # rows = candidates, columns = last word first
    _, inv = np.unique(both, axis=0, return_inverse=True)       # lexicographic, dense
    inv = inv.reshape(-1).astype(np.uint32)
    h, w = Lm.shape
    return o.match(mkl, inv[:len(mkl)], mkr, inv[len(mkl):], w, s), len(mkl), len(mkr)
