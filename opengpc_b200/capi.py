"""ctypes binding of the C ABI in include/gpc_b200.h (libgpc_b200.so).

This is the thin host layer tests and bench.py drive; the drop-in C++ API lives in
include/gpc/*.hpp.  There is no CPU fallback: if the CUDA library is missing or no device is
present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPC_B200_LIB") or os.path.join(PKG, "libgpc_b200.so")   # override: kernel tuning builds

GPC_OK, GPC_E_ARG, GPC_E_WIDTH16, GPC_E_DIMS, GPC_E_CUDA, GPC_E_CAPACITY, GPC_E_UNSUPPORTED, GPC_E_FOREST, GPC_E_IO = range(9)

SUPPORT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("d", "<f4")])   # == ndb::Support, 12 bytes
CORR_DTYPE = np.dtype([("xs", "<i4"), ("ys", "<i4"), ("xt", "<i4"), ("yt", "<i4")])   # == ndb::Correspondence
MATCHER_AUTO, MATCHER_SORT, MATCHER_ROWS_GENERAL = 0, 1, 2

# every symbol include/gpc_b200.h declares (checked by tests/test_capi_cpu.py)
SYMBOLS = ["gpc_create", "gpc_destroy", "gpc_last_error", "gpc_status_string", "gpc_set_stream", "gpc_synchronize",
           "gpc_read_forest", "gpc_set_forest", "gpc_match_pair", "gpc_match_batch", "gpc_match_batch_device",
           "gpc_preprocess", "gpc_hash", "gpc_match_hash_images", "gpc_launch_count", "gpc_enable_kernel_timing",
           "gpc_kernel_times", "gpc_hash_smooth", "gpc_image_upload", "gpc_image_release", "gpc_image_preprocess",
           "gpc_match_images", "gpc_correspond_images", "gpc_find_correspondences", "gpc_hashmatch", "gpc_set_result_mode", "gpc_set_matcher", "gpc_match_pyramid", "gpc_jit_status", "gpc_context_id", "gpc_image_fetch", "gpc_fetch_supports", "gpc_mask_view",
           "gpc_pool_create", "gpc_pool_destroy", "gpc_pool_size", "gpc_pool_context", "gpc_pool_last_error", "gpc_pool_launch_count",
           "gpc_pool_set_forest", "gpc_pool_set_result_mode", "gpc_pool_match_batch", "gpc_host_alloc", "gpc_host_free",
           "gpc_read_forest_tests", "gpc_set_wide_forest", "gpc_match_pair_wide", "gpc_hash_wide"]
KERNEL_NAMES = ["smooth_sobel", "hash_tiles", "match_rows", "scans", "emit_supports"]


class GpcSettings(C.Structure):
    """gpc_settings == gpc::inference::InferenceSettings (inference.hpp:71-131)."""
    _fields_ = [("gradient_threshold", C.c_int32), ("disp_high", C.c_int32), ("vertical_tolerance", C.c_int32),
                ("epipolar_mode", C.c_int32), ("use_hashtable", C.c_int32), ("num_threads", C.c_int32)]


class GpcForest(C.Structure):
    _fields_ = [("n_tests", C.c_int32), ("type", C.c_int32), ("n_discarded", C.c_int32),
                ("ix", C.c_int32 * 32), ("iy", C.c_int32 * 32), ("jx", C.c_int32 * 32), ("jy", C.c_int32 * 32),
                ("tau", C.c_int32 * 32), ("n_ferns", C.c_int32)]


class GpcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gpc status {status}: {message}")
        self.status = status


_lib = None


def load_library():
    """dlopen libgpc_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(opengpc_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gpc_last_error.restype = C.c_char_p
    lib.gpc_last_error.argtypes = [C.c_void_p]
    lib.gpc_status_string.restype = C.c_char_p
    lib.gpc_jit_status.restype = C.c_char_p
    lib.gpc_jit_status.argtypes = [C.c_void_p]
    lib.gpc_launch_count.restype = C.c_int64
    lib.gpc_launch_count.argtypes = [C.c_void_p]
    lib.gpc_destroy.restype = None
    lib.gpc_destroy.argtypes = [C.c_void_p]
    lib.gpc_image_release.restype = None
    lib.gpc_image_release.argtypes = [C.c_void_p]
    lib.gpc_pool_destroy.restype = None
    lib.gpc_pool_destroy.argtypes = [C.c_void_p]
    lib.gpc_pool_last_error.restype = C.c_char_p
    lib.gpc_pool_last_error.argtypes = [C.c_void_p]
    lib.gpc_pool_launch_count.restype = C.c_int64
    lib.gpc_pool_launch_count.argtypes = [C.c_void_p]
    lib.gpc_pool_context.restype = C.c_void_p
    lib.gpc_pool_context.argtypes = [C.c_void_p, C.c_int]
    lib.gpc_host_alloc.restype = C.c_void_p
    lib.gpc_host_alloc.argtypes = [C.c_size_t]
    lib.gpc_host_free.restype = None
    lib.gpc_host_free.argtypes = [C.c_void_p]
    lib.gpc_mask_view.restype = C.c_void_p
    lib.gpc_mask_view.argtypes = [C.c_void_p, C.c_void_p]
    lib.gpc_context_id.restype = C.c_int64
    lib.gpc_context_id.argtypes = [C.c_void_p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    _lib = lib
    return lib


def make_settings(thr=10, disp_high=128, vt=1, epipolar=False, use_hashtable=False, num_threads=1):
    """Defaults are the reference's (inference.hpp:74-89)."""
    return GpcSettings(int(thr), int(disp_high), int(vt), int(bool(epipolar)), int(bool(use_hashtable)), int(num_threads))


def sparsematch_settings():
    """The settings samples/sparsematch.cpp:29-34 fixes."""
    return make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=False)


def read_forest(path):
    """Forest::readForest's parser (inference.hpp:404-446); GpcError(GPC_E_IO) if unreadable."""
    lib = load_library()
    f = GpcForest()
    rc = lib.gpc_read_forest(os.fsencode(path), C.byref(f))
    if rc != GPC_OK:
        raise GpcError(rc, lib.gpc_status_string(rc).decode())
    return f


def read_forest_tests(path):
    """Every test of a forest file, [n, 5] int32 rows of (ix, iy, jx, jy, tau): gpc_read_forest without the 32-test cap."""
    lib = load_library()
    n, nf = C.c_int(0), C.c_int(0)
    rc = lib.gpc_read_forest_tests(os.fsencode(path), None, 0, C.byref(n), C.byref(nf))
    if rc != GPC_OK:
        raise GpcError(rc, lib.gpc_status_string(rc).decode())
    tests = np.zeros((max(n.value, 1), 5), np.int32)
    lib.gpc_read_forest_tests(os.fsencode(path), _ptr(tests), C.c_int(len(tests)), C.byref(n), C.byref(nf))
    return tests[:n.value]


def make_forest(tests, taus, type_=None):
    f = GpcForest()
    tests, taus = list(tests), list(taus)
    f.n_discarded = max(len(tests) - 32, 0)
    nz = any(t != 0 for t in taus)
    tests, taus = tests[:32], taus[:32]
    f.n_tests = len(tests)
    for t, (ix, iy, jx, jy) in enumerate(tests):
        f.ix[t], f.iy[t], f.jx[t], f.jy[t], f.tau[t] = ix, iy, jx, jy, taus[t]
    f.type = int(nz) if type_ is None else int(type_)
    return f


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Context:
    """One resident context per GPU (gpc_create / gpc_destroy)."""

    def __init__(self, device=0, max_w=1024, max_h=436, max_batch=1):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.gpc_create(C.byref(self._h), int(device), int(max_w), int(max_h), int(max_batch))
        if rc != GPC_OK:
            raise GpcError(rc, self.lib.gpc_last_error(None).decode())
        self.device, self.max_w, self.max_h, self.max_batch = device, max_w, max_h, max_batch

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.gpc_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != GPC_OK:
            raise GpcError(rc, self.lib.gpc_last_error(self._h).decode())

    @property
    def launches(self):
        return int(self.lib.gpc_launch_count(self._h))

    @property
    def jit_status(self):
        """'specialised' if kernel A2 was rebuilt for the current forest, else 'generic: <reason>'."""
        return self.lib.gpc_jit_status(self._h).decode()

    def enable_kernel_timing(self, on=True):
        self._check(self.lib.gpc_enable_kernel_timing(self._h, int(bool(on))))

    def kernel_times(self):
        """(dict kernel -> accumulated ms, number of batch runs); synchronises the stream."""
        ms = (C.c_double * len(KERNEL_NAMES))()
        runs = C.c_int64(0)
        self._check(self.lib.gpc_kernel_times(self._h, ms, C.byref(runs)))
        return dict(zip(KERNEL_NAMES, list(ms))), runs.value

    def set_stream(self, cuda_stream):
        self._check(self.lib.gpc_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self.lib.gpc_synchronize(self._h))

    def set_forest(self, forest):
        if isinstance(forest, (str, bytes, os.PathLike)):
            forest = read_forest(forest)
        self._check(self.lib.gpc_set_forest(self._h, C.byref(forest)))
        return forest

    # ---- whole path -------------------------------------------------------------------------
    def match_pair(self, left, right, settings, cap=None):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape
        cap = max((w - 26) * (h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), SUPPORT_DTYPE)
        n, ncl, ncr = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = self.lib.gpc_match_pair(self._h, _ptr(left), _ptr(right), w, h, w, C.byref(settings), _ptr(out),
                                     C.c_int(cap), C.byref(n), C.byref(ncl), C.byref(ncr))
        self._check(rc)
        return out[:n.value].copy(), ncl.value, ncr.value

    def match_batch(self, images, settings, out=None):
        """images: uint8 [n_pairs, 2, h, w] (host).  Returns (supports, offsets[n+1], n_cand[n,2])."""
        images = np.ascontiguousarray(images, np.uint8)
        n_pairs, two, h, w = images.shape
        assert two == 2
        if out is None:
            out = np.empty(max(n_pairs * max(w - 26, 0) * max(h - 26, 0), 1), SUPPORT_DTYPE)
        offsets = np.zeros(n_pairs + 1, np.int64)
        n_cand = np.zeros((n_pairs, 2), np.int32)
        rc = self.lib.gpc_match_batch(self._h, _ptr(images), n_pairs, w, h, C.byref(settings), _ptr(out),
                                      C.c_int64(len(out)), _ptr(offsets), _ptr(n_cand))
        self._check(rc)
        return out[:offsets[-1]], offsets, n_cand

    def match_batch_raw(self, images_ptr, n_pairs, w, h, settings, out_ptr, cap, offsets_ptr, n_cand_ptr=None):
        """Pointer-level gpc_match_batch (pinned host buffers owned by the caller)."""
        self._check(self.lib.gpc_match_batch(self._h, C.c_void_p(images_ptr), n_pairs, w, h, C.byref(settings),
                                             C.c_void_p(out_ptr), C.c_int64(cap), C.c_void_p(offsets_ptr),
                                             C.c_void_p(n_cand_ptr or 0)))

    def match_batch_device(self, d_images, n_pairs, w, h, settings, d_out, cap_per_pair, d_n_out, d_n_cand=0):
        """Device pointers (ints); stream-ordered, no host synchronisation."""
        self._check(self.lib.gpc_match_batch_device(self._h, C.c_void_p(d_images), n_pairs, w, h, C.byref(settings),
                                                    C.c_void_p(d_out), C.c_int(cap_per_pair), C.c_void_p(d_n_out),
                                                    C.c_void_p(d_n_cand or 0)))

    def match_pyramid(self, left, right, n_levels, settings, cap=None):
        """Multi-level matching: (supports, level_offsets[n_levels+1], n_cand[n_levels,2])."""
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape
        cap = max(2 * (w - 26) * (h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), SUPPORT_DTYPE)
        offsets = np.zeros(n_levels + 1, np.int64)
        n_cand = np.zeros((n_levels, 2), np.int32)
        self._check(self.lib.gpc_match_pyramid(self._h, _ptr(left), _ptr(right), w, h, w, int(n_levels), C.byref(settings),
                                               _ptr(out), C.c_int64(cap), _ptr(offsets), _ptr(n_cand)))
        return out[:offsets[-1]].copy(), offsets, n_cand

    def match_pyramid_raw(self, left_ptr, right_ptr, w, h, n_levels, settings, out_ptr, cap, offsets_ptr, n_cand_ptr=None):
        """Pointer-level gpc_match_pyramid (host buffers owned by the caller, ideally pinned)."""
        self._check(self.lib.gpc_match_pyramid(self._h, C.c_void_p(left_ptr), C.c_void_p(right_ptr), w, h, w, int(n_levels),
                                               C.byref(settings), C.c_void_p(out_ptr), C.c_int64(cap), C.c_void_p(offsets_ptr),
                                               C.c_void_p(n_cand_ptr or 0)))

    # ---- stage seams ------------------------------------------------------------------------
    def preprocess(self, img, thr):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        smooth = np.empty((h, w), np.uint8)
        grad = np.empty((h, w), np.uint8)
        mask = np.empty(h * w, np.int32)
        n = C.c_int(0)
        self._check(self.lib.gpc_preprocess(self._h, _ptr(img), w, h, int(thr), _ptr(smooth), _ptr(grad), _ptr(mask),
                                            C.c_int(h * w), C.byref(n)))
        return smooth, grad, mask[:n.value].copy()

    def hash(self, img, thr, want_image=False):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        states = np.empty(h * w, np.uint32)
        mask = np.empty(h * w, np.int32)
        himg = np.empty((h, w), np.uint32) if want_image else None
        n = C.c_int(0)
        self._check(self.lib.gpc_hash(self._h, _ptr(img), w, h, int(thr), _ptr(states), _ptr(mask), C.c_int(h * w),
                                      C.byref(n), _ptr(himg)))
        if want_image:
            return states[:n.value].copy(), mask[:n.value].copy(), himg
        return states[:n.value].copy(), mask[:n.value].copy()

    def match_hash_images(self, hash_l, hash_r, settings):
        hash_l = np.ascontiguousarray(hash_l, np.uint32)
        hash_r = np.ascontiguousarray(hash_r, np.uint32)
        h, w = hash_l.shape
        cap = max(h * w, 1)
        out = np.empty(cap, SUPPORT_DTYPE)
        n = C.c_int(0)
        self._check(self.lib.gpc_match_hash_images(self._h, _ptr(hash_l), _ptr(hash_r), w, h, C.byref(settings),
                                                   _ptr(out), C.c_int(cap), C.byref(n)))
        return out[:n.value].copy()

    def hash_smooth(self, smooth, idx):
        """evalFastMaskOnSubsetSSE on a caller-provided smoothed image: one state per idx entry."""
        smooth = np.ascontiguousarray(smooth, np.uint8)
        idx = np.ascontiguousarray(idx, np.int32)
        h, w = smooth.shape
        states = np.zeros(max(len(idx), 1), np.uint32)
        self._check(self.lib.gpc_hash_smooth(self._h, _ptr(smooth), w, h, _ptr(idx), C.c_int(len(idx)), _ptr(states)))
        return states[:len(idx)].copy()

    def set_matcher(self, matcher):
        self._check(self.lib.gpc_set_matcher(self._h, int(matcher)))

    def find_correspondences(self, src_keys, tar_keys):
        """findCorrespondences on explicit 64-bit keys -> int32 [n, 2] (src index, tar index)."""
        src_keys = np.ascontiguousarray(src_keys, np.uint64)
        tar_keys = np.ascontiguousarray(tar_keys, np.uint64)
        cap = max(min(len(src_keys), len(tar_keys)), 1)
        out = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        self._check(self.lib.gpc_find_correspondences(self._h, _ptr(src_keys), C.c_int(len(src_keys)), _ptr(tar_keys),
                                                      C.c_int(len(tar_keys)), _ptr(out), C.c_int(cap), C.byref(n)))
        return out[:n.value].copy()

    def set_result_mode(self, naive):
        """False: the reference's default (SSE) build; True: its SSE=OFF build (the *Naive functions)."""
        self._check(self.lib.gpc_set_result_mode(self._h, C.c_int(1 if naive else 0)))

    def hashmatch(self, src_keys, tar_keys):
        """The reference's hashtable matcher (useHashtable) on explicit 64-bit keys -> int32 [n, 2]."""
        src_keys = np.ascontiguousarray(src_keys, np.uint64)
        tar_keys = np.ascontiguousarray(tar_keys, np.uint64)
        cap = max(min(len(src_keys), len(tar_keys)), 1)
        out = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        self._check(self.lib.gpc_hashmatch(self._h, _ptr(src_keys), C.c_int(len(src_keys)), _ptr(tar_keys),
                                           C.c_int(len(tar_keys)), _ptr(out), C.c_int(cap), C.byref(n)))
        return out[:n.value].copy()

    # ---- forests of more than 32 tests / 32-test forests in the naive result mode -------------------
    def set_wide_forest(self, tests):
        """tests: [n, 5] int32 rows of (ix, iy, jx, jy, tau) in file order, or a forest file path."""
        if isinstance(tests, (str, bytes, os.PathLike)):
            tests = read_forest_tests(tests)
        tests = np.ascontiguousarray(tests, np.int32).reshape(-1, 5)
        self._check(self.lib.gpc_set_wide_forest(self._h, _ptr(tests), C.c_int(len(tests))))

    def match_pair_wide(self, left, right, settings, cap=None):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape
        cap = max((w - 26) * (h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), SUPPORT_DTYPE)
        n, ncl, ncr = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self.lib.gpc_match_pair_wide(self._h, _ptr(left), _ptr(right), w, h, w, C.byref(settings), _ptr(out),
                                                 C.c_int(cap), C.byref(n), C.byref(ncl), C.byref(ncr)))
        return out[:n.value].copy(), ncl.value, ncr.value

    def hash_wide(self, img, thr, max_words=8):
        """uint32 [n_words, h, w]: candidate flag | state word per pixel under the wide forest."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        words = np.zeros((max_words, h, w), np.uint32)
        n = C.c_int(0)
        self._check(self.lib.gpc_hash_wide(self._h, _ptr(img), w, h, int(thr), _ptr(words), C.c_int(max_words), C.byref(n)))
        return words[:n.value].copy()

    # ---- resident images ----------------------------------------------------------------------
    def upload(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        handle = C.c_void_p()
        self._check(self.lib.gpc_image_upload(self._h, _ptr(img), w, h, w, C.byref(handle)))
        return ResidentImage(self, handle, w, h)

    def match_images(self, left, right, settings, cap=None):
        cap = max((left.w - 26) * (left.h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), SUPPORT_DTYPE)
        n, ncl, ncr = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self.lib.gpc_match_images(self._h, left.handle, right.handle, C.byref(settings), _ptr(out), C.c_int(cap),
                                              C.byref(n), C.byref(ncl), C.byref(ncr)))
        return out[:n.value].copy(), ncl.value, ncr.value

    def correspond_images(self, left, right, settings, cap=None):
        cap = max((left.w - 26) * (left.h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), CORR_DTYPE)
        n = C.c_int(0)
        self._check(self.lib.gpc_correspond_images(self._h, left.handle, right.handle, C.byref(settings), _ptr(out),
                                                   C.c_int(cap), C.byref(n)))
        return out[:n.value].copy()


class Pool:
    """gpc_pool: one resident context and host thread per listed device; match_batch has Context.match_batch's result."""

    def __init__(self, devices, max_w=1024, max_h=436, max_batch_per_device=1):
        self.lib = load_library()
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self.lib.gpc_pool_create(C.byref(self._h), devs, len(devices), int(max_w), int(max_h), int(max_batch_per_device))
        if rc != GPC_OK:
            raise GpcError(rc, self.lib.gpc_last_error(None).decode())
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.gpc_pool_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != GPC_OK:
            raise GpcError(rc, self.lib.gpc_pool_last_error(self._h).decode())

    @property
    def launches(self):
        return int(self.lib.gpc_pool_launch_count(self._h))

    def set_forest(self, forest):
        if isinstance(forest, (str, bytes, os.PathLike)):
            forest = read_forest(forest)
        self._check(self.lib.gpc_pool_set_forest(self._h, C.byref(forest)))

    def set_result_mode(self, naive):
        self._check(self.lib.gpc_pool_set_result_mode(self._h, 1 if naive else 0))

    def set_matcher(self, matcher):
        for i in range(len(self.devices)):
            rc = self.lib.gpc_set_matcher(C.c_void_p(self.lib.gpc_pool_context(self._h, i)), int(matcher))
            if rc != GPC_OK:
                raise GpcError(rc, "gpc_set_matcher")

    def match_batch_raw(self, images_ptr, n_pairs, w, h, settings, out_ptr, cap, offsets_ptr, n_cand_ptr=None):
        self._check(self.lib.gpc_pool_match_batch(self._h, C.c_void_p(images_ptr), n_pairs, w, h, C.byref(settings),
                                                  C.c_void_p(out_ptr), C.c_int64(cap), C.c_void_p(offsets_ptr),
                                                  C.c_void_p(n_cand_ptr or 0)))

    def match_batch(self, images, settings, cap=None):
        """images uint8 [n_pairs, 2, h, w] -> (supports, offsets[n_pairs+1], n_cand[n_pairs,2])."""
        images = np.ascontiguousarray(images, np.uint8)
        n_pairs, _, h, w = images.shape
        cap = max(n_pairs * (w - 26) * (h - 26), 1) if cap is None else cap
        out = np.empty(max(cap, 1), SUPPORT_DTYPE)
        offsets = np.zeros(n_pairs + 1, np.int64)
        n_cand = np.zeros((n_pairs, 2), np.int32)
        self.match_batch_raw(images.ctypes.data, n_pairs, w, h, settings, out.ctypes.data, cap, offsets.ctypes.data, n_cand.ctypes.data)
        return out[:offsets[-1]].copy(), offsets, n_cand


class ResidentImage:
    """A raw image kept on the context's device (gpc_image_upload / gpc_image_release)."""

    def __init__(self, ctx, handle, w, h):
        self.ctx, self.handle, self.w, self.h = ctx, handle, w, h

    def preprocess(self, thr, images=True):
        """preprocessImage on the resident image.  images=False: candidate list only (smooth / grad stay on the
        device until `fetch`); either way the image keeps the kernels' outputs for match_images."""
        smooth = np.empty((self.h, self.w), np.uint8) if images else None
        grad = np.empty((self.h, self.w), np.uint8) if images else None
        mask = np.empty(self.h * self.w, np.int32)
        n = C.c_int(0)
        self.ctx._check(self.ctx.lib.gpc_image_preprocess(self.ctx._h, self.handle, int(thr), _ptr(smooth), _ptr(grad),
                                                          _ptr(mask), C.c_int(self.h * self.w), C.byref(n)))
        return smooth, grad, mask[:n.value].copy()

    def fetch(self, thr):
        """(smooth, grad) of preprocessImage, computed on demand (gpc_image_fetch)."""
        smooth = np.empty((self.h, self.w), np.uint8)
        grad = np.empty((self.h, self.w), np.uint8)
        self.ctx._check(self.ctx.lib.gpc_image_fetch(self.ctx._h, self.handle, int(thr), _ptr(smooth), _ptr(grad)))
        return smooth, grad

    def release(self):
        if self.handle is not None and self.handle.value:
            self.ctx.lib.gpc_image_release(self.handle)
            self.handle = None

    __del__ = release
