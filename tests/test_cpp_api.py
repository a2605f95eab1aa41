"""The drop-in C++ API (include/gpc/{inference,buffer}.hpp): builds on CPU, results on the GPU."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import FORESTS
from oraclelib import ROOT, settings as osettings

API_TEST = os.path.join(ROOT, "tests", "cpp", "api_test")
SPARSEMATCH = os.path.join(ROOT, "samples", "sparsematch")
C_EXAMPLE = os.path.join(ROOT, "samples", "c_api_example")


def _build():
    from opengpc_b200.build import build_native
    build_native()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "samples")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def _write_png(path, arr):
    from PIL import Image
    Image.fromarray(arr).save(path)


def test_cpp_host_builds():
    """sparsematch, api_test and the PNG codec compile against the drop-in headers; the reference's
    own samples/sparsematch.cpp compiles UNCHANGED against them where the tree is present."""
    _build()
    assert os.path.exists(API_TEST) and os.path.exists(SPARSEMATCH) and os.path.exists(C_EXAMPLE)
    ref_src = "/root/reference/samples/sparsematch.cpp"
    if os.path.exists(ref_src):
        with tempfile.TemporaryDirectory() as d:
            r = subprocess.run(["g++", "-std=c++11", "-w", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "sm"), ref_src,
                                "-L" + os.path.join(ROOT, "opengpc_b200"), "-lgpc_b200", "-lz", "-lpthread"],
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr


def test_png_codec_roundtrip():
    """readPNG: gray8 as is, RGB -> (r+g+b)/3, gray16 truncated to the low byte (buffer.hpp:280-299),
    width padded to a multiple of 16; writePNG output is readable by an independent decoder."""
    from PIL import Image
    rng = np.random.default_rng(1)
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "png_test")
        r = subprocess.run(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe,
                            os.path.join(ROOT, "tests", "cpp", "png_test.cpp"), "-lz"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        gray = rng.integers(0, 256, (37, 50), dtype=np.uint8)
        rgb = rng.integers(0, 256, (20, 33, 3), dtype=np.uint8)
        g16 = rng.integers(0, 65536, (9, 16), dtype=np.uint16)
        cases = {"gray": (gray, gray), "rgb": (rgb, (rgb.astype(np.int32).sum(2) // 3).astype(np.uint8)),
                 "g16": (g16, (g16 & 0xff).astype(np.uint8))}
        for name, (src, want) in cases.items():
            pin, pout, praw = (os.path.join(d, f"{name}{e}") for e in (".png", "_out.png", ".raw"))
            if name == "g16":
                Image.fromarray(src).save(pin)        # mode I;16
            else:
                Image.fromarray(src).save(pin)
            r = subprocess.run([exe, pin, pout, praw], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, (name, r.returncode, r.stdout)
            raw = np.fromfile(praw, np.uint8)
            w, h, cols = np.frombuffer(raw[:12].tobytes(), np.int32)
            assert (w, h) == (want.shape[1], want.shape[0]) and cols % 16 == 0
            buf = raw[12:].reshape(h, cols)
            assert np.array_equal(buf[:, :w], want), name
            assert not buf[:, w:].any()
            back = np.array(Image.open(pout))
            assert np.array_equal(back, want), name


@pytest.mark.gpu
@pytest.mark.parametrize("forest,epipolar,vt,dh,thr,w,h,naive", [("tau", 1, 0, 128, 5, 1024, 436, 0), ("zero", 1, 0, 128, 5, 640, 200, 0),
                                                                ("tau", 0, 1, 128, 10, 512, 160, 0), ("deep", 1, 0, 64, 5, 500, 130, 0),
                                                                ("tau", 1, 0, 128, 5, 640, 200, 1), ("zero", 0, 1, 128, 10, 500, 130, 1)])
def test_cpp_api_vs_oracle(oracle, forest, epipolar, vt, dh, thr, w, h, naive):
    """preprocessImage / rectifiedMatch / stereoMatch / evalFastMaskOnSubsetSSE / findCorrespondences through
    the C++ headers, on PNG inputs (width 500 exercises the 16-pixel padding), against the oracle.  naive: the same
    program compiled with -DGPC_B200_NAIVE_RESULTS against the oracle's SSE=OFF restatement."""
    from opengpc_b200.synth import synth_pair
    _build()
    wa = (w + 15) // 16 * 16
    L, R = synth_pair(wa, h, 4321)
    L, R = np.ascontiguousarray(L[:, :w]), np.ascontiguousarray(R[:, :w])
    Lp, Rp = np.zeros((h, wa), np.uint8), np.zeros((h, wa), np.uint8)      # what readPNG hands the library
    Lp[:, :w], Rp[:, :w] = L, R
    of = oracle.read_forest(FORESTS[forest])
    s = osettings(thr, dh, vt, bool(epipolar))
    with tempfile.TemporaryDirectory() as d:
        pl, pr, pout = (os.path.join(d, n) for n in ("l.png", "r.png", "out.bin"))
        _write_png(pl, L); _write_png(pr, R)
        r = subprocess.run([API_TEST + ("_naive" if naive else ""), FORESTS[forest], pl, pr, pout, str(epipolar), str(vt), str(dh), str(thr)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
        if forest == "deep":
            assert r.stdout.count("Note: A maximum of 32 fern features") == 160
        assert "number of ferns:" in r.stdout
        v = np.fromfile(pout, np.int32)
    nL, nR, nS, nC, nS2, nD, nH = v[:7]
    p = 7
    maskL = v[p:p + nL]; p += nL
    maskR = v[p:p + nR]; p += nR
    supp = v[p:p + 3 * nS].reshape(-1, 3); p += 3 * nS
    corr = v[p:p + 4 * nC].reshape(-1, 4); p += 4 * nC
    supp2 = v[p:p + 3 * nS2].reshape(-1, 3); p += 3 * nS2
    states = v[p:p + nD].view(np.uint32); p += nD
    supp_ht = v[p:p + 3 * nH].reshape(-1, 3)
    stages = oracle.stages_naive if naive else oracle.stages
    _, _, omkL, ostL = stages(Lp, of, thr)
    _, _, omkR, _ = stages(Rp, of, thr)
    assert np.array_equal(maskL, omkL) and np.array_equal(maskR, omkR)
    assert np.array_equal(states, ostL)
    if naive:
        ref, _, _ = oracle.pair_naive(Lp, Rp, of, s)
        ht, _, _ = oracle.pair_naive(Lp, Rp, of, s, use_hashtable=True)
    else:
        ref, _, _ = oracle.pair(Lp, Rp, of, s)
        ht = oracle.pair_hashtable(Lp, Rp, of, s)
        assert np.array_equal(corr, oracle.correspondences(Lp, Rp, of, s))
    want = np.stack([ref["x"], ref["y"], ref["d"].astype(np.int32)], 1)
    assert np.array_equal(supp, want)
    assert np.array_equal(supp2, want), "hand-built PreprocessedImage path differs"
    assert np.array_equal(supp_ht, np.stack([ht["x"], ht["y"], ht["d"].astype(np.int32)], 1)), "useHashtable(true) differs"


@pytest.mark.gpu
def test_cpp_api_threads_and_regrowth(oracle):
    """tests/cpp/thread_test.cpp: four host threads through the shared, locked device context give the
    single-threaded result; images preprocessed before a larger image forced a new context still match;
    lazily fetched smooth / grad images feed the hand-built path; the counts equal the oracle's."""
    from opengpc_b200.synth import synth_pair
    _build()
    exe = os.path.join(ROOT, "tests", "cpp", "thread_test")
    L, R = synth_pair(640, 200, 99)
    l, r = synth_pair(256, 96, 98)
    of = oracle.read_forest(FORESTS["tau"])
    want_big = len(oracle.pair(L, R, of, osettings())[0])
    want_small = len(oracle.pair(l, r, of, osettings())[0])
    with tempfile.TemporaryDirectory() as d:
        paths = [os.path.join(d, n) for n in ("L.png", "R.png", "l.png", "r.png")]
        for pth, im in zip(paths, (L, R, l, r)):
            _write_png(pth, im)
        res = subprocess.run([exe, FORESTS["tau"]] + paths + ["4", "6"], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, (res.returncode, res.stdout[-1500:], res.stderr[-1500:])
        assert f"ok {want_big} {want_small}" in res.stdout, res.stdout[-500:]


@pytest.mark.gpu
def test_sparsematch_cli(oracle):
    from opengpc_b200.synth import synth_pair
    _build()
    L, R = synth_pair(1024, 436, 1234)
    with tempfile.TemporaryDirectory() as d:
        pl, pr, po = (os.path.join(d, n) for n in ("l.png", "r.png", "disp.png"))
        _write_png(pl, L); _write_png(pr, R)
        r = subprocess.run([SPARSEMATCH, FORESTS["tau"], pl, pr, po, "3"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "#candidatesL:377030, #candidatesR:378097" in r.stdout and "num matches:40839" in r.stdout, r.stdout
        from PIL import Image
        assert np.array(Image.open(po)).shape == (436, 1024, 3)


@pytest.mark.gpu
def test_pool_from_plain_c(oracle):
    """samples/pool_example.c: gpc_pool_match_batch from C99 with gpc_host_alloc'ed buffers (two contexts on device 0),
    23 copies of one pair; the count equals the oracle's for that pair."""
    _build()
    w, h = 512, 120
    s = np.uint32(12345)
    vals = np.empty(w * h, np.uint8)
    x = 12345
    for i in range(w * h):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        vals[i] = x >> 24
    L = vals.reshape(h, w)
    R = np.roll(L, -7, axis=1)
    ref, _, _ = oracle.pair(L, R, oracle.read_forest(FORESTS["tau"]), osettings())
    r = subprocess.run([os.path.join(ROOT, "samples", "pool_example"), FORESTS["tau"], str(w), str(h), "23"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"ok {len(ref)}" in r.stdout, (r.stdout, len(ref))


@pytest.mark.gpu
def test_c_api_from_plain_c(oracle):
    """samples/c_api_example.c (C99, no C++ on the caller's side): the SURVEY 8c golden values of the Sintel-sized
    synthetic pair, and the SSE=OFF result mode against the oracle."""
    from oraclelib import digest
    from opengpc_b200.synth import synth_pair
    _build()
    L, R = synth_pair(1024, 436, 1234)
    with tempfile.TemporaryDirectory() as d:
        pl, pr = os.path.join(d, "l.raw"), os.path.join(d, "r.raw")
        L.tofile(pl); R.tofile(pr)
        r = subprocess.run([C_EXAMPLE, FORESTS["tau"], "1024", "436", pl, pr], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "#candidatesL:377030, #candidatesR:378097, num matches:40839, digest:fb3f3b2728785637" in r.stdout, r.stdout
        r = subprocess.run([C_EXAMPLE, FORESTS["zero"], "1024", "436", pl, pr, "naive"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        ref, ncl, ncr = oracle.pair_naive(L, R, oracle.read_forest(FORESTS["zero"]), osettings())
        assert f"#candidatesL:{ncl}, #candidatesR:{ncr}, num matches:{len(ref)}, digest:{digest(ref):016x}" in r.stdout, r.stdout
