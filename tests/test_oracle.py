"""CPU tests: the C restatement (oracle/) against the golden vectors generated from the
compiled, unmodified reference (scripts/make_golden.py) and, where oracle/_ref is present,
against the reference itself."""
import numpy as np
import pytest

from helpers import FORESTS, make_pair, supp_to_i32, write_forest
from oraclelib import digest, settings


def test_synth_generators_agree(oracle):
    from opengpc_b200.synth import synth_pair
    for w, h, seed in ((1024, 436, 1234), (256, 64, 7), (480, 270, 1300)):
        Lc, Rc = oracle.synth(w, h, seed)
        Lp, Rp = synth_pair(w, h, seed)
        assert np.array_equal(Lc, Lp) and np.array_equal(Rc, Rp)


def test_struct_sizes(golden):
    # sizeof(Descriptor)=24, Support=12, Correspondence=16
    assert golden["sizeof"] == {"Descriptor": 24, "Support": 12, "Correspondence": 16}


def test_matcher_kats(oracle, golden):
    for kat in golden["kats"]:
        got = oracle.find_correspondences(np.array(kat["src"], np.uint64), np.array(kat["tar"], np.uint64))
        assert got.tolist() == kat["pairs"], kat


def test_survey_kats(oracle):
    # SURVEY.md 8c, src x = index, tar x = index (100 + j in the survey)
    fc = lambda s, t: oracle.find_correspondences(np.array(s, np.uint64), np.array(t, np.uint64)).tolist()
    assert fc([1, 2, 3], [1, 2, 3, 9]) == [[0, 0], [1, 1], [2, 2]]
    assert fc([1, 2, 3], [1, 2, 3]) == [[0, 0], [1, 1]]            # last tar element never matches
    assert fc([1, 2, 2, 3], [1, 2, 3, 9]) == [[0, 0], [3, 2]]
    assert fc([1, 2, 3], [1, 2, 2, 3, 9]) == [[0, 0], [2, 3]]
    assert fc([1, 5], [1, 5, 5]) == [[0, 0], [1, 1]]               # tail-dup-2 quirk
    assert fc([1, 5], [1, 5, 5, 5]) == [[0, 0]]
    assert fc([1], [1]) == []
    assert fc([1], []) == []


def test_forest_reader(oracle):
    f = oracle.read_forest(FORESTS["tau"])
    assert (f.n_tests, f.type, f.n_discarded) == (30, 1, 0)
    assert (f.ix[0], f.iy[0], f.jx[0], f.jy[0], f.tau[0]) == (0, 3, -3, -2, 1)
    assert (f.ix[29], f.iy[29], f.jx[29], f.jy[29], f.tau[29]) == (5, -7, 0, 4, 9)
    z = oracle.read_forest(FORESTS["zero"])
    assert (z.n_tests, z.type) == (30, 0)
    d = oracle.read_forest(FORESTS["deep"])
    assert (d.n_tests, d.type, d.n_discarded) == (32, 1, 160)
    with pytest.raises(FileNotFoundError):
        oracle.read_forest("/nonexistent/forest.txt")


def test_forest_reader_vs_reference(oracle, reference):
    for name, path in FORESTS.items():
        f = oracle.read_forest(path)
        off, tau, typ = reference.read_forest(path, 1024, 436)
        assert typ == f.type and len(off) == 2 * f.n_tests
        for t in range(f.n_tests):
            assert off[2 * t] == f.ix[t] + f.iy[t] * 1024 and off[2 * t + 1] == f.jx[t] + f.jy[t] * 1024
            if typ == 1:
                assert tau[t] == f.tau[t]


def test_small_cases_all_stages(oracle, small_cases):
    """Every intermediate of every small fixture, bit for bit."""
    for name, c in small_cases.items():
        thr, epi, vt, dh = (int(v) for v in c["cfg"])
        path = write_forest(c["forest"])
        f = oracle.read_forest(path)
        st = {}
        for side in "LR":
            sm, gr, mk, states = oracle.stages(c[side], f, thr)
            assert np.array_equal(sm, c["smooth" + side]), (name, side, "smooth")
            # grad columns 0,1 read the byte in front of the row in the reference (not canonical)
            assert np.array_equal(gr[:, 2:], c["grad" + side][:, 2:]), (name, side, "grad")
            assert np.array_equal(mk, c["mask" + side]), (name, side, "mask")
            assert np.array_equal(states, c["states" + side]), (name, side, "states")
            st[side] = (mk, states)
        supp = oracle.match(st["L"][0], st["L"][1], st["R"][0], st["R"][1], c["L"].shape[1],
                            settings(thr, dh, vt, epi))
        assert np.array_equal(supp_to_i32(supp), c["supp"].reshape(-1, 3)), (name, "supports")


@pytest.mark.parametrize("idx", range(16))
def test_golden_pairs(oracle, golden, idx):
    """BASELINE shapes: candidate counts, support count and ordered-list digest."""
    if idx >= len(golden["pairs"]):
        pytest.skip("no such golden record")
    rec = golden["pairs"][idx]
    L, R = make_pair(rec)
    f = oracle.read_forest(FORESTS[rec["forest"]])
    supp, ncl, ncr = oracle.pair(L, R, f, settings(5, rec["disp_high"], rec["vt"], rec["epipolar"]))
    assert (ncl, ncr, len(supp)) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"])
    assert "%016x" % oracle.digest(supp) == rec["digest"]


def test_python_digest_matches_c(oracle):
    L, R = oracle.synth(256, 64, 3)
    f = oracle.read_forest(FORESTS["zero"])
    supp, _, _ = oracle.pair(L, R, f, settings())
    assert digest(supp) == oracle.digest(supp)


def test_oracle_vs_reference_random(oracle, reference):
    """Fresh random images / thresholds / settings against the compiled reference."""
    rng = np.random.default_rng(5)
    for it in range(6):
        w = int(rng.integers(4, 20)) * 16
        h = int(rng.integers(30, 90))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if it % 2:
            img = (img // 16 * 16).astype(np.uint8)
        R = np.roll(img, -int(rng.integers(0, 9)), axis=1)
        thr = int(rng.choice([0, 3, 5, 10, 40, 181, 182, 255]))
        epi = bool(it % 3)
        fname = ["tau", "zero", "deep"][it % 3]
        sm, gr, mk = reference.preprocess(img, thr)
        f = oracle.read_forest(FORESTS[fname])
        osm, ogr, omk, ost = oracle.stages(img, f, thr)
        assert np.array_equal(sm, osm) and np.array_equal(gr[:, 2:], ogr[:, 2:]) and np.array_equal(mk, omk)
        assert np.array_equal(reference.hash(img, thr, FORESTS[fname]), ost)
        rs, ncl, ncr, _ = reference.pair(img, R, FORESTS[fname], thr=thr, disp_high=50, vt=1, epipolar=epi)
        os_, ocl, ocr = oracle.pair(img, R, f, settings(thr, 50, 1, epi))
        assert (ncl, ncr) == (ocl, ocr)
        assert np.array_equal(supp_to_i32(rs), supp_to_i32(os_)), (it, len(rs), len(os_))
