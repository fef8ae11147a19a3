"""CPU: pin oracle/smle_oracle.c against (a) the known-answer values SURVEY.md section 4 records
from the reference, (b) the golden fixture generated from the compiled reference."""
import numpy as np
import pytest

from conftest import named_matrix, rel_rownorm_err
from oracle import oracle as O


# ---- SURVEY.md section 4 known answers ----------------------------------------------------------
def test_survey_merge_coords_wheel10(orc):
    ro, ci, va = orc.gen_wheel(10)
    assert ro.tolist() == [0, 10] + list(range(11, 21))
    assert orc.merge_partition(ro, 4).tolist() == [[0, 0], [0, 8], [3, 13], [7, 17], [11, 20]]


def test_survey_merge_coords_dense(orc):
    ro, _, _ = orc.gen_dense(4, 3)
    assert orc.merge_partition(ro, 3).tolist() == [[0, 0], [1, 5], [3, 9], [4, 12]]


def test_survey_merge_coords_wheel_1m(orc):
    ro, _, _ = orc.gen_wheel(1000000)
    want = [[0, 0], [0, 375001], [0, 750002], [62502, 1062501], [250002, 1250002], [437503, 1437502],
            [625003, 1625003], [812504, 1812503], [1000001, 2000000]]
    assert orc.merge_partition(ro, 8).tolist() == want


def test_survey_merge_coords_grid2d_1000(orc):
    ro, ci, _ = orc.gen_grid2d(1000, True)
    assert len(ci) == 4996000
    want = [[0, 0], [125125, 624375], [250083, 1248917], [375041, 1873459], [500000, 2498000],
            [624958, 3122542], [749916, 3747084], [874874, 4371626], [1000000, 4996000]]
    assert orc.merge_partition(ro, 8).tolist() == want


def test_survey_merge_coords_grid3d_24(orc):
    ro, ci, _ = orc.gen_grid3d(24, True)
    assert (len(ro) - 1, len(ci)) == (13824, 93312)
    want = [[0, 0], [1785, 11607], [3495, 23289], [5204, 34972], [6912, 46656], [8619, 58341],
            [10328, 70024], [12038, 81706], [13824, 93312]]
    assert orc.merge_partition(ro, 8).tolist() == want


def test_survey_spmm_wheel10(orc):
    ro, ci, va = orc.gen_wheel(10)
    X = (np.arange(22) + 1.0).reshape(11, 2)
    Y = orc.merge_csrmm(4, ro, ci, va, X, 2)
    want = [120, 130] + [v for i in range(2, 11) for v in (2 * i + 1, 2 * i + 2)] + [3, 4]
    assert Y.ravel().tolist() == [float(v) for v in want]


@pytest.mark.parametrize("w,k,raw,thr_iters,single_raw,single_thr,thr", [
    (24, 4, 64, None, 64, None, None),
    (50, 4, 129, 72, 121, 72, 2.043200e-03),
])
def test_survey_cg_iteration_counts(orc, w, k, raw, thr_iters, single_raw, single_thr, thr):
    ro, ci, va = orc.gen_grid3d(w, True, 6.0, -1.0)
    n = len(ro) - 1
    B = orc.rhs_rand(42, n * k).reshape(n, k)
    it, _, hist = orc.cg_multi(ro, ci, va, B, k, 10000, 1e-5, O.MERGE, 8)
    assert it == raw and len(hist) == raw
    if w == 24:
        b0 = np.ascontiguousarray(B[:, 0])          # SURVEY: column 0 of the row-major block
    else:
        b0 = np.ascontiguousarray(B.ravel()[:n])    # SURVEY: B[0:n]
    assert orc.cg_single(ro, ci, va, b0, 10000, 1e-5)[0] == single_raw
    if thr is not None:
        t = orc.driver_threshold(B.ravel(), n, 1e-5)
        assert abs(t - thr) / thr < 1e-6
        assert orc.cg_multi(ro, ci, va, B, k, 10000, t, O.MERGE, 8)[0] == thr_iters
        assert orc.cg_single(ro, ci, va, b0, 10000, t)[0] == single_thr


# ---- golden fixture from the compiled reference ---------------------------------------------------
def test_golden_partition(orc, golden):
    cache = {}
    for rec in golden["partition"]:
        ro = cache.setdefault(rec["matrix"], named_matrix(orc, rec["matrix"]))[0]
        assert orc.merge_partition(ro, rec["threads"]).tolist() == rec["coords"], rec["matrix"]


def test_golden_spmv(orc, golden):
    for rec in golden["spmv"]:
        ro, ci, va = named_matrix(orc, rec["matrix"])
        x = np.array(rec["x"])
        y = orc.merge_csrmv(rec["threads"], ro, ci, va, x)
        assert rel_rownorm_err(y, np.array(rec["y"])) <= 1e-14
        assert rel_rownorm_err(orc.spmv_gold(ro, ci, va, x), np.array(rec["y_gold"])) <= 1e-14


def test_golden_spmm(orc, golden):
    for rec in golden["spmm"]:
        ro, ci, va = named_matrix(orc, rec["matrix"])
        X = np.array(rec["X"]).reshape(-1, rec["k"])
        Y = orc.merge_csrmm(rec["threads"], ro, ci, va, X, rec["k"])
        assert rel_rownorm_err(Y, np.array(rec["Y"]).reshape(-1, rec["k"])) <= 1e-14


def test_golden_cg(orc, golden):
    for rec in golden["cg"]:
        w, k = rec["grid3d"], rec["k"]
        ro, ci, va = orc.gen_grid3d(w, True, 6.0, -1.0)
        n = len(ro) - 1
        B = orc.rhs_rand(42, n * k).reshape(n, k)
        it, X, hist = orc.cg_multi(ro, ci, va, B, k, 10000, rec["tol"], O.MERGE, 8)
        assert it == rec["multi_iters"]
        assert abs(X.sum() - rec["multi_x_sum"]) <= 1e-9 * max(1.0, abs(rec["multi_x_sum"]))
        np.testing.assert_allclose(hist[-3:], rec["multi_hist_tail"], rtol=1e-6)
        it1, x1 = orc.cg_single(ro, ci, va, np.ascontiguousarray(B.ravel()[:n]), 10000, rec["tol"])
        assert it1 == rec["single_iters_on_flat_b0"]


# ---- internal consistency of the three SpmmKernel choices ---------------------------------------------
@pytest.mark.parametrize("T", [1, 3, 8])
def test_spmm_kernels_agree(orc, T):
    ro, ci, va = orc.gen_grid2d(9, True, 4.0, -1.0)
    X = np.random.default_rng(1).random((len(ro) - 1, 3))
    a = orc.merge_csrmm(T, ro, ci, va, X, 3)
    b = orc.nonzero_split_csrmm(T, ro, ci, va, X, 3)
    c = orc.row_split_csrmm(T, ro, ci, va, X, 3)
    assert rel_rownorm_err(a, c) <= 1e-14 and rel_rownorm_err(b, c) <= 1e-14


def test_merge_csrmv_rejects_more_than_256_threads(orc):
    ro, ci, va = orc.gen_wheel(10)
    with pytest.raises(ValueError):
        orc.merge_csrmv(257, ro, ci, va, np.ones(11))
