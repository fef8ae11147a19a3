"""GPU: CG through the C ABI against the oracle / golden fixture.
Bar (north_star): same residual tolerance reached within +-2 % of the reference's iteration count."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _close_iters(got, want):
    return abs(got - want) <= max(1, round(0.02 * want))


def test_golden_cg_counts(gpu, golden):
    for rec in golden["cg"]:
        w, k = rec["grid3d"], rec["k"]
        ro, ci, va = gpu.gen_grid3d(w, True, 6.0, -1.0)
        n = len(ro) - 1
        a = gpu.CsrMatrix(ro, ci, va)
        B = gpu.gen_rhs_rand(42, n * k).reshape(n, k)
        it, X, hist, rel = a.cg_solve_multiple(B, 10000, rec["tol"])
        assert _close_iters(it, rec["multi_iters"]), (rec, it)
        assert len(hist) == it and rel < rec["tol"]
        assert abs(X.sum() - rec["multi_x_sum"]) <= 1e-5 * max(1.0, abs(rec["multi_x_sum"]))
        it1, x1, rel1 = a.cg_solve_single(np.ascontiguousarray(B.ravel()[:n]), 10000, rec["tol"])
        assert _close_iters(it1, rec["single_iters_on_flat_b0"]), (rec, it1)
        a.close()


@pytest.mark.parametrize("k", [1, 2, 3, 8, 32])
def test_multi_rhs_against_oracle(gpu, orc, k):
    ro, ci, va = gpu.gen_grid3d(20, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    B = gpu.gen_rhs_rand(42, n * k).reshape(n, k)
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, 1e-8)
    it_o, X_o, hist_o = orc.cg_multi(ro, ci, va, B, k, 10000, 1e-8, O.MERGE, 8)
    assert _close_iters(it, it_o)
    np.testing.assert_allclose(X, X_o, rtol=1e-6, atol=1e-9)
    nh = min(len(hist), len(hist_o)) - 2
    np.testing.assert_allclose(hist[:nh], hist_o[:nh], rtol=1e-5)
    # true residual of the returned solution
    for c in range(k):
        r = B[:, c] - orc.spmv_gold(ro, ci, va, np.ascontiguousarray(X[:, c]))
        assert np.linalg.norm(r) / np.linalg.norm(B[:, c]) < 1e-7
    a.close()


def test_single_rhs_against_oracle_2d(gpu, orc):
    ro, ci, va = gpu.gen_grid2d(64, True, 4.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    b = gpu.gen_rhs_rand(7, n)
    it, x, rel = a.cg_solve_single(b, 10000, 1e-9)
    it_o, x_o = orc.cg_single(ro, ci, va, b, 10000, 1e-9)
    assert _close_iters(it, it_o)
    np.testing.assert_allclose(x, x_o, rtol=1e-6, atol=1e-9)
    a.close()


def _arrow_spd(n, fan, dtype=np.float64):
    """SPD matrix with skewed rows: row 0 is a hub coupled to every vertex (spans several tiles), every
    `fan`-th row is a medium hub coupled to the `fan` rows behind it, plus a tridiagonal part; strictly
    diagonally dominant.  Exercises the general tiles of the single-vector kernel inside CG (fused p.Ap)."""
    rows, cols, vals = [], [], []

    def add(i, j, v):
        rows.extend([i, j]); cols.extend([j, i]); vals.extend([v, v])

    for j in range(1, n):
        add(0, j, -1.0 / n)
    for i in range(1, n - 1):
        add(i, i + 1, -0.5)
    for h in range(fan, n - fan, fan):
        for j in range(h + 2, h + fan):
            add(h, j, -0.25 / fan)
    import scipy.sparse as sp
    a = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    a = a + sp.diags(np.asarray(abs(a).sum(axis=1)).ravel() + 1.0)
    a = a.tocsr(); a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(dtype)


def test_single_rhs_cg_on_skewed_spd_matrix(gpu, orc):
    ro, ci, va = _arrow_spd(20000, 200)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    b = gpu.gen_rhs_rand(5, n)
    # the product alone first: hub row, medium rows and short rows against the gold loop
    y = a.spmv(b)
    y_o = orc.spmv_gold(ro, ci, va, b)
    np.testing.assert_allclose(y, y_o, rtol=1e-12, atol=1e-13)
    it, x, rel = a.cg_solve_single(b, 10000, 1e-10)
    it_o, x_o = orc.cg_single(ro, ci, va, b, 10000, 1e-10)
    assert _close_iters(it, it_o), (it, it_o)
    np.testing.assert_allclose(x, x_o, rtol=1e-7, atol=1e-10)
    a.close()


def test_max_iters_cap_and_zero_iters(gpu):
    ro, ci, va = gpu.gen_grid3d(10, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    B = gpu.gen_rhs_rand(1, n * 2).reshape(n, 2)
    for cap in (1, 5, 16, 17, 33):
        it, X, hist, rel = a.cg_solve_multiple(B, cap, 1e-30)
        assert it == cap and len(hist) == cap      # not converged -> iterations == max_iters
    it, X, hist, rel = a.cg_solve_multiple(B, 0, 1e-5)
    assert it == 0 and np.array_equal(X, np.zeros_like(X))
    a.close()


def test_per_column_latch_freezes_converged_columns(gpu, orc):
    """columns converge at different iterations; latched ones must stop moving (alpha=beta=0)."""
    ro, ci, va = gpu.gen_grid3d(12, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    B = gpu.gen_rhs_rand(3, n * 3).reshape(n, 3).copy()
    B[:, 1] = 0.0
    B[0, 1] = 1.0                                  # a very different right-hand side
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, 1e-6)
    it_o, X_o, hist_o = orc.cg_multi(ro, ci, va, B, 3, 10000, 1e-6, O.MERGE, 8)
    assert _close_iters(it, it_o)
    np.testing.assert_allclose(X, X_o, rtol=1e-5, atol=1e-9)
    a.close()


def test_device_pointer_solve_and_driver_threshold(gpu, orc):
    import torch
    ro, ci, va = gpu.gen_grid3d(30, True, 6.0, -1.0)
    n, k = len(ro) - 1, 4
    a = gpu.CsrMatrix(ro, ci, va)
    Bh = gpu.gen_rhs_rand(42, n * k).reshape(n, k)
    thr = gpu.driver_threshold(Bh.ravel(), n, 1e-5)          # cpu_multicg.cpp:168 semantics
    B = torch.from_numpy(Bh).cuda()
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, thr)
    it_o, X_o, _ = orc.cg_multi(ro, ci, va, Bh, k, 10000, thr, O.MERGE, 8)
    assert _close_iters(it, it_o)
    np.testing.assert_allclose(X.cpu().numpy(), X_o, rtol=1e-5, atol=1e-9)
    a.close()


def test_cg_at_baseline_size_converges(gpu):
    """config 2 (150^3) property check: converged solution has a small TRUE residual."""
    import torch
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    b = torch.from_numpy(gpu.gen_rhs_rand(42, n)).cuda()
    it, x, rel = a.cg_solve_single(b, 10000, 1e-8)
    r = b - a.spmv(x)
    gpu.sync()
    true_rel = (r.norm() / b.norm()).item()
    assert 0 < it < 2000 and rel < 1e-8 and true_rel < 2e-8, (it, rel, true_rel)
    a.close()


def test_single_batch_matches_one_by_one(gpu, orc):
    """smle_cg_single_batch_f64 (the solve loop of TestCGSolveSingle with overlapped host copies) gives
    exactly the iterates of L separate smle_cg_single_f64 calls."""
    ro, ci, va = gpu.gen_grid3d(30, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    L = 5
    B = gpu.gen_rhs_rand(42, n * L).reshape(L, n).copy()
    its, X = a.cg_solve_single_batch(B, 10000, 1e-7)
    for v in range(L):
        it1, x1, _ = a.cg_solve_single(B[v], 10000, 1e-7)
        assert it1 == its[v]
        assert np.array_equal(x1, X[v]), v
    it_ref, x_ref = orc.cg_single(ro, ci, va, B[L - 1], 10000, 1e-7)
    assert abs(its[-1] - it_ref) <= max(1, round(0.02 * it_ref))
    assert np.abs(X[-1] - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    its0, _ = a.cg_solve_single_batch(B[:0], 10, 1e-7)
    assert its0 == []
    a.close()
