"""GPU: parity with the oracle AT THE SIZES BASELINE.json names (not only through size-independent
properties): SpMV on grid2d 1000^2 (configs[0]) and grid3d 150^3, SpMM 150^3 x 32, the CG iteration
counts SURVEY.md section 4 recorded from the reference at 50^3 / 100^3, the 150^3 single-RHS solve
the round-1 headline was built on (configs[1]), the wheel 2^20 hub row (configs[3]), and -- without
an oracle, by the true residual -- the 300^3 system of configs[4].

Error metrics.  north_star asks for "1e-12 relative per row-norm (fp64)".  Two numbers are asserted:
  * rel_rownorm_err(..., csr, X): |y_i - yref_i| / (|A||x|)_i   -- the componentwise bound for a
    different summation order; always meaningful, asserted for every row;
  * rownorm_err: |y_i - yref_i| / ||yref_i||                     -- the stated metric; asserted
    for every row whose result is not a cancellation (||yref_i|| >= 1e-3 (|A||x|)_i), which on the
    all-positive grid2d matrix of the reference driver (values 1.0) is every row.
"""
import numpy as np
import pytest

from conftest import rel_rownorm_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rownorm_err(Y, Y_ref, csr, X):
    """(max over non-cancelling rows of |y_i - yref_i| / ||yref_i||, fraction of rows that qualify)"""
    import scipy.sparse as sp
    ro, ci, va = csr
    R = np.asarray(Y_ref, dtype=np.float64).reshape(len(ro) - 1, -1)
    Yv = np.asarray(Y, dtype=np.float64).reshape(R.shape)
    Xa = np.abs(np.asarray(X, dtype=np.float64)).reshape(-1, R.shape[1])
    A = sp.csr_matrix((np.abs(va.astype(np.float64)), ci, ro), shape=(len(ro) - 1, Xa.shape[0]))
    scale = np.linalg.norm(A @ Xa, axis=1)
    norm = np.linalg.norm(R, axis=1)
    ok = norm >= 1e-3 * np.maximum(scale, 1e-300)
    err = np.abs(Yv - R).max(axis=1)
    return float((err[ok] / norm[ok]).max()) if ok.any() else 0.0, float(ok.mean())


def _close_iters(got, want):
    return abs(got - want) <= max(1, round(0.02 * want))


# ---- configs[0]: cpu_spmv merge SpMV on grid2d 1000^2 ------------------------------------------
@pytest.mark.parametrize("self_loop,diag,offd", [(False, 1.0, 1.0), (True, 4.0, -1.0)])
def test_spmv_grid2d_1000_against_oracle(gpu, orc, self_loop, diag, offd):
    ro, ci, va = gpu.gen_grid2d(1000, self_loop, diag, offd)
    assert len(ci) == (4996000 if self_loop else 3996000)          # C1 sizes
    a = gpu.CsrMatrix(ro, ci, va)
    n = len(ro) - 1
    rng = np.random.default_rng(1)
    for x in (np.full(n, 0.0019), rng.random(n)):                  # the driver's x (cpu_spmv.cpp:855) and a random one
        y = a.spmv(x)
        y_gold = orc.spmv_gold(ro, ci, va, x)                       # cpu_spmv.cpp:862 gold check
        y_merge = orc.merge_csrmv(8, ro, ci, va, x)                 # OmpMergeCsrmv, 8 threads
        for ref in (y_gold, y_merge):
            assert rel_rownorm_err(y, ref, (ro, ci, va), x) <= 1e-12
            e, frac = rownorm_err(y, ref, (ro, ci, va), x)
            assert e <= 1e-12
            if not self_loop:
                assert frac == 1.0                                  # all-positive matrix: no row cancels
    a.close()


# ---- grid3d 150^3 (configs[1] matrix): SpMV and SpMM x 32 ---------------------------------------
def test_spmv_and_spmm32_grid3d_150_against_oracle(gpu, orc):
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    assert (n, len(ci)) == (3375000, 23490000)
    a = gpu.CsrMatrix(ro, ci, va)
    rng = np.random.default_rng(2)
    x = rng.random(n)
    y = a.spmv(x)
    y_ref = orc.merge_csrmv(8, ro, ci, va, x)
    assert rel_rownorm_err(y, y_ref, (ro, ci, va), x) <= 1e-12
    assert rownorm_err(y, y_ref, (ro, ci, va), x)[0] <= 1e-12
    X = rng.random((n, 32))
    Y = a.spmm(X)
    Y_ref = orc.merge_csrmm(8, ro, ci, va, X, 32)                   # OmpMergeCsrmm (merge_based.hpp:49-153)
    assert rel_rownorm_err(Y, Y_ref, (ro, ci, va), X) <= 1e-12
    e, frac = rownorm_err(Y, Y_ref, (ro, ci, va), X)
    assert e <= 1e-12 and frac > 0.99
    a.close()


# ---- CG iteration counts SURVEY.md section 4 recorded from the reference --------------------------
@pytest.mark.parametrize("w,multi_raw,multi_thr,single_raw,single_thr,threshold",
                         [(50, 129, 72, 121, 72, 2.043200e-03), (100, 239, 133, 237, 133, 5.774754e-03)])
def test_cg_known_iteration_counts(gpu, w, multi_raw, multi_thr, single_raw, single_thr, threshold):
    ro, ci, va = gpu.gen_grid3d(w, True, 6.0, -1.0)
    n, k = len(ro) - 1, 4
    a = gpu.CsrMatrix(ro, ci, va)
    flat = gpu.gen_rhs_rand(42, n * k)
    B = flat.reshape(n, k)
    thr = gpu.driver_threshold(flat, n, 1e-5)
    assert abs(thr - threshold) <= 1e-6 * threshold
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, 1e-5)          # CGSolveMultiple(MERGE), raw tolerance
    assert _close_iters(it, multi_raw), it
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, thr)           # ... the drivers' threshold semantics
    assert _close_iters(it, multi_thr), it
    b0 = np.ascontiguousarray(flat[:n])
    it, x, rel = a.cg_solve_single(b0, 10000, 1e-5)                 # CGSolveSingle on B[0:n]
    assert _close_iters(it, single_raw), it
    it, x, rel = a.cg_solve_single(b0, 10000, thr)
    assert _close_iters(it, single_thr), it
    a.close()


# ---- configs[1]: the 150^3 single-RHS solve against CGSolveSingle ---------------------------------
def test_cg_single_grid3d_150_against_oracle(gpu, orc):
    be = O.ref() or orc                                             # the compiled reference when it travelled
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    b = gpu.gen_rhs_rand(42, n)
    it, x, rel = a.cg_solve_single(b, 10000, 1e-5)
    it_ref, x_ref = be.cg_single(ro, ci, va, b, 10000, 1e-5)
    assert _close_iters(it, it_ref), (it, it_ref)
    assert rel < 1e-5
    assert np.abs(x - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    # the same system through the row-partitioned solver (a partition of one rank), host buffers
    from smle_b200 import dist as D
    A = D.RowPartitionedCsr.from_global(ro, ci, va, 0, 1, lambda obj: [obj])
    it2, x2, _ = A.cg_solve_single(b, 10000, 1e-5)
    assert _close_iters(it2, it_ref), (it2, it_ref)
    assert np.abs(x2 - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    A.close()
    a.close()


# ---- configs[3]: the wheel hub row (one row of 2^20 nonzeros spanning hundreds of tiles) ----------
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_wheel_2_20_hub_row(gpu, orc, dtype, tol):
    s = 1 << 20
    ro, ci, va = gpu.gen_wheel(s, 1.0, dtype)
    a = gpu.CsrMatrix(ro, ci, va)
    rng = np.random.default_rng(3)
    x = rng.random(s + 1).astype(dtype)
    y = a.spmv(x)
    y_ref = orc.spmv_gold(ro, ci, va, x)
    # the hub: a sum of 2^20 positive terms, so |A||x| = y and both metrics coincide.  The serial gold
    # itself carries up to n*eps of rounding there, so the hub is held against the exactly rounded sum
    exact = float(np.sum(x[1:].astype(np.longdouble)))
    assert abs(float(y[0]) - exact) <= tol * exact
    assert abs(float(y_ref[0]) - exact) <= 100 * tol * exact         # sanity of the oracle's own value
    assert rel_rownorm_err(y[1:], y_ref[1:]) <= tol                 # rim rows: one nonzero each, exact
    for k in (8, 32):
        X = rng.random((s + 1, k)).astype(dtype)
        Y = a.spmm(X)
        Y_ref = orc.merge_csrmm(8, ro, ci, va, X, k)
        assert rel_rownorm_err(Y[1:], Y_ref[1:]) <= tol, k
        hub = np.sum(X[1:].astype(np.longdouble), axis=0).astype(np.float64)
        assert np.abs(Y[0] - hub).max() <= tol * np.abs(hub).max(), k
    a.close()


# ---- configs[4]: 300^3 through the row-partitioned solver; checked by the TRUE residual -----------
def test_rowcg_grid3d_300_true_residual(gpu):
    import torch
    from smle_b200 import dist as D
    A = D.RowPartitionedCsr.grid3d(300, 0, 1, lambda obj: [obj])
    n = A.n_local
    assert (n, A.num_nonzeros_global) == (27000000, 188460000)
    b = torch.from_numpy(gpu.gen_rhs_rand(42, n)).cuda()
    it, x, rel = A.cg_solve_single(b, 10000, 1e-5)
    assert 500 < it < 1000 and rel < 1e-5
    r = b - A.spmv(x)                                               # recomputed, not the recurrence's r
    true_rel = float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b))
    assert true_rel < 1.01e-5, true_rel
    # Poisson row sums: A 1 = (6 - #neighbours), zero in the interior
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    y = A.spmv(ones)
    assert float(y.min()) == 0.0 and float(y.max()) == 3.0 and abs(float(y.sum()) - 6 * 300 * 300) < 1e-6
    A.close()


# ---- one handle, interleaved calls: graphs must follow the scratch they were captured with --------
def test_interleaved_cg_spmm_cg_on_one_handle(gpu, orc):
    """cg(k) -> spmm(k2 > k) -> cg(k) on ONE handle: the second product grows the carry / partial
    scratch the first solve's CUDA graph had baked in: the replay must not use it."""
    ro, ci, va = gpu.gen_grid3d(30, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    rng = np.random.default_rng(4)
    b = gpu.gen_rhs_rand(42, n)
    it_ref, x_ref = orc.cg_single(ro, ci, va, b, 10000, 1e-8)
    for k2 in (8, 32, 64):
        it, x, _ = a.cg_solve_single(b, 10000, 1e-8)
        assert _close_iters(it, it_ref) and np.abs(x - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
        X = rng.random((n, k2))
        Y = a.spmm(X)
        assert rel_rownorm_err(Y, orc.merge_csrmm(8, ro, ci, va, X, k2), (ro, ci, va), X) <= 1e-12
        y = a.spmv(b)
        assert rel_rownorm_err(y, orc.spmv_gold(ro, ci, va, b), (ro, ci, va), b) <= 1e-12
    it, x, _ = a.cg_solve_single(b, 10000, 1e-8)
    assert _close_iters(it, it_ref) and np.abs(x - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    # multi-RHS: k = 4 solve, k = 32 product (grows the per-tile carry slots), k = 4 solve again
    B = gpu.gen_rhs_rand(42, n * 4).reshape(n, 4)
    it_m, X_m, _ = orc.cg_multi(ro, ci, va, B, 4, 10000, 1e-8, O.MERGE, 8)
    for _ in range(2):
        it, Xs, hist, rel = a.cg_solve_multiple(B, 10000, 1e-8)
        assert _close_iters(it, it_m)
        np.testing.assert_allclose(Xs, X_m, rtol=1e-6, atol=1e-9)
        a.spmm(rng.random((n, 128)))
    a.close()


# ---- skewed matrices large enough for the skewed-matrix configuration (R-MAT scale 20: 17.8 M merge items) ----
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_rmat_20_general_tiles_skewed_configuration(gpu, orc, dtype, tol):
    """R-MAT scale 20 x16: most tiles hold a row segment longer than 32, so the fp64 handle picks two stages of
    1920 items (more L1 for the scattered gathers; fp32 keeps tiles of 2880); every tier of the product-staged general path (thread / warp / CTA per segment, tiles
    inside one row) is exercised against the gold loop, componentwise in |A||x|."""
    ro, ci, va = gpu.gen_rmat(20, 16, seed=7, dtype=dtype)
    n = len(ro) - 1
    x = np.random.default_rng(3).random(n).astype(dtype)
    a = gpu.CsrMatrix(ro, ci, va, n)
    _, items = a.tile_coords(1)
    assert items == (1920 if dtype == np.float64 else 2880)
    y = a.spmv(x)
    gold = orc.spmv_gold(ro, ci, va, x)
    assert rel_rownorm_err(y, gold, (ro, ci, va), x) <= tol
    # the path is deterministic by construction (fixed summation trees, wait-free but order-independent carries):
    # a race in the double / triple buffered shared-memory queues would show up as a run-to-run difference
    for _ in range(8):
        assert np.array_equal(a.spmv(x), y)
    a.close()
