"""CPU: the C restatement against the UNMODIFIED reference compiled into oracle/_ref
(skipped where the reference sources were never available, e.g. a box without the prebuilt .so)."""
import numpy as np
import pytest

from conftest import rel_rownorm_err
from oracle import oracle as O


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_generators_identical(orc, ref, dtype):
    for w in (1, 2, 3, 5, 11):
        for loop in (False, True):
            assert _same(orc.gen_grid2d(w, loop, 4.0, -1.0, dtype), ref.gen_grid2d(w, loop, 4.0, -1.0, dtype)) or w == 1
            assert _same(orc.gen_grid3d(w, loop, 6.0, -1.0, dtype), ref.gen_grid3d(w, loop, 6.0, -1.0, dtype))
    for s in (1, 2, 7, 300):
        assert _same(orc.gen_wheel(s, 1.0, dtype), ref.gen_wheel(s, 1.0, dtype))
    assert _same(orc.gen_dense(6, 5, 2.0, dtype), ref.gen_dense(6, 5, 2.0, dtype))


def test_merge_path_search_every_diagonal(orc, ref):
    rng = np.random.default_rng(7)
    for trial in range(20):
        m = int(rng.integers(1, 60))
        deg = rng.integers(0, 9, size=m)
        deg[rng.random(m) < 0.3] = 0                      # empty rows
        ro = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        nnz = int(ro[-1])
        for d in range(m + nnz + 1):
            assert orc.merge_path_search(d, ro[1:], nnz) == ref.merge_path_search(d, ro[1:], nnz)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-14), (np.float32, 1e-6)])
def test_kernels_match_reference(orc, ref, dtype, tol):
    rng = np.random.default_rng(3)
    for (ro, ci, va) in (ref.gen_grid3d(7, True, 6.0, -1.0, dtype), ref.gen_wheel(500, 1.0, dtype),
                         ref.gen_dense(9, 13, 0.5, dtype)):
        m = len(ro) - 1
        n = int(ci.max()) + 1
        x = rng.random(max(m, n)).astype(dtype)
        for T in (1, 5, 8):
            assert rel_rownorm_err(orc.merge_csrmv(T, ro, ci, va, x), ref.merge_csrmv(T, ro, ci, va, x)) <= tol
        assert rel_rownorm_err(orc.spmv_gold(ro, ci, va, x), ref.spmv_gold(ro, ci, va, x)) <= tol
        for k in (1, 3, 8):
            X = rng.random((max(m, n), k)).astype(dtype)
            for fn in ("merge_csrmm", "nonzero_split_csrmm", "row_split_csrmm"):
                assert rel_rownorm_err(getattr(orc, fn)(4, ro, ci, va, X, k),
                                       getattr(ref, fn)(4, ro, ci, va, X, k)) <= tol, fn


@pytest.mark.parametrize("kernel", [O.SIMPLE, O.MERGE, O.NONZERO_SPLIT])
def test_cg_matches_reference(orc, ref, kernel):
    ro, ci, va = ref.gen_grid3d(12, True, 6.0, -1.0)
    n = len(ro) - 1
    B = orc.rhs_rand(42, n * 3).reshape(n, 3)
    it_o, X_o, h_o = orc.cg_multi(ro, ci, va, B, 3, 5000, 1e-8, kernel, 8)
    it_r, X_r, h_r = ref.cg_multi(ro, ci, va, B, 3, 5000, 1e-8, kernel, 8)
    assert it_o == it_r and len(h_o) == len(h_r)
    np.testing.assert_allclose(X_o, X_r, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(h_o, h_r, rtol=1e-6)
    b0 = np.ascontiguousarray(B[:, 1])
    it_so, x_so = orc.cg_single(ro, ci, va, b0, 5000, 1e-8)
    it_sr, x_sr = ref.cg_single(ro, ci, va, b0, 5000, 1e-8)
    assert it_so == it_sr
    np.testing.assert_allclose(x_so, x_sr, rtol=1e-9, atol=1e-12)


def test_cg_max_iters_and_zero_rhs(orc, ref):
    ro, ci, va = ref.gen_grid3d(6, True, 6.0, -1.0)
    n = len(ro) - 1
    B = orc.rhs_rand(1, n * 2).reshape(n, 2)
    assert orc.cg_multi(ro, ci, va, B, 2, 5, 1e-12, O.MERGE, 4)[0] == ref.cg_multi(ro, ci, va, B, 2, 5, 1e-12, O.MERGE, 4)[0] == 5
    assert orc.cg_single(ro, ci, va, B[:, 0].copy(), 0, 1e-5)[0] == ref.cg_single(ro, ci, va, B[:, 0].copy(), 0, 1e-5)[0] == 0
