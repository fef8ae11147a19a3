"""GPU: SPAI-preconditioned multi-RHS CG (smle_pcg_spai_multi_f64) against the oracle's restatement
of SPAISolveMultiple (work_2025/main/sparse_approximate_inverse.hpp:31-230): iteration counts within
2 %, solutions, the per-iteration error history; M from the product's host construction."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _close_iters(got, want):
    return abs(got - want) <= max(1, round(0.02 * want))


@pytest.mark.parametrize("k", [1, 2, 4, 8, 32])
def test_pcg_spai_against_oracle(gpu, orc, k):
    ro, ci, va = gpu.gen_grid3d(20, True, 6.0, -1.0)
    n = len(ro) - 1
    mv = gpu.spai_build(ro, ci, va)
    assert np.abs(mv - orc.spai_build(ro, ci, va)).max() <= 1e-12 * np.abs(mv).max()
    a, m = gpu.CsrMatrix(ro, ci, va), gpu.CsrMatrix(ro, ci, mv)
    B = gpu.gen_rhs_rand(42, n * k).reshape(n, k)
    it, X, hist, rel = a.pcg_spai_solve_multiple(m, B, 10000, 1e-8)
    it_o, X_o, hist_o = orc.spai_solve_multi(ro, ci, va, mv, B, k, 10000, 1e-8, O.MERGE, 8)
    assert _close_iters(it, it_o), (it, it_o)
    np.testing.assert_allclose(X, X_o, rtol=1e-6, atol=1e-9)
    nh = min(len(hist), len(hist_o)) - 2
    np.testing.assert_allclose(hist[:nh], hist_o[:nh], rtol=1e-5)
    assert rel < 1e-8 and len(hist) == it
    # fewer iterations than the plain solver on the same handle, and the plain solver still works after
    it_plain, X_plain, _, _ = a.cg_solve_multiple(B, 10000, 1e-8)
    assert it < it_plain
    np.testing.assert_allclose(X_plain, X_o, rtol=1e-5, atol=1e-8)
    for c in range(k):
        r = B[:, c] - orc.spmv_gold(ro, ci, va, np.ascontiguousarray(X[:, c]))
        assert np.linalg.norm(r) / np.linalg.norm(B[:, c]) < 1e-7
    a.close(); m.close()


def test_pcg_spai_all_kernel_values_device_pointers_and_cap(gpu, orc):
    import torch
    ro, ci, va = gpu.gen_grid2d(60, True, 4.0, -1.0)
    n = len(ro) - 1
    mv = gpu.spai_build(ro, ci, va)
    a, m = gpu.CsrMatrix(ro, ci, va), gpu.CsrMatrix(ro, ci, mv)
    B = gpu.gen_rhs_rand(42, n * 4).reshape(n, 4)
    it_o, X_o, _ = orc.spai_solve_multi(ro, ci, va, mv, B, 4, 10000, 1e-7, O.MERGE, 8)
    Bd = torch.from_numpy(B).cuda()
    for kernel in (O.SIMPLE, O.MERGE, O.NONZERO_SPLIT):     # all three compute Y = A X here (see the header)
        it, X, hist, rel = a.pcg_spai_solve_multiple(m, Bd, 10000, 1e-7, kernel)
        assert _close_iters(it, it_o), (kernel, it, it_o)
        np.testing.assert_allclose(X.cpu().numpy(), X_o, rtol=1e-6, atol=1e-9)
    it, X, hist, rel = a.pcg_spai_solve_multiple(m, B, 5, 1e-30)     # max_iters cap
    assert it == 5 and len(hist) == 5
    it, X, hist, rel = a.pcg_spai_solve_multiple(m, B, 0, 1e-7)
    assert it == 0 and np.all(X == 0)
    a.close(); m.close()


def test_pcg_spai_at_baseline_size(gpu):
    """150^3 x 8: converges to the tolerance in fewer iterations than plain CG; checked by the true residual"""
    import torch
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    mv = gpu.spai_build(ro, ci, va)
    a, m = gpu.CsrMatrix(ro, ci, va), gpu.CsrMatrix(ro, ci, mv)
    B = torch.from_numpy(gpu.gen_rhs_rand(42, n * 8).reshape(n, 8)).cuda()
    it, X, hist, rel = a.pcg_spai_solve_multiple(m, B, 10000, 1e-5)
    it_plain, Xp, _, _ = a.cg_solve_multiple(B, 10000, 1e-5)
    assert it < 0.8 * it_plain and rel < 1e-5
    R = B - a.spmm(X)
    true_rel = (torch.linalg.vector_norm(R, dim=0) / torch.linalg.vector_norm(B, dim=0)).max().item()
    assert true_rel < 1.05e-5, true_rel
    a.close(); m.close()
