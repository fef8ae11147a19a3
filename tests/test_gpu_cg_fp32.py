"""GPU: the fp32 solvers against the reference's templates instantiated for <float,int>
(oracle port; oracle/_ref where it travelled).  In fp32 the recurrence is sensitive to the order of the
dot-product sums (the reference accumulates them in float with OpenMP reductions, the GPU in double), so
the bar is: the same tolerance reached, iteration counts within 5 %, solutions within 1e-3 relative."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [1, 3, 4, 16, 32])
def test_fp32_multi_rhs_cg(gpu, orc, k):
    ro, ci, va = gpu.gen_grid3d(24, True, 6.0, -1.0, np.float32)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    B = gpu.gen_rhs_rand(42, n * k).reshape(n, k).astype(np.float32)
    it, X, hist, rel = a.cg_solve_multiple(B, 10000, 1e-4)
    it_o, X_o, hist_o = orc.cg_multi(ro, ci, va, B, k, 10000, np.float32(1e-4), O.MERGE, 8)
    assert X.dtype == np.float32 and rel < 1e-4 and len(hist) == it
    assert abs(it - it_o) <= max(1, round(0.05 * it_o)), (it, it_o)
    assert np.abs(X - X_o).max() <= 1e-3 * np.abs(X_o).max()
    # the true residual in double
    A64 = va.astype(np.float64)
    for c in range(min(k, 4)):
        r = B[:, c].astype(np.float64) - orc.spmv_gold(ro, ci, A64, X[:, c].astype(np.float64))
        assert np.linalg.norm(r) / np.linalg.norm(B[:, c]) < 2e-4
    a.close()


def test_fp32_single_rhs_cg_host_and_device(gpu, orc):
    import torch
    ro, ci, va = gpu.gen_grid2d(64, True, 4.0, -1.0, np.float32)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    b = gpu.gen_rhs_rand(42, n).astype(np.float32)
    it, x, rel = a.cg_solve_single(b, 10000, 1e-4)
    it_o, x_o = orc.cg_single(ro, ci, va, b, 10000, np.float32(1e-4))
    assert abs(it - it_o) <= max(1, round(0.05 * it_o)), (it, it_o)
    assert np.abs(x - x_o).max() <= 1e-3 * np.abs(x_o).max()
    it_d, x_d, _ = a.cg_solve_single(torch.from_numpy(b).cuda(), 10000, 1e-4)
    assert it_d == it and np.array_equal(x_d.cpu().numpy(), x)
    it_c, _, _ = a.cg_solve_single(b, 7, 1e-30)          # max_iters cap
    assert it_c == 7
    a.close()
    a64 = gpu.CsrMatrix(ro, ci, va.astype(np.float64))    # the fp64 entry refuses fp32 blocks and vice versa
    with pytest.raises(gpu.SmleError):
        a64.cg_solve_single(b, 10, 1e-4)
    a64.close()
