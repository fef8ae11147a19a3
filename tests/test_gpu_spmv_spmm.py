"""GPU: SpMV / SpMM through the C ABI against the oracle and the golden fixture.
Tolerance (north_star): 1e-12 relative per row-norm in fp64, 1e-5 in fp32."""
import numpy as np
import pytest

from conftest import named_matrix, rel_rownorm_err

pytestmark = pytest.mark.gpu
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _matrices(S, dtype):
    yield "grid3d_poisson_20", S.gen_grid3d(20, True, 6.0, -1.0, dtype)
    yield "grid2d_noloop_70", S.gen_grid2d(70, False, 1.0, 1.0, dtype)
    yield "wheel_30000", S.gen_wheel(30000, 1.0, dtype)
    yield "dense_40x33", S.gen_dense(40, 33, 0.5, dtype)
    yield "rmat_12", S.gen_rmat(12, 16, seed=3, dtype=dtype)


def _with_empty_rows(rng, m, n, dtype):
    deg = rng.integers(0, 6, size=m)
    deg[rng.random(m) < 0.4] = 0
    deg[m // 2] = 5000
    ro = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ci = np.concatenate([np.sort(rng.integers(0, n, size=d)) for d in deg]).astype(np.int32)
    va = (rng.random(len(ci)) - 0.5).astype(dtype)
    return ro, ci, va


def test_golden_spmv(gpu, golden):
    for rec in golden["spmv"]:
        ro, ci, va = named_matrix(gpu, rec["matrix"])
        x = np.array(rec["x"])
        ncols = len(x)
        a = gpu.CsrMatrix(ro, ci, va, ncols)
        y = a.spmv(x)
        assert rel_rownorm_err(y, np.array(rec["y"]), (ro, ci, va), x) <= 1e-12, rec["matrix"]
        assert rel_rownorm_err(y, np.array(rec["y_gold"]), (ro, ci, va), x) <= 1e-12, rec["matrix"]
        a.close()


def test_golden_spmm(gpu, golden):
    for rec in golden["spmm"]:
        ro, ci, va = named_matrix(gpu, rec["matrix"])
        X = np.array(rec["X"]).reshape(-1, rec["k"])
        a = gpu.CsrMatrix(ro, ci, va, X.shape[0])
        Y = a.spmm(X)
        assert rel_rownorm_err(Y, np.array(rec["Y"]).reshape(-1, rec["k"]), (ro, ci, va), X) <= 1e-12, (rec["matrix"], rec["k"])
        a.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spmv_against_oracle(gpu, orc, dtype):
    rng = np.random.default_rng(11)
    for name, (ro, ci, va) in _matrices(gpu, dtype):
        m, n = len(ro) - 1, int(ci.max()) + 1
        x = rng.random(max(m, n)).astype(dtype)
        a = gpu.CsrMatrix(ro, ci, va, len(x))
        y = a.spmv(x)
        assert rel_rownorm_err(y, orc.spmv_gold(ro, ci, va, x), (ro, ci, va), x) <= TOL[dtype], name
        assert rel_rownorm_err(y, orc.merge_csrmv(8, ro, ci, va, x), (ro, ci, va), x) <= TOL[dtype], name
        a.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 12, 16, 32, 33, 64, 128])
def test_spmm_against_oracle(gpu, orc, dtype, k):
    rng = np.random.default_rng(100 + k)
    for name, (ro, ci, va) in _matrices(gpu, dtype):
        if k > 32 and name.startswith("rmat"):
            continue
        m, n = len(ro) - 1, int(ci.max()) + 1
        X = (rng.random((max(m, n), k)) - 0.5).astype(dtype)
        a = gpu.CsrMatrix(ro, ci, va, X.shape[0])
        Y = a.spmm(X)
        assert rel_rownorm_err(Y, orc.merge_csrmm(8, ro, ci, va, X, k), (ro, ci, va), X) <= TOL[dtype], (name, k)
        a.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_empty_rows_long_row_and_rectangular(gpu, orc, dtype):
    rng = np.random.default_rng(5)
    ro, ci, va = _with_empty_rows(rng, 3000, 777, dtype)
    a = gpu.CsrMatrix(ro, ci, va, 777)
    x = rng.random(777).astype(dtype)
    assert rel_rownorm_err(a.spmv(x), orc.spmv_gold(ro, ci, va, x, n=777), (ro, ci, va), x) <= TOL[dtype]
    X = rng.random((777, 6)).astype(dtype)
    assert rel_rownorm_err(a.spmm(X), orc.merge_csrmm(4, ro, ci, va, X, 6, n=777), (ro, ci, va), X) <= TOL[dtype]
    a.close()


def test_all_rows_empty_and_single_row(gpu, orc):
    ro = np.zeros(11, dtype=np.int32)
    a = gpu.CsrMatrix(ro, np.zeros(0, np.int32), np.zeros(0), 10)
    assert np.array_equal(a.spmv(np.ones(10)), np.zeros(10))
    assert np.array_equal(a.spmm(np.ones((10, 4))), np.zeros((10, 4)))
    a.close()
    ro, ci, va = gpu.gen_dense(1, 5000, 2.0)
    a = gpu.CsrMatrix(ro, ci, va, 5000)
    x = np.arange(5000, dtype=np.float64)
    assert abs(a.spmv(x)[0] - 2.0 * x.sum()) <= 1e-12 * 2.0 * x.sum()
    a.close()


def test_device_pointers_and_determinism(gpu, orc):
    import torch
    ro, ci, va = gpu.gen_rmat(13, 16, seed=9)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    X = torch.rand(n, 8, dtype=torch.float64, device="cuda")
    Y1 = a.spmm(X)
    Y2 = a.spmm(X)
    gpu.sync()
    assert torch.equal(Y1, Y2), "SpMM must be deterministic run to run"
    assert rel_rownorm_err(Y1.cpu().numpy(), orc.merge_csrmm(8, ro, ci, va, X.cpu().numpy(), 8), (ro, ci, va), X.cpu().numpy()) <= 1e-12
    a.close()


def test_linearity_at_scale(gpu):
    """size-independent property at a BASELINE-sized matrix: A(ax + by) == a Ax + b Ay."""
    import torch
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    lhs = a.spmv((2.5 * x - 0.75 * y).contiguous())
    rhs = 2.5 * a.spmv(x) - 0.75 * a.spmv(y)
    gpu.sync()
    assert (lhs - rhs).abs().max().item() <= 1e-12 * rhs.abs().max().item()
    # Poisson row sums: A * ones = 6 - (number of neighbours)
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    s = a.spmv(ones)
    gpu.sync()
    deg = torch.from_numpy(np.diff(ro).astype(np.float64)).cuda() - 1.0
    assert torch.equal(s, 6.0 - deg)
    a.close()


def test_spmm_at_scale_matches_column_spmv(gpu):
    """BASELINE-sized SpMM (200^3 would need 4 GB of blocks; 150^3 x 32 exercises the same
    round-robin tile deal over all 148 CTAs): every column of A X equals the single-vector product
    of that column, and A * ones has the exact Poisson row sums."""
    import torch
    ro, ci, va = gpu.gen_grid3d(150, True, 6.0, -1.0)
    n = len(ro) - 1
    a = gpu.CsrMatrix(ro, ci, va)
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.rand(n, 32, dtype=torch.float64, device="cuda", generator=g) - 0.5
    Y = a.spmm(X)
    Y2 = a.spmm(X)
    gpu.sync()
    assert torch.equal(Y, Y2), "SpMM must be deterministic run to run"
    for c in (0, 13, 31):
        y = a.spmv(X[:, c].contiguous())
        gpu.sync()
        assert (Y[:, c] - y).abs().max().item() <= 1e-12 * 12.0 * 0.5
    ones = torch.ones(n, 32, dtype=torch.float64, device="cuda")
    S = a.spmm(ones)
    gpu.sync()
    deg = torch.from_numpy(np.diff(ro).astype(np.float64)).cuda() - 1.0
    assert torch.equal(S, (6.0 - deg)[:, None].expand(n, 32))
    a.close()


ROWS_KERNEL_CHILD = r'''
import sys
import numpy as np
sys.path[:0] = [sys.argv[1], sys.argv[1] + "/tests", sys.argv[1] + "/sparse-matrix-linear-equations_b200/python"]
import smle_b200 as S
from oracle import oracle as O
from conftest import rel_rownorm_err
orc = O.port()
S.init(0)
rng = np.random.default_rng(3)
def empty_rows(m, n):
    deg = rng.integers(0, 6, size=m); deg[rng.random(m) < 0.4] = 0; deg[m // 2] = 9000
    ro = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ci = np.concatenate([np.sort(rng.integers(0, n, size=d)) for d in deg]).astype(np.int32)
    return ro, ci, (rng.random(len(ci)) - 0.5)
cases = [("wheel", S.gen_wheel(30000)), ("dense", S.gen_dense(40, 3300, 0.5)), ("rmat", S.gen_rmat(12, 16, seed=3)),
         ("empty_rows", empty_rows(3000, 777)), ("grid", S.gen_grid3d(20, True, 6.0, -1.0))]
for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
    for name, (ro, ci, va) in cases:
        va = va.astype(dtype)
        n = int(ci.max()) + 1
        for k in (2, 5, 32, 33):
            X = (rng.random((max(n, len(ro) - 1), k)) - 0.5).astype(dtype)
            a = S.CsrMatrix(ro, ci, va, X.shape[0])
            coords, items = a.tile_coords(k)
            assert items in (1920, 1440), (name, k, items)   # the row-per-worker kernel's tilings
            Y = a.spmm(X)
            err = rel_rownorm_err(Y, orc.merge_csrmm(8, ro, ci, va, X, k, n=X.shape[0]), (ro, ci, va), X)
            assert err <= tol, (name, k, dtype.__name__, err)
            assert np.array_equal(Y, a.spmm(X)), (name, k, "not deterministic")
            a.close()
print("ROWS_KERNEL_OK")
'''


def test_rows_kernel_on_skewed_matrices(gpu):
    """The row-per-worker SpMM kernel is only dispatched for matrices without long rows, but its
    carry exchange is general: forced (SMLE_SPMM_ROWS_MAXLEN) onto wheel / dense / R-MAT / empty-row
    matrices it must still match the oracle -- rows spanning many tiles exercise the pass-on path."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, SMLE_SPMM_ROWS_MAXLEN="2000000000")
    r = subprocess.run([sys.executable, "-c", ROWS_KERNEL_CHILD, str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert "ROWS_KERNEL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spmm_against_all_three_spmmkernel_oracles(gpu, orc, dtype):
    """SpmmKernel {SIMPLE, MERGE, NONZERO_SPLIT} (work_2025/types.hpp:11-16) are three CPU threading
    strategies for the same product; on the GPU all three map to the merge-path kernel
    (include/smle_b200.h).  The one GPU result must agree with each of the three reference kernels --
    OmpCsrSpmmT (row_splitting.hpp:18-54), OmpMergeCsrmm (merge_based.hpp:49-153),
    OmpNonzeroSplitCsrmm (nonzero_splitting.hpp:52-150, on a zeroed output: it adds into the last row)."""
    rng = np.random.default_rng(5)
    for name, (ro, ci, va) in _matrices(gpu, dtype):
        m, n = len(ro) - 1, max(len(ro) - 1, int(ci.max()) + 1)
        a = gpu.CsrMatrix(ro, ci, va, n)
        for k in (3, 16):
            X = rng.random((n, k)).astype(dtype)
            Y = a.spmm(X)
            for T in (1, 8):
                for oracle_kernel in (orc.row_split_csrmm, orc.merge_csrmm, orc.nonzero_split_csrmm):
                    Y_ref = oracle_kernel(T, ro, ci, va, X, k, n)
                    assert rel_rownorm_err(Y, Y_ref, (ro, ci, va), X) <= TOL[dtype], (name, k, T, oracle_kernel.__name__)
        a.close()
