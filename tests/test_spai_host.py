"""CPU: the SPAI preconditioner.

Three constructions of M must agree: the compiled reference (SparseApproximateInversion from
/root/reference with the shim's LAPACKE_dgels), the oracle's C restatement, and the product's host
code smle_spai_build_f64.  The oracle's SPAISolveMultiple restatement is pinned to the reference's
(iteration counts, solutions, error history) -- including the reference's NONZERO_SPLIT quirk."""
import numpy as np
import pytest

from oracle import oracle as O


def _spd_random(m, seed):
    import scipy.sparse as sp
    W = sp.random(m, m, density=5.0 / m, random_state=np.random.RandomState(seed), format="csr")
    W = W + W.T
    A = (sp.diags(np.asarray(W.sum(axis=1)).ravel() + 1.0) - W).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def _systems(orc):
    yield "poisson3d_8", orc.gen_grid3d(8, True, 6.0, -1.0)
    yield "poisson2d_17", orc.gen_grid2d(17, True, 4.0, -1.0)
    yield "random_spd_300", _spd_random(300, 3)


def test_spai_build_three_ways(S, orc, ref):
    for name, (ro, ci, va) in _systems(orc):
        m_ref = ref.spai_build(ro, ci, va)
        m_orc = orc.spai_build(ro, ci, va)
        m_lib = S.spai_build(ro, ci, va)
        scale = np.abs(m_ref).max()
        assert np.abs(m_orc - m_ref).max() <= 1e-13 * scale, name
        assert np.abs(m_lib - m_ref).max() <= 1e-12 * scale, name
        # symmetric by construction (sparse_approximate_inversion.hpp:268-318)
        import scipy.sparse as sp
        M = sp.csr_matrix((m_lib, ci, ro), shape=(len(ro) - 1,) * 2)
        assert abs(M - M.T).max() == 0.0, name


def test_spai_columns_minimise_the_residual(S, orc):
    """definition check, independent of any QR: column k of the un-symmetrised M solves the normal
    equations of min ||A(:,J) m - e_k||; after symmetrisation ||A M - I||_F must still be well below
    ||A diag(A)^-1 - I||_F (Jacobi), otherwise the preconditioner would be pointless"""
    import scipy.sparse as sp
    ro, ci, va = orc.gen_grid3d(7, True, 6.0, -1.0)
    n = len(ro) - 1
    A = sp.csr_matrix((va, ci, ro), shape=(n, n))
    M = sp.csr_matrix((S.spai_build(ro, ci, va), ci, ro), shape=(n, n))
    err_spai = sp.linalg.norm(A @ M - sp.identity(n))
    err_jacobi = sp.linalg.norm(A @ sp.diags(1.0 / A.diagonal()) - sp.identity(n))
    assert err_spai < 0.75 * err_jacobi


@pytest.mark.parametrize("kernel", [O.SIMPLE, O.MERGE])
def test_oracle_spai_solver_matches_reference(orc, ref, kernel):
    for name, (ro, ci, va) in _systems(orc):
        n = len(ro) - 1
        mv = ref.spai_build(ro, ci, va)
        B = orc.rhs_rand(42, n * 4).reshape(n, 4)
        it_o, X_o, h_o = orc.spai_solve_multi(ro, ci, va, mv, B, 4, 10000, 1e-8, kernel, 8)
        it_r, X_r, h_r = ref.spai_solve_multi(ro, ci, va, mv, B, 4, 10000, 1e-8, kernel, 8)
        assert it_o == it_r and len(h_o) == len(h_r) == it_r, name
        np.testing.assert_allclose(X_o, X_r, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(h_o, h_r, rtol=1e-6)
        it_plain = orc.cg_multi(ro, ci, va, B, 4, 10000, 1e-8, kernel, 8)[0]
        assert it_o < it_plain, (name, it_o, it_plain)     # the preconditioner earns its two products


def test_reference_nonzero_split_quirk_is_reproduced(orc, ref):
    """OmpNonzeroSplitCsrmm adds into the last row of Y (nonzero_splitting.hpp:137-149) and
    SPAISolveMultiple never clears Z / AP: with NONZERO_SPLIT the reference does not converge.  The
    oracle reproduces that; the GPU library documents it and computes Y = A X for all three values."""
    ro, ci, va = orc.gen_grid3d(6, True, 6.0, -1.0)
    n = len(ro) - 1
    mv = ref.spai_build(ro, ci, va)
    B = orc.rhs_rand(42, n * 2).reshape(n, 2)
    it_r = ref.spai_solve_multi(ro, ci, va, mv, B, 2, 300, 1e-8, O.NONZERO_SPLIT, 8)[0]
    it_o = orc.spai_solve_multi(ro, ci, va, mv, B, 2, 300, 1e-8, O.NONZERO_SPLIT, 8)[0]
    assert it_r == it_o == 300
