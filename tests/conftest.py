"""pytest configuration: `gpu` marker, import paths, shared fixtures."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The plain-C oracle (oracle/smle_oracle.c)."""
    from oracle import oracle as O
    return O.port()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref); tests that need it skip when it is absent."""
    from oracle import oracle as O
    r = O.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return r


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "reference_vectors.json").read_text())


@pytest.fixture(scope="session")
def S():
    """The product binding; initialised on cuda:0 for gpu tests."""
    import smle_b200
    return smle_b200


@pytest.fixture(scope="session")
def gpu(S):
    if S.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible")
    S.init(0)
    return S


def named_matrix(gen, name):
    """The matrices of tests/golden/make_golden.py, from any generator back-end."""
    table = {
        "wheel10": lambda: gen.gen_wheel(10),
        "wheel1000": lambda: gen.gen_wheel(1000),
        "dense4x3": lambda: gen.gen_dense(4, 3),
        "dense17x9": lambda: gen.gen_dense(17, 9),
        "grid2d_12_noloop": lambda: gen.gen_grid2d(12, False),
        "grid2d_12_poisson": lambda: gen.gen_grid2d(12, True, 4.0, -1.0),
        "grid3d_6_poisson": lambda: gen.gen_grid3d(6, True, 6.0, -1.0),
        "grid3d_24_loop": lambda: gen.gen_grid3d(24, True),
    }
    return table[name]()


def rel_rownorm_err(Y, Y_ref, csr=None, X=None):
    """north_star's "relative per row-norm" error:  max_i |y_i - yref_i| / scale_i.

    With csr=(row_offsets, column_indices, values) and X given, scale = (|A| |X|)_i -- the
    magnitude the row's dot product is built from (the standard componentwise bound for a
    rounding-order difference); otherwise scale_i = ||yref row i||_2 (only meaningful when no
    cancellation occurs, e.g. the golden vectors)."""
    R = np.asarray(Y_ref, dtype=np.float64)
    R = R.reshape(len(R), -1)
    Y = np.asarray(Y, dtype=np.float64).reshape(R.shape)
    if not len(R):
        return 0.0
    if csr is not None:
        import scipy.sparse as sp
        ro, ci, va = csr
        Xa = np.abs(np.asarray(X, dtype=np.float64)).reshape(-1, R.shape[1])
        A = sp.csr_matrix((np.abs(va.astype(np.float64)), ci, ro), shape=(len(ro) - 1, Xa.shape[0]))
        scale = A @ Xa
        scale[scale == 0] = 1.0
        return float((np.abs(Y - R) / scale).max())
    norm = np.linalg.norm(R, axis=1)
    norm[norm == 0] = 1.0
    return float((np.abs(Y - R).max(axis=1) / norm).max())
