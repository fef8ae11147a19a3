"""GPU: the band-window SpMM variant (SMLE_SPMM_BAND=1, k = 32 fp64: dense rows of the +-band around the
tile served from a shared-memory ring) against the oracle -- product and fused p.Ap (through the
multi-RHS CG).  The switch is read once per process, so the check runs in a child process."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

CHILD = r'''
import json, sys
import numpy as np
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/sparse-matrix-linear-equations_b200/python"); sys.path.insert(0, "{root}/tests")
import smle_b200 as S
from oracle import oracle as O
from conftest import rel_rownorm_err
S.init(0)
orc = O.port()
out = {{}}
rng = np.random.default_rng(9)
for name, (ro, ci, va) in (("grid3d_40", S.gen_grid3d(40, True, 6.0, -1.0)), ("grid3d_100", S.gen_grid3d(100, True, 6.0, -1.0)),
                           ("grid2d_300", S.gen_grid2d(300, True, 4.0, -1.0))):
    n = len(ro) - 1
    a = S.CsrMatrix(ro, ci, va)
    X = rng.random((n, 32))
    Y = a.spmm(X)
    out[name + "_spmm_err"] = rel_rownorm_err(Y, orc.merge_csrmm(8, ro, ci, va, X, 32), (ro, ci, va), X)
    Y2 = a.spmm(X)
    out[name + "_deterministic"] = bool(np.array_equal(Y, Y2))
    if n <= 100000:
        B = S.gen_rhs_rand(42, n * 32).reshape(n, 32)
        it, Xs, hist, rel = a.cg_solve_multiple(B, 10000, 1e-8)
        it_o, X_o, _ = orc.cg_multi(ro, ci, va, B, 32, 10000, 1e-8, O.MERGE, 8)
        out[name + "_cg"] = [it, it_o, float(np.abs(Xs - X_o).max() / np.abs(X_o).max())]
    a.close()
print("RESULT " + json.dumps(out))
'''


def test_band_window_spmm_against_oracle(gpu):
    env = dict(os.environ, SMLE_SPMM_BAND="1", SMLE_SPMM_BAND_CHUNK="4")
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    for key, val in res.items():
        if key.endswith("_spmm_err"):
            assert val <= 1e-12, (key, val)
        elif key.endswith("_deterministic"):
            assert val, key
        else:
            it, it_o, err = val
            assert abs(it - it_o) <= max(1, round(0.02 * it_o)) and err <= 1e-6, (key, val)
