"""Small end-to-end pass over every kernel family, for `compute-sanitizer --tool memcheck` where that
tool is available (it is closed on the round-1 pool: the gpu tests' comparisons with the oracle on
small and ragged cases are the bounds check there)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
import smle_b200 as S  # noqa: E402

S.init(0)
rng = np.random.default_rng(0)
for name, (ro, ci, va) in (("grid3d_12", S.gen_grid3d(12, True, 6.0, -1.0)), ("rmat_9", S.gen_rmat(9, 16, seed=1)),
                           ("wheel_5000", S.gen_wheel(5000))):
    n = max(len(ro) - 1, int(ci.max()) + 1)
    for dt in (np.float64, np.float32):
        a = S.CsrMatrix(ro, ci, va.astype(dt), n)
        a.spmv(rng.random(n).astype(dt))
        for k in (2, 8, 32, 33):
            a.spmm(rng.random((n, k)).astype(dt))
        a.close()
ro, ci, va = S.gen_grid3d(12, True, 6.0, -1.0)
n = len(ro) - 1
a = S.CsrMatrix(ro, ci, va)
B = S.gen_rhs_rand(42, n * 4).reshape(n, 4)
print("cg multi", a.cg_solve_multiple(B, 500, 1e-6)[0])
print("cg single", a.cg_solve_single(np.ascontiguousarray(B[:, 0]), 500, 1e-6)[0])
print("cg batch", a.cg_solve_single_batch(np.ascontiguousarray(B.T), 500, 1e-6)[0])
a.close()
print("sanitize_small done")
