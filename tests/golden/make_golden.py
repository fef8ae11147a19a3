"""Generate tests/golden/reference_vectors.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
It drives oracle/_ref (the reference sources compiled where they lie by oracle/Makefile) and
records, for small seeded inputs, what the reference's own functions return:
merge-path thread coordinates (MergePathSearch), OmpMergeCsrmv / OmpMergeCsrmm outputs,
SpmvGold outputs, and CGSolveSingle / CGSolveMultiple iteration counts + solution digests.
The fixture travels to the GPU box; /root/reference does not.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402


def matrices(ref):
    yield "wheel10", ref.gen_wheel(10)
    yield "wheel1000", ref.gen_wheel(1000)
    yield "dense4x3", ref.gen_dense(4, 3)
    yield "dense17x9", ref.gen_dense(17, 9)
    yield "grid2d_12_noloop", ref.gen_grid2d(12, False)
    yield "grid2d_12_poisson", ref.gen_grid2d(12, True, 4.0, -1.0)
    yield "grid3d_6_poisson", ref.gen_grid3d(6, True, 6.0, -1.0)
    yield "grid3d_24_loop", ref.gen_grid3d(24, True)


def main():
    ref = O.ref()
    assert ref is not None, "oracle/_ref is not built (needs /root/reference)"
    port = O.port()
    out = {"generated_by": "tests/golden/make_golden.py", "source": "oracle/_ref (unmodified reference)",
           "partition": [], "spmv": [], "spmm": [], "cg": []}
    rng = np.random.default_rng(20261018)
    for name, (ro, ci, va) in matrices(ref):
        m, nnz = len(ro) - 1, len(ci)
        ncols = int(ci.max()) + 1 if nnz else m
        for T in (1, 2, 3, 4, 8, 17, 64):
            out["partition"].append({"matrix": name, "threads": T,
                                     "coords": ref.merge_partition(ro, T).tolist()})
        if m <= 1100:
            x = np.round(rng.random(max(ncols, m)) * 8 - 4, 3)
            for T in (1, 4, 8):
                y = ref.merge_csrmv(T, ro, ci, va, x)
                out["spmv"].append({"matrix": name, "threads": T, "x": x.tolist(), "y": y.tolist(),
                                    "y_gold": ref.spmv_gold(ro, ci, va, x).tolist()})
            for k in (1, 2, 5, 8):
                X = np.round(rng.random((max(ncols, m), k)) * 8 - 4, 3)
                Y = ref.merge_csrmm(4, ro, ci, va, X, k)
                out["spmm"].append({"matrix": name, "threads": 4, "k": k, "X": X.tolist(), "Y": Y.tolist()})
    # CG on 3-D Poisson, RHS = srand(42) stream
    for w, k in ((8, 1), (8, 4), (16, 3), (24, 4)):
        ro, ci, va = ref.gen_grid3d(w, True, 6.0, -1.0)
        n = len(ro) - 1
        B = port.rhs_rand(42, n * k).reshape(n, k)
        thr = port.driver_threshold(B.ravel(), n, 1e-5)
        for tol_name, tol in (("raw_1e-5", 1e-5), ("driver_threshold", thr)):
            it, X, hist = ref.cg_multi(ro, ci, va, B, k, 10000, tol, O.MERGE, 8)
            it1, x1 = ref.cg_single(ro, ci, va, np.ascontiguousarray(B.ravel()[:n]), 10000, tol)
            out["cg"].append({"grid3d": w, "k": k, "tol_name": tol_name, "tol": tol,
                              "multi_iters": it, "multi_hist_tail": hist[-3:].tolist(),
                              "multi_x_sum": float(X.sum()), "multi_x_absmax": float(np.abs(X).max()),
                              "single_iters_on_flat_b0": it1, "single_x_sum": float(x1.sum())})
    path = Path(__file__).with_name("reference_vectors.json")
    path.write_text(json.dumps(out))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
