"""Single-RHS CG on systems that fit in L2 (3-D Poisson 64^3 .. 100^3): iterations per second of full solves with device
buffers, for the A/B of the matrix stream's L2 priority (SMLE_SPMV_KEEP_MB=0: always evict-first, 96: default).
usage: python tools/small_cg_keep_ab.py [width ...]"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
import torch  # noqa: E402
import smle_b200 as S  # noqa: E402

S.init(0)
st = torch.cuda.Stream()
S.set_stream(st.cuda_stream)
for w in [int(v) for v in sys.argv[1:]] or [64, 80, 100]:
    ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
    n = len(ro) - 1
    a = S.CsrMatrix(ro, ci, va)
    with torch.cuda.stream(st):
        b = torch.from_numpy(S.gen_rhs_rand(42, n)).cuda()
        x = torch.empty_like(b)
        it, _, rel = a.cg_solve_single(b, 10000, 1e-8, out=x)      # warm-up (graphs, partitions)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            it, _, rel = a.cg_solve_single(b, 10000, 1e-8, out=x)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    foot = (len(ci) * 12 + (n + 1) * 4 + 5 * n * 8) / 2**20
    print(json.dumps({"grid3d": w, "rows": n, "system_MB": round(foot, 1), "keep_mb": os.environ.get("SMLE_SPMV_KEEP_MB", "96"),
                      "iterations": it, "us_per_iteration": round(dt / it * 1e6, 2), "iterations_per_s": round(it / dt, 1), "rel_res": rel}), flush=True)
    a.close()
