"""Small driver for ncu: runs the hot kernels a few times on a BASELINE-sized input.
usage: python tools/prof_kernels.py {spmv|spmm<k>|cg|cgmulti|rmat<k>|wheel<k>|grid2d} [grid_width or scale]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
import torch  # noqa: E402
import smle_b200 as S  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "spmv"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 150
S.init(0)
st = torch.cuda.Stream()
S.set_stream(st.cuda_stream)
if what.startswith("rmat"):      # rmat<k>: R-MAT scale w, 16 edges per vertex
    ro, ci, va = S.gen_rmat(w, 16, seed=42)
    what = ("spmm" + what[4:]) if what[4:] not in ("", "1") else "spmv"
elif what.startswith("wheel"):
    ro, ci, va = S.gen_wheel(1 << w)
    what = ("spmm" + what[5:]) if what[5:] not in ("", "1") else "spmv"
elif what.startswith("grid2d"):
    ro, ci, va = S.gen_grid2d(w, True)
    what = "spmv"
else:
    ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
n = len(ro) - 1
a = S.CsrMatrix(ro, ci, va)
with torch.cuda.stream(st):
    if what == "spmv":
        x = torch.rand(n, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        for _ in range(5):
            a.spmv(x, out=y)
        import os
        if os.environ.get("PROF_TIME"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                a.spmv(x, out=y)
            e1.record(st)
            torch.cuda.synchronize()
            print(f"spmv {what} {w} cfg={os.environ.get('SMLE_SPMV_CFG', 'default')}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
    elif what.startswith("spmm"):
        k = int(what[4:])
        X = torch.rand(n, k, dtype=torch.float64, device="cuda")
        Y = torch.empty_like(X)
        for _ in range(4):
            a.spmm(X, out=Y)
    elif what == "cg":
        b = torch.rand(n, dtype=torch.float64, device="cuda")
        x = torch.empty_like(b)
        print(a.cg_profile(b, x, 6))
    elif what == "cgmulti":
        B = torch.rand(n, 32, dtype=torch.float64, device="cuda")
        X = torch.empty_like(B)
        print(a.cg_profile(B, X, 3))
torch.cuda.synchronize()
print("ok", what, n)
