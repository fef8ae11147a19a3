"""Sweep SpMV kernel configurations (SMLE_SPMV_CFG) in sub-processes and time SpMV / CG kernels.
usage: python tools/sweep_spmv.py [grid_width] cfg cfg ...   (cfg = <ipt>x<stages>)"""
import os
import subprocess
import sys

CHILD = r'''
import sys, os
sys.path.insert(0, "sparse-matrix-linear-equations_b200/python")
import torch, smle_b200 as S
w = int(sys.argv[1])
S.init(0); st = torch.cuda.Stream(); S.set_stream(st.cuda_stream)
ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
n = len(ro) - 1; nnz = len(ci)
a = S.CsrMatrix(ro, ci, va)
with torch.cuda.stream(st):
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    for _ in range(5): a.spmv(x, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): a.spmv(x, out=y)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    b = torch.rand(n, dtype=torch.float64, device="cuda"); xs = torch.empty_like(b)
    k = a.cg_profile(b, xs, 30)
    a.cg_run_fixed(b.view(n, 1), xs.view(n, 1), 32); torch.cuda.synchronize()
    e0.record(st); a.cg_run_fixed(b.view(n, 1), xs.view(n, 1), 320); e1.record(st); torch.cuda.synchronize()
    it_ms = e0.elapsed_time(e1) / 320
bytes_ = nnz * 12 + (n + 1) * 4 + 2 * n * 8
print(f"cfg={os.environ.get('SMLE_SPMV_CFG','default'):6s} spmv {ms*1e3:7.1f} us {bytes_/ms/1e6:7.0f} GB/s | cg kernels us {k[0]*1e3:6.1f} {k[1]*1e3:6.1f} {k[2]*1e3:6.1f} | graph iter {it_ms*1e3:6.1f} us")
'''

w = sys.argv[1] if len(sys.argv) > 1 else "150"
for cfg in sys.argv[2:] or ["256x12x2"]:
    env = dict(os.environ, SMLE_SPMV_CFG=cfg)
    r = subprocess.run([sys.executable, "-c", CHILD, w], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
