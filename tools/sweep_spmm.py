"""Sweep SpMM kernel configurations in sub-processes and time SpMM / multi-RHS CG kernels.
usage: python tools/sweep_spmm.py [grid_width] [k] cfg[@chunk] ...
cfg = <threads>x<tile>x<stages>x<minb> (SMLE_SPMM_CFG), chunk = SMLE_SPMM_CHUNK, "v1" = merge-walk kernel,
"sched0" / "sched1" = default configuration without / with the structure-aware tile schedule"""
import os
import subprocess
import sys

CHILD = r'''
import sys, os
sys.path.insert(0, "sparse-matrix-linear-equations_b200/python")
import torch, smle_b200 as S
w = int(sys.argv[1]); k = int(sys.argv[2])
S.init(0); st = torch.cuda.Stream(); S.set_stream(st.cuda_stream)
ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
n = len(ro) - 1; nnz = len(ci)
a = S.CsrMatrix(ro, ci, va)
with torch.cuda.stream(st):
    X = torch.rand(n, k, dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    for _ in range(3): a.spmm(X, out=Y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): a.spmm(X, out=Y)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kk = a.cg_profile(X, Y, 4) if os.environ.get("SWEEP_CG") else (0, 0, 0)
bytes_ = nnz * 12 + (n + 1) * 4 + 2 * n * k * 8
print(f"{os.environ.get('SWEEP_TAG',''):24s} spmm {ms*1e3:8.1f} us {bytes_/ms/1e6:7.0f} GB/s {2*nnz*k/ms/1e6:7.0f} GFLOP/s | cg kernels us {kk[0]*1e3:7.1f} {kk[1]*1e3:7.1f} {kk[2]*1e3:7.1f}")
'''

w = sys.argv[1] if len(sys.argv) > 1 else "200"
k = sys.argv[2] if len(sys.argv) > 2 else "32"
for spec in sys.argv[3:] or ["256x1024x2x2@2"]:
    env = dict(os.environ, SWEEP_TAG=spec)
    if spec == "v1":
        env["SMLE_SPMM_V1"] = "1"
    elif spec.startswith("band"):        # band-window variant (k = 32 fp64), band<chunk>
        env["SMLE_SPMM_BAND"] = "1"
        if spec[4:]:
            env["SMLE_SPMM_BAND_CHUNK"] = spec[4:]
    elif spec.startswith("carve"):       # default configuration, shared-memory carve-out: carve0 = driver's choice, carve<pct>
        env["SMLE_SPMM_CARVE"] = spec[5:]
    elif spec in ("sched0", "sched1"):   # default configuration without / with the structure-aware tile schedule
        env["SMLE_SPMM_SCHED"] = spec[-1]
    else:
        cfg, _, chunk = spec.partition("@")
        env["SMLE_SPMM_CFG"] = cfg
        if chunk:
            env["SMLE_SPMM_CHUNK"] = chunk
    r = subprocess.run([sys.executable, "-c", CHILD, w, k], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr.strip()[-600:], flush=True)
    for line in r.stderr.splitlines():
        if line.startswith("[smle]"):
            print("    " + line, flush=True)
