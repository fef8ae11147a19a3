"""A/B of the general-tile paths of the single-vector kernel (SMLE_SPMV_DEBUG bits) on skewed matrices.

The parent generates each matrix once (host C++ generators), parks it in /dev/shm, and runs one child process
per variant (the switches are read once per process).  Every child times 20 products after 5 warm-ups with
CUDA events on the launch stream and leaves y behind; the parent reports the time per product, the fraction
of the measured HBM peak on the algorithmic model, and the largest difference of y against the first variant.

usage: python tools/ab_general_tiles.py [spec ...]      spec = rmat:<scale>[:f32] | wheel:<log2 spokes>[:f32]
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))

VARIANTS = [
    ("per-row gathers (SMLE_SPMV_DEBUG=4)", {"SMLE_SPMV_DEBUG": "4"}),
    ("product-staged (default)", {"SMLE_SPMV_DEBUG": "0"}),
    ("product-staged + L2 priorities (SMLE_SPMV_DEBUG=8)", {"SMLE_SPMV_DEBUG": "8"}),
]
SHM = Path("/dev/shm")


def child(tag):
    import torch
    import smle_b200 as S
    ro, ci, va = (np.load(SHM / f"smle_ab_{tag}_{n}.npy") for n in ("ro", "ci", "va"))
    S.init(0)
    st = torch.cuda.Stream()
    S.set_stream(st.cuda_stream)
    n = len(ro) - 1
    a = S.CsrMatrix(ro, ci, va)
    tdt = torch.float64 if va.dtype == np.float64 else torch.float32
    with torch.cuda.stream(st):
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.rand(n, dtype=tdt, device="cuda", generator=g)
        y = torch.empty_like(x)
        for _ in range(5):
            a.spmv(x, out=y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            a.spmv(x, out=y)
        e1.record(st)
        torch.cuda.synchronize()
    np.save(SHM / f"smle_ab_{tag}_y_{os.environ.get('SMLE_AB_VARIANT', '0')}.npy", y.cpu().numpy())
    print(json.dumps({"us": e0.elapsed_time(e1) / 20 * 1e3}))


def main():
    import smle_b200 as S
    specs = sys.argv[1:] or ["rmat:22", "rmat:23", "rmat:24", "rmat:24:f32", "wheel:24"]
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("hbm_gbs", 6459.0))
    for spec in specs:
        parts = spec.split(":")
        dtype = np.float32 if parts[-1] == "f32" else np.float64
        if parts[0] == "rmat":
            ro, ci, va = S.gen_rmat(int(parts[1]), 16, seed=42, dtype=dtype)
        else:
            ro, ci, va = S.gen_wheel(1 << int(parts[1]), 1.0, dtype)
        tag = spec.replace(":", "_")
        for n, arr in (("ro", ro), ("ci", ci), ("va", va)):
            np.save(SHM / f"smle_ab_{tag}_{n}.npy", arr)
        m, nnz, vb = len(ro) - 1, len(ci), va.dtype.itemsize
        alg = nnz * (4 + vb) + m * (4 + 2 * vb)          # CSR stream + x once + y once
        del ro, ci, va
        y0 = None
        for vi, (name, env) in enumerate(VARIANTS):
            e = dict(os.environ, SMLE_AB_VARIANT=str(vi), **env)
            r = subprocess.run([sys.executable, __file__, "--child", tag], env=e, capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode or not line:
                print(json.dumps({"matrix": spec, "variant": name, "error": (r.stderr or r.stdout)[-300:]}), flush=True)
                continue
            us = json.loads(line[-1])["us"]
            y = np.load(SHM / f"smle_ab_{tag}_y_{vi}.npy")
            if y0 is None:
                y0 = y
            diff = float(np.abs(y.astype(np.float64) - y0.astype(np.float64)).max() / max(np.abs(y0).max(), 1e-300))
            print(json.dumps({"matrix": spec, "rows": m, "nnz": nnz, "variant": name, "us": round(us, 1),
                              "algorithmic_GBs": round(alg / us / 1e3, 1), "frac_of_measured_hbm": round(alg / us / 1e3 / peak, 3),
                              "max_diff_vs_first_variant": diff}), flush=True)
        for f in SHM.glob(f"smle_ab_{tag}_*.npy"):
            f.unlink()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()
