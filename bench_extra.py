"""bench_extra.py -- measurements of the remaining BASELINE.json configs (not the driver's line).

    python bench.py --workload spmv     configs[0]: merge-path SpMV, grid2d 1000^2 (4-pt and 5-pt),
                                        warm (L2 resident, 80 MB < 126 MB L2) and cold (rotating
                                        buffers larger than L2)
    python bench.py --workload multicg  configs[2]: multi-RHS CG k=32 on 3-D Poisson 200^3 (columns
                                        sharded over ranks under torchrun)
    python bench.py --workload stress   configs[3]: RMAT / wheel SpMV and SpMM, fp32 + fp64, k in {1,8,32}

Each prints one JSON line per case with GFLOP/s (reference formulas) and the
fraction of the measured HBM peak on the algorithmic byte model of SURVEY.md section 8(d).
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))


def _peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"])
    except Exception:
        return 6650.0


def _time(fn, stream, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def spmm_bytes(m, n, nnz, k, V):
    return nnz * (V + 4) + (m + 1) * 4 + (n + m) * k * V


def _setup():
    import torch
    import smle_b200 as S
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    S.init(local)
    st = torch.cuda.Stream()
    S.set_stream(st.cuda_stream)
    return torch, S, st


def run_spmv(args):
    torch, S, st = _setup()
    peak = _peak()
    with torch.cuda.stream(st):
        for loop, name in ((False, "grid2d_1000 4-pt (reference driver: self_loop=false)"), (True, "grid2d_1000 5-pt")):
            ro, ci, va = S.gen_grid2d(1000, loop)
            m, nnz = len(ro) - 1, len(ci)
            a = S.CsrMatrix(ro, ci, va)
            x = torch.full((m,), 0.0019, dtype=torch.float64, device="cuda")
            y = torch.empty_like(x)
            warm_ms = _time(lambda: a.spmv(x, out=y), st, 200)
            # cold: rotate through copies of A and x so that nothing survives in the 126 MB L2
            copies = [(S.CsrMatrix(ro, ci, va), x.clone(), torch.empty_like(x)) for _ in range(4)]
            i = [0]

            def cold():
                b, xb, yb = copies[i[0] % 4]
                i[0] += 1
                b.spmv(xb, out=yb)
            cold_ms = _time(cold, st, 200)
            B = spmm_bytes(m, m, nnz, 1, 8)
            for tag, ms in (("warm(L2-resident)", warm_ms), ("cold(rotating 4 copies, 320 MB)", cold_ms)):
                print(json.dumps({"workload": f"merge SpMV fp64 {name}", "cache": tag, "ms": ms,
                                  "gflops": 2.0 * nnz / ms / 1e6, "algorithmic_GBs": B / ms / 1e6,
                                  "frac_of_measured_hbm": B / ms / 1e6 / peak}), flush=True)


def run_multicg(args):
    import torch.distributed as dist
    torch, S, st = _setup()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    w = int(os.environ.get("SMLE_MULTICG_GRID", "200"))
    K = 32
    kloc = K // world
    ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
    n, nnz = len(ro) - 1, len(ci)
    a = S.CsrMatrix(ro, ci, va)
    # row-major n x 32 block from the srand(42) stream; rank r owns columns [r*kloc, (r+1)*kloc)
    Bfull = S.gen_rhs_rand(42, n * K).reshape(n, K)
    B = torch.from_numpy(np.ascontiguousarray(Bfull[:, rank * kloc:(rank + 1) * kloc])).cuda()
    del Bfull
    X = torch.empty_like(B)
    with torch.cuda.stream(st):
        a.cg_run_fixed(B, X, 16)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        iters = 48
        a.cg_run_fixed(B, X, iters)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        kms = a.cg_profile(B, X, 5)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    if rank == 0:
        per_gpu_bytes = nnz * 12 + (n + 1) * 4 + 11 * n * kloc * 8
        print(json.dumps({"workload": f"multi-RHS CG k={K} 3-D Poisson {w}^3, columns sharded over {world} GPU(s)",
                          "ms_per_iteration": ms, "iterations_per_s": 1e3 / ms,
                          "gflops": (2.0 * nnz + 10.0 * n) * K / ms / 1e6,
                          "per_gpu_algorithmic_GBs": per_gpu_bytes / ms / 1e6,
                          "frac_of_measured_hbm": per_gpu_bytes / ms / 1e6 / _peak(),
                          "kernel_ms": {"spmm_dot": kms[0], "update_r_dot": kms[1], "update_xp": kms[2]}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_stress(args):
    torch, S, st = _setup()
    peak = _peak()
    scale = int(os.environ.get("SMLE_RMAT_SCALE", "24"))   # configs[3] names scale 24 / 2^24; smaller for quick runs
    with torch.cuda.stream(st):
        for mname, gen in ((f"rmat scale {scale} x16", lambda dt: S.gen_rmat(scale, 16, seed=42, dtype=dt)),
                           (f"wheel 2^{scale}", lambda dt: S.gen_wheel(1 << scale, 1.0, dt))):
            for dt, tdt, V in ((np.float64, torch.float64, 8), (np.float32, torch.float32, 4)):
                ro, ci, va = gen(dt)
                m, nnz = len(ro) - 1, len(ci)
                a = S.CsrMatrix(ro, ci, va)
                del ro, ci, va
                for k in (1, 8, 32):
                    X = torch.rand(m, k, dtype=tdt, device="cuda") if k > 1 else torch.rand(m, dtype=tdt, device="cuda")
                    Y = torch.empty_like(X)
                    fn = (lambda: a.spmm(X, out=Y)) if k > 1 else (lambda: a.spmv(X, out=Y))
                    ms = _time(fn, st, 10, warm=2)
                    B = spmm_bytes(m, m, nnz, k, V)
                    print(json.dumps({"workload": f"merge SpMV/SpMM {mname}", "dtype": "f64" if V == 8 else "f32", "k": k,
                                      "rows": m, "nnz": nnz, "ms": ms, "gflops": 2.0 * nnz * k / ms / 1e6,
                                      "algorithmic_GBs": B / ms / 1e6, "frac_of_measured_hbm": B / ms / 1e6 / peak}), flush=True)
                    del X, Y
                a.close()


def run_rowcg(args):
    """configs[4] with a FIXED iteration count (no convergence exit): per-iteration time and the
    per-kernel timeline (K1 / K2 / K3 including their waits for the peers) of the row-partitioned CG.
    The driver's line is bench.py's default workload; this is the builder's breakdown of it."""
    import time
    import torch.distributed as dist
    torch, S, st = _setup()
    from smle_b200 import dist as D
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)   # plumbing only (request blobs, IPC handles)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    w = int(os.environ.get("SMLE_ROWCG_GRID", "300"))
    iters = int(os.environ.get("SMLE_ROWCG_ITERS", "400"))
    t0 = time.time()
    A = D.RowPartitionedCsr.grid3d(w, rank, world, gather)
    m, nnz = A.num_rows_global, A.num_nonzeros_global
    b = torch.from_numpy(S.gen_rhs_rand_range(42, A.r0, A.n_local)).cuda()
    x = torch.empty_like(b)
    setup_s = time.time() - t0
    with torch.cuda.stream(st):
        A.cg_solve_single(b, 32, 1e-300, out=x)          # warm-up (graph build)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        it, _, rel = A.cg_solve_single(b, iters, 1e-300, out=x)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        dist.barrier()
        kms = A.cg_profile(b, 40)
    allms = gather((ms, kms))
    if rank == 0:
        ms = max(a[0] for a in allms)
        per_iter = ms / it
        bytes_iter = nnz * 12 + (m + 1) * 4 + 11 * m * 8
        print(json.dumps({"workload": f"row-partitioned CG fp64 3-D Poisson {w}^3 ({m} rows, {nnz} nnz) over {world} GPU(s)",
                          "iterations": it, "ms_per_iteration": per_iter, "iterations_per_s": 1e3 / per_iter,
                          "aggregate_algorithmic_GBs": bytes_iter / per_iter / 1e6,
                          "frac_of_measured_hbm_per_gpu": bytes_iter / per_iter / 1e6 / world / _peak(),
                          "kernel_us_per_rank_ungraphed": [[round(1e3 * v, 1) for v in a[1]] for a in allms],
                          "halo_entries_this_rank": A.n_halo, "rows_per_rank": [int(v) for v in np.diff(A.bounds)],
                          "final_rel_res": rel, "setup_s": setup_s, "scaling": "strong"}), flush=True)
    dist.barrier()
    A.close()
    dist.destroy_process_group()


def run(args):
    return {"spmv": run_spmv, "multicg": run_multicg, "stress": run_stress, "rowcg_fixed": run_rowcg}[args.workload](args)
