#!/usr/bin/env python
"""bench.py -- headline benchmark of the merge-path SpMV / SpMM / CG hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Default workload (BASELINE.json configs[4], the metric's "CG iterations/s at 1/2/4/8 B200"):
ROW-PARTITIONED single-RHS fp64 CG on the 3-D 7-point Poisson grid 300^3 (27 000 000 rows,
188 460 000 nonzeros), solver semantics of the reference's CGSolveSingle
(work_2025/main/single_strategy.hpp:105-170) on the global system.  The SAME global problem is
solved at every N (strong scaling): rows are cut at the reference's merge-path coordinates, every
rank generates and owns only its slab, the halo of p travels by NVLink peer stores and the two dot
products per iteration by flag-in-data mailboxes -- no library collective on the data path.  One
STEP = one full solve A x = b (b = the srand(42) stream of cpu_singlecg.cpp:88-90, raw tolerance
1e-5, max_iters 10000).  Metric: CG iterations per second = SpMV applications / time.

Before the timed region the same ranks solve a 40^3 system and compare iteration count (+-2 %)
and solution (1e-6) with the CPU oracle: `"parity_check": "ok"` in the JSON line, or the run aborts.

The line also carries, as `extra.*`, this round's numbers for BASELINE.json configs[1] (single-RHS
CG 150^3 on every rank, independent right-hand sides) and configs[2] (multi-RHS CG k=32 on 200^3,
columns sharded over the N ranks); `--no-extras` skips them.

Other workloads (builder measurements of the remaining configs; not the driver's line):
    --workload singlecg configs[1] as the main line (round-1 headline)
    --workload spmv     configs[0] merge-path SpMV, grid2d 1000^2 (+ the cold-cache rotation)
    --workload multicg  configs[2] multi-RHS CG k=32 on 200^3, columns sharded over ranks
    --workload stress   configs[3] RMAT / wheel SpMV & SpMM sweep
    --workload rowcg_fixed  configs[4] with a fixed iteration count + the per-rank kernel timeline

`roofline`: dominant kernel = the TMA-staged merge-path SpMV + p.Ap (spmv_kernel<DOT>) on the rank's
slab, algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json.  `cpu_baseline` / `--impl
reference`: the reference's own OpenMP CGSolveSingle (oracle/_ref, the unmodified sources; else the
oracle port) on this box's host cores, on the same 300^3 system.
"""
from __future__ import annotations

import argparse
import json
import os

# Host cores and OpenMP binding must be settled before any library that embeds an OpenMP runtime
# is loaded: with OMP_PROC_BIND=true the runtime pins the main thread, after which the affinity
# mask no longer tells how many cores the process may use.  (The reference wrappers bind with
# KMP_AFFINITY=granularity=core,scatter; SURVEY.md section 6 measured 8x without binding.)
import sys

HOST_CORES = len(os.sched_getaffinity(0))
if int(os.environ.get("WORLD_SIZE", "1")) == 1 or "reference" in sys.argv:
    # only the process that times the CPU reference binds its OpenMP team; under torchrun the
    # ranks must not all pin their launch threads to core 0
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ.setdefault("OMP_PLACES", "cores")
    os.environ["OMP_NUM_THREADS"] = str(HOST_CORES)
import statistics
import subprocess
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))

GRID = int(os.environ.get("SMLE_BENCH_GRID", "300"))   # configs[4]: 3-D Poisson 300^3
PARITY_GRID = 40      # in-run oracle check on the same ranks
C2_GRID = 150         # configs[1]
C3_GRID = int(os.environ.get("SMLE_BENCH_C3_GRID", "200"))   # configs[2]
C3_K = 32
TOL = 1e-5            # raw relative tolerance (the drivers' default --tolerance)
MAX_ITERS = 10000     # cpu_singlecg.cpp:226
VEC_PER_STEP = 4      # configs[1] extra: right-hand sides per step and per GPU (the driver uses L = 16)
REF_BUDGET_S = float(os.environ.get("SMLE_REF_BUDGET_S", "240"))   # CPU time the reference arm may spend in steps
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# byte models (SURVEY.md section 8d / BASELINE.md section 3), V = 8 bytes
# ----------------------------------------------------------------------------------------------
def spmm_bytes(m, n, nnz, k, V=8):
    return nnz * (V + 4) + (m + 1) * 4 + n * k * V + m * k * V


def cg_iter_bytes(m, nnz, k, V=8):
    return nnz * (V + 4) + (m + 1) * 4 + 11 * m * k * V


def grid3d_shape(w):
    """rows, nnz of InitGrid3d(w, true) (sparse_matrix.h:541-550)"""
    return w ** 3, 7 * w ** 3 - 6 * w * w


def workload_config(world):
    n, nnz = grid3d_shape(GRID)
    return {"workload": f"row-partitioned single-RHS CG (CGSolveSingle semantics) fp64, 3-D 7-point Poisson {GRID}^3 "
                        f"(InitGrid3d(w,true), diag 6 / off-diag -1), b = srand(42) stream; one step = one full solve",
            "rows": n, "nnz": nnz, "tolerance": TOL, "max_iters": MAX_ITERS, "rhs_vectors_per_step": 1,
            "reference_driver_rhs_vectors": 16,   # cpu_singlecg.cpp:160 solves L = 16 vectors per timing pass
            "sharding": f"rows cut at MergePathSearch(g*ceil((m+nnz)/G)).x over {world} rank(s); every rank generates "
                        f"and owns only its slab; halo of p by NVLink peer stores (fused into the p update), p.Ap and "
                        f"r.r by flag-in-data mailboxes in peer memory, summed in rank order; no NCCL on the data path",
            "cache": "inputs larger than L2 at N = 1 (A 2.4 GB + 5 vectors 1.1 GB vs 126 MB L2); no flush"}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_backend():
    from oracle import oracle as O
    O.build()
    r = O.ref()
    return (r, "reference") if r is not None else (O.port(), "port")


def cpu_system(w):
    """the bench system from the CHECKER's own generators (no product code in the CPU legs)"""
    from oracle import oracle as O
    orc = O.port()
    ro, ci, va = orc.gen_grid3d_sorted(w, True, 6.0, -1.0)
    b = orc.rhs_rand(42, len(ro) - 1)
    return ro, ci, va, b


def cpu_cg_step(be, kind, ro, ci, va, b, cap):
    """one CPU step: CGSolveSingle capped at `cap` iterations -> (seconds, iterations).
    With oracle/_ref the reference's own wrapper TestCGSolveSingle (single_strategy.hpp:179-240: one
    vector, one timing pass, its CpuTimer) does the timing."""
    if kind == "reference":
        ms, it, _ = be.test_cg_single(ro, ci, va, b, 1, cap, TOL, 1)
        return ms * 1e-3, int(it)
    t0 = time.perf_counter()
    it, _ = be.cg_single(ro, ci, va, b, cap, TOL)
    return time.perf_counter() - t0, it


def cpu_iter_cap(be, kind, ro, ci, va, b, budget_s, steps):
    """iterations per step such that `steps` steps fit in budget_s (calibrated on 4 iterations)"""
    cpu_cg_step(be, kind, ro, ci, va, b, 2)
    dt, it = cpu_cg_step(be, kind, ro, ci, va, b, 4)
    per_iter = dt / max(it, 1)
    return max(8, min(MAX_ITERS, int(budget_s / max(steps, 1) / per_iter))), per_iter


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    be, kind = cpu_backend()
    cores = HOST_CORES
    be.set_threads(cores)
    ro, ci, va, b = cpu_system(GRID)
    n, nnz = len(ro) - 1, len(ci)
    total_steps = args.steps + args.warmup
    cap, per_iter = cpu_iter_cap(be, kind, ro, ci, va, b, REF_BUDGET_S, total_steps)
    for _ in range(args.warmup):
        cpu_cg_step(be, kind, ro, ci, va, b, cap)
    times, iters = [], []
    for _ in range(args.steps):
        dt, it = cpu_cg_step(be, kind, ro, ci, va, b, cap)
        times.append(dt); iters.append(it)
    total = sum(times)
    value = sum(iters) / total
    full = all(it < cap for it in iters)
    sample = (f"every step is the full solve ({iters[0]} iterations to tol {TOL})" if full else
              f"every step is the first {cap} iterations of the solve (bounded so that {total_steps} steps fit "
              f"{REF_BUDGET_S:.0f} s of CPU time; the rate is per iteration)")
    sample += (f"; CGSolveSingle, {cores} OpenMP threads, OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')} "
               f"OMP_PLACES={os.environ.get('OMP_PLACES')}; timed by " +
               ("the reference's TestCGSolveSingle wrapper" if kind == "reference" else "perf_counter around the oracle port") +
               f"; fastest step {max(i / t for i, t in zip(iters, times)):.1f} iter/s")
    line = {
        "impl": "reference", "metric": "cg_iterations_per_s", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "iterations_per_step": sum(iters) / max(args.steps, 1),
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json names the report it was read from); None
    when no capture of this workload is committed."""
    try:
        rec = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text())[key]
        return int(rec["dram_bytes_read"]) + int(rec["dram_bytes_write"])
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    """process group, device, stream"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import smle_b200 as S
        self.torch, self.dist, self.S = torch, dist, S
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        S.init(self.local)
        self.stream = torch.cuda.Stream()
        S.set_stream(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def timed(self, fn, steps):
        """barrier, CUDA events on the launch stream around `steps` calls of fn, barrier -> (ms, sum of fn())"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        acc = 0
        for _ in range(steps):
            acc += fn()
        e1.record(self.stream)
        self.barrier()
        return e0.elapsed_time(e1), acc

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def parity_check(cx):
    """the same ranks, a PARITY_GRID^3 system: iteration count within 2 % and solution within 1e-6 of
    the CPU oracle's CGSolveSingle restatement on the global system.  Raises on mismatch."""
    from oracle import oracle as O
    from smle_b200 import dist as D
    torch, S = cx.torch, cx.S
    orc = O.port()
    A = D.RowPartitionedCsr.grid3d(PARITY_GRID, cx.rank, cx.world, cx.gather)
    ro, ci, va = orc.gen_grid3d_sorted(PARITY_GRID, True, 6.0, -1.0)
    b = orc.rhs_rand(42, len(ro) - 1)
    it_ref, x_ref = orc.cg_single(ro, ci, va, b, MAX_ITERS, TOL)
    b_local = torch.from_numpy(S.gen_rhs_rand_range(42, A.r0, A.n_local)).cuda()
    with torch.cuda.stream(cx.stream):
        it, x, rel = A.cg_solve_single(b_local, MAX_ITERS, TOL)
    cx.barrier()
    err = float(np.abs(x.cpu().numpy() - x_ref[A.r0:A.r1]).max() / np.abs(x_ref).max())
    A.close()
    ok = abs(it - it_ref) <= max(1, round(0.02 * it_ref)) and err <= 1e-6
    oks = cx.gather((ok, it, it_ref, err))
    if not all(o[0] for o in oks):
        raise RuntimeError(f"parity check against the oracle failed on {PARITY_GRID}^3: (ok, iters, oracle iters, err) per rank = {oks}")
    return {"status": "ok", "grid": PARITY_GRID, "iterations": it, "oracle_iterations": it_ref,
            "max_solution_err": max(o[3] for o in oks)}


def extra_c2(cx, steps):
    """configs[1]: single-RHS CG on 150^3, every rank its own right-hand sides (round-1 headline)"""
    torch, S = cx.torch, cx.S
    ro, ci, va = S.gen_grid3d(C2_GRID, True, 6.0, -1.0)
    n, nnz = len(ro) - 1, len(ci)
    a = S.CsrMatrix(ro, ci, va)
    L = VEC_PER_STEP
    b_host = torch.from_numpy(S.gen_rhs_rand_range(42, cx.rank * L * n, L * n).reshape(L, n)).pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()
    b_dev = b_host.cuda()
    x_dev = torch.empty_like(b_dev)

    def step_device():
        return sum(a.cg_solve_single(b_dev[v], MAX_ITERS, TOL, out=x_dev[v])[0] for v in range(L))

    def step_host():
        return sum(a.cg_solve_single_batch(b_host, MAX_ITERS, TOL, out=x_host)[0])

    with torch.cuda.stream(cx.stream):
        step_device()
        ms, iters = cx.timed(step_device, steps)
        step_host()
        ms_h, iters_h = cx.timed(step_host, steps)
        kms = a.cg_profile(b_dev[0], x_dev[0], 20)
    ms, ms_h = cx.reduce([ms, ms_h], "MAX")
    it_all, ith_all = cx.reduce([float(iters), float(iters_h)], "SUM")
    a.close()
    peak, _ = measured_hbm_peak()
    iter_ms = ms / (it_all / cx.world)
    return {"workload": f"configs[1]: single-RHS CG fp64 3-D Poisson {C2_GRID}^3, {L} RHS vectors per step on each of {cx.world} GPU(s) (independent units)",
            "value": it_all / (ms * 1e-3), "unit": "iter/s", "scaling": "weak", "steps": steps,
            "e2e": {"value": ith_all / (ms_h * 1e-3), "unit": "iter/s", "h2d_bytes_per_step": L * n * 8, "d2h_bytes_per_step": L * n * 8},
            "ms_per_iteration": iter_ms, "iteration_frac_of_measured_hbm": cg_iter_bytes(n, nnz, 1) / (iter_ms * 1e-3) / 1e9 / peak,
            "kernel_ms": {"spmv_dot": kms[0], "update_r_dot": kms[1], "update_xp": kms[2]},
            "spmv_dot_frac_of_measured_hbm": spmm_bytes(n, n, nnz, 1) / (kms[0] * 1e-3) / 1e9 / peak}


def extra_c3(cx, iters=32):
    """configs[2]: multi-RHS CG k=32 on 200^3, columns sharded over the ranks, A replicated"""
    torch, S = cx.torch, cx.S
    K, kloc = C3_K, C3_K // cx.world
    ro, ci, va = S.gen_grid3d(C3_GRID, True, 6.0, -1.0)
    n, nnz = len(ro) - 1, len(ci)
    a = S.CsrMatrix(ro, ci, va)
    del ro, ci, va
    # row-major n x 32 block from the srand(42) stream; rank r owns columns [r*kloc, (r+1)*kloc)
    Bfull = S.gen_rhs_rand(42, n * K).reshape(n, K)
    B = torch.from_numpy(np.ascontiguousarray(Bfull[:, cx.rank * kloc:(cx.rank + 1) * kloc])).cuda()
    del Bfull
    X = torch.empty_like(B)
    with torch.cuda.stream(cx.stream):
        a.cg_run_fixed(B, X, 16)

        def run():
            a.cg_run_fixed(B, X, iters)
            return iters
        ms, _ = cx.timed(run, 1)
        kms = a.cg_profile(B, X, 4)
    ms = cx.reduce([ms], "MAX")[0] / iters
    a.close()
    peak, _ = measured_hbm_peak()
    per_gpu_bytes = nnz * 12 + (n + 1) * 4 + 11 * n * kloc * 8
    return {"workload": f"configs[2]: multi-RHS CG k={K} fp64 3-D Poisson {C3_GRID}^3, columns sharded over {cx.world} GPU(s), "
                        f"{iters} lock-step iterations (fixed count)",
            "value": 1e3 / ms, "unit": "iter/s", "scaling": "strong", "ms_per_iteration": ms,
            "gflops": (2.0 * nnz + 10.0 * n) * K / ms / 1e6,
            "per_gpu_frac_of_measured_hbm": per_gpu_bytes / ms / 1e6 / peak,
            "kernel_ms": {"spmm_dot": kms[0], "update_r_dot": kms[1], "update_xp": kms[2]}}


def run_rowcg(args):
    cx = Ctx()
    torch, S = cx.torch, cx.S
    from smle_b200 import dist as D
    world, rank = cx.world, cx.rank
    warmup = max(args.warmup, 3)

    parity = parity_check(cx)

    t_setup = time.time()
    A = D.RowPartitionedCsr.grid3d(GRID, rank, world, cx.gather)
    n, nnz = A.num_rows_global, A.num_nonzeros_global
    nloc, nnz_loc = A.n_local, A.plan.nnz_local
    b_host = torch.from_numpy(S.gen_rhs_rand_range(42, A.r0, nloc)).pin_memory()   # this rank's rows of the stream
    x_host = torch.empty_like(b_host).pin_memory()
    b_dev = b_host.cuda()
    x_dev = torch.empty_like(b_dev)
    setup_s = time.time() - t_setup

    def step_device():
        return A.cg_solve_single(b_dev, MAX_ITERS, TOL, out=x_dev)[0]

    def step_host():
        # the reference-facing call with HOST buffers: smle_dist_cg_f64(is_device_ptr = 0) uploads this rank's
        # rows of b and downloads its rows of x inside the call
        return A.cg_solve_single(b_host, MAX_ITERS, TOL, out=x_host)[0]

    with torch.cuda.stream(cx.stream):
        for _ in range(warmup):
            step_device()
        # ---- timed region: device-resident inputs ------------------------------------------------
        sampler = ClockSampler(cx.local)
        sampler.start()
        l0 = S.launch_count()
        ms, iters = cx.timed(step_device, args.steps)
        clocks = sampler.stop()
        launches = S.launch_count() - l0
        # ---- end-to-end: host (pinned) buffers through the C ABI, copies inside the timed region ----
        step_host()
        ms_h, iters_h = cx.timed(step_host, args.steps)
        # ---- per-kernel timing for the roofline (CUDA events around each launch, no graph) ----------
        cx.barrier()
        kms = A.cg_profile(b_dev, 40)
    torch.cuda.synchronize()
    x_check = float(torch.linalg.vector_norm(x_dev).item())   # the device -> host read of the result

    ms, ms_h, k1, k2, k3 = cx.reduce([ms, ms_h] + kms, "MAX")
    launches_all, = cx.reduce([float(launches)], "SUM")
    h2d, d2h = cx.reduce([float(nloc * 8), float(nloc * 8)], "SUM")
    k1_bytes_all = cx.gather(spmm_bytes(nloc, nloc + A.n_halo, nnz_loc, 1))
    A.close()

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        it_step = iters / args.steps
        iter_ms = ms / iters
        k1_bytes = max(k1_bytes_all)
        achieved = k1_bytes / (k1 * 1e-3) / 1e9
        line = {
            "metric": "cg_iterations_per_s", "value": iters / (ms * 1e-3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "parity_check": parity["status"], "parity": parity,
            "iterations_per_step": it_step, "ms_per_iteration": iter_ms,
            "gflops": (2.0 * nnz + 10.0 * n) * iters / (ms * 1e-3) / 1e9,   # cpu_singlecg.cpp:94,108
            "clocks": clocks,
            "e2e": {"value": iters_h / (ms_h * 1e-3), "unit": "iter/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "smle_dist_cg_f64(is_device_ptr=0): every rank uploads its rows of b from pinned host memory and "
                           "downloads its rows of x inside the call"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "spmv_kernel<double,480,6,2,DOT> on the rank's slab (TMA-staged merge-path SpMV + p.Ap)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(f"spmv_dot_grid3d_{GRID}_n{world}"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes,
                         "kernel_ms": {"spmv_dot": k1, "update_r_dot": k2, "update_xp": k3},
                         "how": "CUDA events around every launch of 40 un-graphed iterations on the launch stream, max over ranks; "
                                "at N > 1 the kernels' waits for the peers are inside these times",
                         "iteration": {"ms": iter_ms, "algorithmic_bytes": cg_iter_bytes(n, nnz, 1),
                                       "achieved_aggregate": cg_iter_bytes(n, nnz, 1) / (iter_ms * 1e-3) / 1e9,
                                       "frac_per_gpu": cg_iter_bytes(n, nnz, 1) / (iter_ms * 1e-3) / 1e9 / peak / world}},
            "setup_s": setup_s, "x_norm_local": x_check,
        }
    if not args.no_extras:
        c2 = extra_c2(cx, 2)
        c3 = extra_c3(cx)
        if rank == 0:
            line["extra"] = {"c2_singlecg_150": c2, "c3_multicg_200_k32": c3}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            be, kind = cpu_backend()
            be.set_threads(HOST_CORES)
            ro, ci, va, b = cpu_system(GRID)
            cap, _ = cpu_iter_cap(be, kind, ro, ci, va, b, 12.0, 1)   # ~12 s of CPU work per sample, two samples
            samples = [cpu_cg_step(be, kind, ro, ci, va, b, cap) for _ in range(2)]
            dt, it = min(samples, key=lambda s: s[0] / max(s[1], 1))
            line["cpu_baseline"] = {"value": it / dt, "unit": "iter/s", "cores": HOST_CORES, "kind": kind,
                                    "sample": f"first {it} CG iterations of the same {GRID}^3 solve (CGSolveSingle, OpenMP bound to cores), "
                                              f"best of 2 samples ({', '.join(f'{i / t:.1f}' for t, i in samples)} iter/s)"}
        print(json.dumps(line), flush=True)
    cx.close()


def run_singlecg(args):
    """configs[1] as the main line (round-1 headline): kept for builder measurements"""
    cx = Ctx()
    c2 = extra_c2(cx, args.steps)
    if cx.rank == 0:
        print(json.dumps(c2), flush=True)
    cx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rowcg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] / configs[2] extras")
    args, _ = ap.parse_known_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "rowcg":
        return run_rowcg(args)
    if args.workload == "singlecg":
        return run_singlecg(args)
    import bench_extra
    return bench_extra.run(args)


if __name__ == "__main__":
    main()
