#!/usr/bin/env python
"""bench.py -- headline benchmark of the merge-path SpMV / SpMM / CG hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Default workload (BASELINE.json configs[1], the `cpu_singlecg` path): single-RHS CG, fp64, 3-D
7-point Poisson grid 150^3 (3 375 000 rows, 23 490 000 nonzeros).  One STEP is what the
reference's TestCGSolveSingle times (single_strategy.hpp:215-239): L right-hand sides taken from
one srand(42) stream (cpu_singlecg.cpp:88-90), each a contiguous length-n vector, solved one
after another with CGSolveSingle.  Metric: CG iterations per second (total SpMV applications
over all vectors / time).  With N GPUs every rank solves its own L vectors of the stream
(independent units, no data-path collective): weak scaling, value = sum over ranks.

Other workloads (my own measurements of the remaining BASELINE configs; not the driver's line):
    --workload spmv     configs[0] merge-path SpMV, grid2d 1000^2 (+ the cold-cache rotation)
    --workload multicg  configs[2] multi-RHS CG k=32 on 200^3, columns sharded over ranks
    --workload stress   configs[3] RMAT / wheel SpMV & SpMM sweep

The JSON line carries `roofline` (dominant kernel = SpMV+dot merge kernel, algorithmic bytes /
CUDA-event time against MEASURED_PEAKS.json) and `cpu_baseline` (the reference's own OpenMP
CGSolveSingle from oracle/_ref, or the oracle port, on this box's host cores).
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
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))

GRID = 150            # configs[1]: 3-D Poisson 150^3
TOL = 1e-5            # raw relative tolerance (the drivers' default --tolerance)
MAX_ITERS = 10000     # cpu_singlecg.cpp:226
VEC_PER_STEP = 4      # right-hand sides solved per step and per GPU (the driver uses L = 16)
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


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_backend():
    from oracle import oracle as O
    O.build()
    r = O.ref()
    return (r, "reference") if r is not None else (O.port(), "port")


def cpu_cg_sample(ro, ci, va, b, iters_cap):
    """CGSolveSingle on one vector, capped at iters_cap iterations -> (seconds, iterations)."""
    be, kind = cpu_backend()
    cores = HOST_CORES
    be.set_threads(cores)
    t0 = time.perf_counter()
    it, _ = be.cg_single(ro, ci, va, b, iters_cap, TOL)
    dt = time.perf_counter() - t0
    return dt, it, kind, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    import smle_b200 as S
    ro, ci, va = S.gen_grid3d(GRID, True, 6.0, -1.0)
    n = len(ro) - 1
    b = S.gen_rhs_rand(42, n)
    cap = 40   # bounded sample: the first 40 CG iterations of vector 0 per step
    for _ in range(args.warmup):
        cpu_cg_sample(ro, ci, va, b, cap)
    times, iters = [], 0
    kind, cores = "port", 1
    for _ in range(args.steps):
        dt, it, kind, cores = cpu_cg_sample(ro, ci, va, b, cap)
        times.append(dt); iters += it
    total = sum(times)
    value = iters / total
    sample = f"first {cap} CG iterations of RHS vector 0 per step (CGSolveSingle, {cores} OpenMP threads)"
    line = {
        "impl": "reference", "metric": "cg_iterations_per_s", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, len(ci), args.gpus),
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


def workload_config(n, nnz, gpus):
    return {"workload": f"single-RHS CG (cpu_singlecg path) fp64, 3-D 7-point Poisson {GRID}^3 "
                        f"(InitGrid3d(w,true), diag 6 / off-diag -1), RHS srand(42) stream",
            "rows": n, "nnz": nnz, "rhs_vectors_per_step_per_gpu": VEC_PER_STEP, "tolerance": TOL,
            "max_iters": MAX_ITERS, "sharding": f"independent RHS vectors over {gpus} rank(s), no collective",
            "cache": "inputs larger than L2 (A 295 MB + 5 vectors 135 MB vs 126 MB L2); no flush"}


# ----------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------
def run_singlecg(args):
    import torch
    import torch.distributed as dist
    import smle_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    S.init(local)
    stream = torch.cuda.Stream()
    S.set_stream(stream.cuda_stream)

    ro, ci, va = S.gen_grid3d(GRID, True, 6.0, -1.0)
    n, nnz = len(ro) - 1, len(ci)
    a = S.CsrMatrix(ro, ci, va)
    L = VEC_PER_STEP
    # one srand(42) stream, cut into contiguous length-n vectors (cpu_singlecg.cpp:88-90, column-major);
    # rank r takes vectors [r*L, (r+1)*L)
    stream_all = S.gen_rhs_rand(42, n * L * world)
    b_host = torch.from_numpy(stream_all[rank * L * n:(rank + 1) * L * n].reshape(L, n).copy()).pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()
    b_dev = b_host.cuda()
    x_dev = torch.empty_like(b_dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        it = 0
        for v in range(L):
            i, _, _ = a.cg_solve_single(b_dev[v], MAX_ITERS, TOL, out=x_dev[v])
            it += i
        return it

    def step_host():
        # the reference-facing call for this workload: TestCGSolveSingle's loop over the L vectors with
        # HOST buffers (smle_cg_single_batch_f64); uploads of b and downloads of x are inside the call
        its, _ = a.cg_solve_single_batch(b_host, MAX_ITERS, TOL, out=x_host)
        return sum(its)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        # ---- timed region: device-resident inputs ------------------------------------------------
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        l0 = S.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        iters = 0
        for _ in range(args.steps):
            iters += step_device()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        launches = S.launch_count() - l0
        ms = e0.elapsed_time(e1)

        # ---- end-to-end: host (pinned) buffers through the C ABI, copies inside the timed region ----
        step_host()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        iters_h = 0
        for _ in range(args.steps):
            iters_h += step_host()
        h1.record(stream)
        barrier()
        ms_h = h0.elapsed_time(h1)

        # ---- per-kernel timing for the roofline (CUDA events around each launch, no graph) ----------
        kms = a.cg_profile(b_dev[0], x_dev[0], 40)
    torch.cuda.synchronize()

    t = torch.tensor([ms, ms_h], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(iters), float(iters_h), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, ms_h = t.tolist()
    iters_all, iters_h_all, launches_all = cnt.tolist()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        k1_bytes = spmm_bytes(n, n, nnz, 1)
        achieved = k1_bytes / (kms[0] * 1e-3) / 1e9
        iter_ms = ms / (iters_all / world)
        line = {
            "metric": "cg_iterations_per_s", "value": iters_all / (ms * 1e-3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, nnz, world),
            "iterations_per_step_per_gpu": iters_all / world / args.steps,
            "gflops": (2.0 * nnz + 10.0 * n) * iters_all / (ms * 1e-3) / 1e9,   # cpu_singlecg.cpp:94,108
            "clocks": clocks,
            "e2e": {"value": iters_h_all / (ms_h * 1e-3), "unit": "iter/s",
                    "h2d_bytes_per_step": L * n * 8, "d2h_bytes_per_step": L * n * 8,
                    "api": "smle_cg_single_batch_f64(host b_vectors -> host x_solutions), pinned buffers, copies overlapped with the neighbouring solves"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "spmv_kernel<double,480,6,2,DOT> (TMA-staged merge-path SpMV + p.Ap)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("spmv_dot_grid3d_150"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes,
                         "kernel_ms": {"spmv_dot": kms[0], "update_r_dot": kms[1], "update_xp": kms[2]},
                         "how": "CUDA events around every launch of 40 un-graphed CG iterations on the launch stream",
                         "iteration": {"ms": iter_ms, "algorithmic_bytes": cg_iter_bytes(n, nnz, 1),
                                       "achieved": cg_iter_bytes(n, nnz, 1) / (iter_ms * 1e-3) / 1e9,
                                       "frac": cg_iter_bytes(n, nnz, 1) / (iter_ms * 1e-3) / 1e9 / peak}},
        }
        if world == 1 and not args.no_cpu:
            cap = 40
            cpu_cg_sample(ro, ci, va, stream_all[:n].copy(), 5)
            dt, it, kind, cores = cpu_cg_sample(ro, ci, va, stream_all[:n].copy(), cap)
            line["cpu_baseline"] = {"value": it / dt, "unit": "iter/s", "cores": cores, "kind": kind,
                                    "sample": f"first {cap} CG iterations of RHS vector 0 (CGSolveSingle, OpenMP)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="singlecg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args, _ = ap.parse_known_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "singlecg":
        return run_singlecg(args)
    import bench_extra
    return bench_extra.run(args)


if __name__ == "__main__":
    main()
