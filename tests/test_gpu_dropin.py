"""GPU: the drop-in boundary against the reference's REAL headers and call sites.

oracle/Makefile compiles, where /root/reference exists, the reference's own drivers twice from the
sources where they lie (nothing is copied): once as they are (ref_*: the CPU reference), once with
the single change INTEGRATION.md describes -- the callee renamed to the adapter of the GPU library
(dropin_*: cpu_singlecg.cpp:101 TestCGSolveSingle -> TestGpuCGSolveSingle;
verification/efficiency/parallel_efficiency.cpp:102 TestCGMultipleRHS -> TestGpuCGMultipleRHS).
The drop-in translation units include the reference's sparse_matrix.h (CsrMatrix<double,int>),
utils.h (CommandLineArgs), work_2025/types.hpp (int2, SpmmKernel) and hyper_parameters.hpp, so a
build proves the adapters accept the real types; here both binaries run on the same Matrix Market
file and their CSVs are compared.  The binaries travel to the GPU box with the snapshot
(oracle/_ref is git-ignored, not gpurun-ignored); /root/reference is not needed at run time."""
import csv
import os
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
BIN = ROOT / "sparse-matrix-linear-equations_b200" / "bin"


def _need(*names):
    missing = [n for n in names if not (REF / n).exists()]
    if missing:
        pytest.skip(f"oracle/_ref/{missing} not built (reference sources were absent at build time)")


def _run(cmd, cwd):
    env = dict(os.environ, OMP_NUM_THREADS="8", OMP_PROC_BIND="true", OMP_PLACES="cores")
    r = subprocess.run([str(c) for c in cmd], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (cmd, r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


def _rows(path):
    with open(path) as f:
        return list(csv.DictReader(f))


def _close(a, b):
    return abs(a - b) <= max(1, round(0.02 * b))


def test_cpu_singlecg_call_site_on_the_gpu(gpu, tmp_path):
    _need("dropin_singlecg", "ref_singlecg")
    mtx = tmp_path / "poisson3d_24.mtx"
    _run([BIN / "mtx_tool", "--grid3d=24", "--poisson", f"--out={mtx}"], tmp_path)
    out_gpu = _run([REF / "dropin_singlecg", f"--mtx={mtx}", f"--output={tmp_path / 'gpu.csv'}", "--threads=8", "--quiet"], tmp_path)
    out_cpu = _run([REF / "ref_singlecg", f"--mtx={mtx}", f"--output={tmp_path / 'cpu.csv'}", "--threads=8", "--quiet"], tmp_path)
    g, c = _rows(tmp_path / "gpu.csv"), _rows(tmp_path / "cpu.csv")
    assert len(g) == len(c) == 1                                    # num_vectors_list = {16} (cpu_singlecg.cpp:160)
    for key in ("matrix_name", "kernel", "num_vectors"):
        assert g[0][key] == c[0][key]
    assert g[0]["num_vectors"] == "16"
    # total iterations over the 16 vectors, driver-threshold tolerance (cpu_singlecg.cpp:92,101)
    assert _close(int(g[0]["iterations"]), int(c[0]["iterations"])), (g, c)
    assert "Rows: 13824" in out_gpu and "Rows: 13824" in out_cpu
    assert float(g[0]["gflops"]) > 0


def test_parallel_efficiency_call_site_on_the_gpu(gpu, tmp_path):
    _need("dropin_parallel_efficiency", "ref_parallel_efficiency")
    d = tmp_path / "mtx"
    d.mkdir()
    _run([BIN / "mtx_tool", "--grid3d=16", "--poisson", f"--out={d / 'poisson3d_16.mtx'}"], tmp_path)
    _run([BIN / "mtx_tool", "--grid2d=40", "--poisson", f"--out={d / 'poisson2d_40.mtx'}"], tmp_path)
    for name, exe in (("gpu", "dropin_parallel_efficiency"), ("cpu", "ref_parallel_efficiency")):
        _run([REF / exe, f"--mtx_dir={d}", f"--output_dir={tmp_path / name}", "--num_vectors=4", "--timing_iters=1"], tmp_path)
    g = _rows(tmp_path / "gpu" / "parallel_efficiency_detailed.csv")
    c = _rows(tmp_path / "cpu" / "parallel_efficiency_detailed.csv")
    assert len(g) == len(c) == 2 * 11                               # 2 matrices x thread_counts (:305)
    for rg, rc in zip(g, c):
        assert (rg["matrix_name"], rg["num_threads"]) == (rc["matrix_name"], rc["num_threads"])
        # CGSolveMultiple, raw 1e-5, SpmmKernel NONZERO_SPLIT (:300-303): lock-step iteration count
        assert _close(int(rg["iterations"]), int(rc["iterations"])), (rg, rc)
    assert len(_rows(tmp_path / "gpu" / "parallel_efficiency.csv")) == 11
