"""GPU: the gpu_* CLI drivers (reference flag surface) run end to end and agree with the oracle."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
BIN = ROOT / "sparse-matrix-linear-equations_b200" / "bin"


def _run(name, *flags, cwd=None):
    exe = BIN / name
    assert exe.exists(), f"{exe} not built (run __graft_entry__.build())"
    r = subprocess.run([str(exe), *flags], capture_output=True, text=True, cwd=cwd, timeout=180)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2000:])
    return r.stdout


def test_gpu_spmv_quiet_csv_and_pass(gpu):
    out = _run("gpu_spmv", "--grid2d=200", "--i=20")
    assert "PASS" in out and "FAIL" not in out
    out = _run("gpu_spmv", "--quiet", "--grid2d=200", "--i=20")
    fields = [f.strip() for f in out.strip().split(",") if f.strip()]
    # file, rows, cols, nnz, mean, std, variation, skew, method, setup, avg_ms, gflops, GB/s
    assert fields[0] == "grid2d_200" and fields[1] == "40000" and fields[3] == "159200"
    assert fields[8].startswith("Merge CsrMV") and float(fields[11]) > 0
    for flags in (("--wheel=5000",), ("--rmat=12",), ("--grid3d=20", "--fp32"), ("--dense=1024",)):
        assert "PASS" in _run("gpu_spmv", *flags, "--i=5")


def test_gpu_spmm_pass(gpu):
    for flags in (("--grid3d=24", "--num_vectors=32"), ("--wheel=3000", "--num_vectors=8", "--random_x"),
                  ("--grid2d=64", "--num_vectors=5", "--fp32")):
        out = _run("gpu_spmm", *flags, "--i=5")
        assert "PASS" in out and "FAIL" not in out


def test_gpu_singlecg_matches_reference_driver_semantics(gpu, orc, tmp_path):
    out = _run("gpu_singlecg", "--grid3d=24", "--num_vectors=3", f"--output={tmp_path}/r.csv")
    m = re.search(r"method=SINGLE_LOOP: [\d.]+ ms, (\d+) iters", out)
    assert m, out
    total = int(m.group(1))
    # oracle: same driver semantics -- column-major vectors of one srand(42) stream, threshold quirk
    ro, ci, va = orc.gen_grid3d(24, True, 6.0, -1.0)
    n = len(ro) - 1
    b = orc.rhs_rand(42, n * 3)
    thr = orc.driver_threshold(b, n, 1e-5)
    want = sum(orc.cg_single(ro, ci, va, b[v * n:(v + 1) * n].copy(), 10000, thr)[0] for v in range(3))
    assert abs(total - want) <= max(1, round(0.02 * want))
    csv = (tmp_path / "r.csv").read_text().splitlines()
    assert csv[0] == "matrix_name,kernel,num_vectors,min_ms,gflops,iterations"
    assert csv[1].startswith("grid3d_24,SINGLE_LOOP,3,")


def test_gpu_multicg_matches_oracle(gpu, orc, tmp_path):
    (tmp_path / "data" / "error_data").mkdir(parents=True)
    out = _run("gpu_multicg", "--grid3d=20", "--num_vectors=4", "--timing_iters=1", cwd=tmp_path)
    m = re.search(r"Iters:\s+([\d.]+)", out)
    assert m, out
    iters = float(m.group(1))
    ro, ci, va = orc.gen_grid3d(20, True, 6.0, -1.0)
    n = len(ro) - 1
    B = orc.rhs_rand(42, n * 4).reshape(n, 4)
    thr = orc.driver_threshold(B.ravel(), n, 1e-5)
    want, _, hist = orc.cg_multi(ro, ci, va, B, 4, 50000, thr, O.NONZERO_SPLIT, 8)
    assert abs(iters - want) <= max(1, round(0.02 * want))
    lines = (tmp_path / "data" / "error_data" / "grid3d_20_cg_errors.csv").read_text().splitlines()
    assert lines[0] == "iteration,max_error" and len(lines) - 1 == int(iters)
    np.testing.assert_allclose([float(l.split(",")[1]) for l in lines[1:6]], hist[:5], rtol=1e-5)


def test_mtx_roundtrip_through_gpu_driver(gpu, tmp_path):
    _run("mtx_tool", "--grid2d=30", "--poisson", f"--out={tmp_path}/g.mtx")
    out = _run("gpu_spmv", f"--mtx={tmp_path}/g.mtx", "--i=5")
    assert "PASS" in out


def test_gpu_multicg_sweep_is_the_missing_cpu_multicg2(gpu, orc, tmp_path):
    """`gpu_multicg --sweep` = the cpu_multicg2 the reference's Makefile:191 names but does not ship:
    invoked as eval_gflops.sh:61-66 does (--mtx --output --threads --timing_iters --quiet), it must
    write the CSV verification/gflops/gflop_analyze.py:13-52 consumes (pivot kernel x num_vectors of
    "gflops(iterations)") and the three SpmmKernel values must all reproduce CGSolveMultiple."""
    import pandas as pd
    _run("mtx_tool", "--grid3d=12", "--poisson", f"--out={tmp_path}/p12.mtx")
    csv_path = tmp_path / "p12_gflops.csv"
    _run("gpu_multicg", f"--mtx={tmp_path}/p12.mtx", f"--output={csv_path}", "--threads=17", "--timing_iters=1", "--quiet", "--sweep")
    lines = csv_path.read_text().splitlines()
    assert lines[0] == "matrix_name,kernel,num_vectors,min_ms,gflops,iterations"
    df = pd.read_csv(csv_path)
    assert len(df) == 21 and set(df["kernel"]) == {"SIMPLE", "MERGE", "NONZERO_SPLIT"}
    assert sorted(set(df["num_vectors"])) == [2, 4, 8, 16, 32, 64, 128]      # cpu_singlecg.cpp:159
    assert (df["matrix_name"] == "p12").all() and (df["gflops"] > 0).all() and (df["min_ms"] > 0).all()
    # the analysis script's own steps (gflop_analyze.py:27-52)
    df["formatted_value"] = df["gflops"].astype(str) + "(" + df["iterations"].astype(str) + ")"
    for kernel in df["kernel"].unique():
        piv = df[df["kernel"] == kernel].pivot_table(index="matrix_name", columns="num_vectors", values="formatted_value", aggfunc="first")
        piv = piv.reindex(sorted(piv.columns), axis=1)
        assert piv.shape == (1, 7) and not piv.isna().any().any()
    # iteration counts: each (kernel, k) against the oracle's CGSolveMultiple with that SpmmKernel
    ro, ci, va = orc.gen_grid3d(12, True, 6.0, -1.0)
    n = len(ro) - 1
    for kname, kid in (("SIMPLE", O.SIMPLE), ("MERGE", O.MERGE), ("NONZERO_SPLIT", O.NONZERO_SPLIT)):
        for k in (2, 16, 128):
            B = orc.rhs_rand(42, n * k).reshape(n, k)
            thr = orc.driver_threshold(B.ravel(), n, 1e-5)
            want = orc.cg_multi(ro, ci, va, B, k, 50000, thr, kid, 8)[0]
            got = int(df[(df["kernel"] == kname) & (df["num_vectors"] == k)]["iterations"].iloc[0])
            assert abs(got - want) <= max(1, round(0.02 * want)), (kname, k, got, want)


def test_gpu_singlecg_row_partitioned_over_two_gpus(gpu, orc, tmp_path):
    """--gpus=2: one forked worker per GPU, planner and IPC handles through shared memory
    (host/smle_multi.hpp); same totals as the single-GPU driver and as the oracle."""
    if gpu.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run("gpu_singlecg", "--grid3d=24", "--num_vectors=3", "--gpus=2", "--check", f"--output={tmp_path}/r2.csv")
    m = re.search(r"method=SINGLE_LOOP: [\d.]+ ms, (\d+) iters", out)
    assert m, out
    total = int(m.group(1))
    ro, ci, va = orc.gen_grid3d(24, True, 6.0, -1.0)
    n = len(ro) - 1
    b = orc.rhs_rand(42, n * 3)
    thr = orc.driver_threshold(b, n, 1e-5)
    want = sum(orc.cg_single(ro, ci, va, b[v * n:(v + 1) * n].copy(), 10000, thr)[0] for v in range(3))
    assert abs(total - want) <= max(1, round(0.02 * want))
    assert "row partition over 2 GPU(s): [0,6912) halo 576 [6912,13824) halo 576" in out     # merge-path cut, one 24x24 plane each
    res = float(re.search(r"true residual of vector 0: ([\d.e+-]+)", out).group(1))
    assert res < thr * 1.01
    assert (tmp_path / "r2.csv").read_text().splitlines()[1].startswith("grid3d_24,SINGLE_LOOP,3,")


def test_gpu_singlecg_partitioned_flag_on_one_gpu(gpu, tmp_path):
    """--partitioned with one GPU goes through the same forked-worker path (world = 1)"""
    out = _run("gpu_singlecg", "--grid3d=16", "--num_vectors=2", "--partitioned", "--check", f"--output={tmp_path}/r1.csv")
    assert "row partition over 1 GPU(s): [0,4096) halo 0" in out
    assert re.search(r"method=SINGLE_LOOP: [\d.]+ ms, (\d+) iters", out)


def test_gpu_multicg_columns_sharded_over_two_gpus(gpu, orc, tmp_path):
    """--gpus=2: the num_vectors columns split over two forked workers (A replicated); iteration count and error
    history of the whole block are merged from the shards (host/smle_multi.hpp)."""
    if gpu.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    (tmp_path / "data" / "error_data").mkdir(parents=True)
    out = _run("gpu_multicg", "--grid3d=20", "--num_vectors=6", "--timing_iters=1", "--gpus=2", "--check", cwd=tmp_path)
    iters = float(re.search(r"Iters:\s+([\d.]+)", out).group(1))
    ro, ci, va = orc.gen_grid3d(20, True, 6.0, -1.0)
    n = len(ro) - 1
    B = orc.rhs_rand(42, n * 6).reshape(n, 6)
    thr = orc.driver_threshold(B.ravel(), n, 1e-5)
    want, _, hist = orc.cg_multi(ro, ci, va, B, 6, 50000, thr, O.NONZERO_SPLIT, 8)
    assert abs(iters - want) <= max(1, round(0.02 * want))
    lines = (tmp_path / "data" / "error_data" / "grid3d_20_cg_errors.csv").read_text().splitlines()
    assert len(lines) - 1 == int(iters)
    np.testing.assert_allclose([float(l.split(",")[1]) for l in lines[1:6]], hist[:5], rtol=1e-5)
    assert float(re.search(r"true residual of column 0: ([\d.e+-]+)", out).group(1)) < thr * 1.01
