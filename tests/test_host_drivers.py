"""CPU: host side of the drivers -- Matrix Market reader/writer against the reference's
CooMatrix::InitMarket (sparse_matrix.h:211-380), the drivers exist and fail loudly without a GPU."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
BIN = ROOT / "sparse-matrix-linear-equations_b200" / "bin"


def _need_bin(name):
    p = BIN / name
    if not p.exists():
        pytest.skip(f"{p} not built")
    return p


SYMMETRIC = """%%MatrixMarket matrix coordinate real symmetric
% a comment
4 4 6
1 1 4.0
2 1 -1.0
2 2 4.0
3 2 -1.5
4 4 2.0
4 1 0.25
"""

PATTERN = """%%MatrixMarket matrix coordinate pattern general
3 5 4
1 5
3 1
2 2
1 1
"""

SKEW = """%%MatrixMarket matrix coordinate real skew-symmetric
3 3 2
2 1 1.5
3 1 -2.0
"""


@pytest.mark.parametrize("text", [SYMMETRIC, PATTERN, SKEW])
def test_matrix_market_reader_matches_reference(ref, tmp_path, text):
    tool = _need_bin("mtx_tool")
    src, out = tmp_path / "in.mtx", tmp_path / "out.mtx"
    src.write_text(text)
    subprocess.run([str(tool), f"--mtx={src}", f"--out={out}"], check=True, capture_output=True)
    a = ref.read_mtx(src)      # reference reader on the original file
    b = ref.read_mtx(out)      # reference reader on what OUR reader+writer produced
    assert a[3] == b[3]
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)


def test_generated_poisson_roundtrips_through_mtx(ref, tmp_path):
    tool = _need_bin("mtx_tool")
    out = tmp_path / "p5.mtx"
    subprocess.run([str(tool), "--grid3d=5", "--poisson", f"--out={out}"], check=True, capture_output=True)
    ro, ci, va, n = ref.read_mtx(out)
    want = ref.gen_grid3d(5, True, 6.0, -1.0)
    assert n == 125 and all(np.array_equal(x, y) for x, y in zip((ro, ci, va), want))


@pytest.mark.parametrize("drv", ["gpu_spmv", "gpu_spmm", "gpu_singlecg", "gpu_multicg"])
def test_drivers_fail_loudly_without_gpu(S, drv):
    exe = _need_bin(drv)
    if S.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([str(exe), "--grid2d=8", "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
