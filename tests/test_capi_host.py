"""CPU: the C-ABI library loads, exports every symbol include/smle_b200.h declares, its host-side
generators equal the oracle's restatement of the reference generators, and compute entry points
fail loudly (no CPU fallback) when no CUDA device is visible.  No compute calls here."""
import ctypes as C

import numpy as np
import pytest


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_library_exports_every_declared_symbol(S):
    assert S.lib_path().exists(), "libsmle_b200.so not built: run __graft_entry__.build()"
    assert len(S.DECLARED_SYMBOLS) >= 35
    missing = [s for s in S.DECLARED_SYMBOLS if not hasattr(S.lib(), s)]
    assert not missing, missing
    assert S.lib().smle_version() >= 100


def test_library_has_sm100a_code(S):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", str(S.lib_path())],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_generators_match_oracle(S, orc, dtype):
    for w in (1, 2, 3, 6, 13):
        for loop in (False, True):
            assert _same(S.gen_grid3d(w, loop, 6.0, -1.0, dtype), orc.gen_grid3d(w, loop, 6.0, -1.0, dtype))
            assert _same(S.gen_grid2d(w, loop, 4.0, -1.0, dtype), orc.gen_grid2d(w, loop, 4.0, -1.0, dtype))
    for s in (1, 2, 9, 1000):
        assert _same(S.gen_wheel(s, 1.0, dtype), orc.gen_wheel(s, 1.0, dtype))
    assert _same(S.gen_dense(7, 4, 3.0, dtype), orc.gen_dense(7, 4, 3.0, dtype))
    assert _same(S.gen_dense(0, 4, 3.0, dtype), orc.gen_dense(0, 4, 3.0, dtype))


def test_grid_shapes_match_baseline_configs(S):
    # BASELINE.md section 3 sizes
    ro, ci, _ = S.gen_grid2d(1000, True)
    assert (len(ro) - 1, len(ci)) == (1000000, 4996000)
    ro, ci, _ = S.gen_grid2d(1000, False)
    assert len(ci) == 3996000
    ro, ci, va = S.gen_grid3d(150, True, 6.0, -1.0)
    assert (len(ro) - 1, len(ci)) == (3375000, 23490000)
    assert va[ci == np.repeat(np.arange(len(ro) - 1), np.diff(ro))].min() == 6.0


def test_rhs_and_threshold_match_oracle(S, orc):
    b = S.gen_rhs_rand(42, 5000)
    assert np.array_equal(b, orc.rhs_rand(42, 5000))
    assert abs(S.driver_threshold(b, 1250, 1e-5) - orc.driver_threshold(b, 1250, 1e-5)) < 1e-18


def test_rmat_is_valid_csr_and_deterministic(S):
    ro, ci, va = S.gen_rmat(10, 8, seed=7)
    assert len(ro) == 1025 and ro[0] == 0 and ro[-1] == 8192 == len(ci)
    assert np.all(np.diff(ro) >= 0) and ci.min() >= 0 and ci.max() < 1024
    for r in range(1024):
        assert np.all(np.diff(ci[ro[r]:ro[r + 1]]) >= 0)       # sorted, duplicates kept
    assert 0.0 < va.min() and va.max() <= 1.0
    ro2, ci2, va2 = S.gen_rmat(10, 8, seed=7)
    assert _same((ro, ci, va), (ro2, ci2, va2))
    assert np.diff(ro).max() > 8 * 8                            # power-law skew
    assert np.all(S.gen_rmat(8, 4, unit_values=True)[2] == 1.0)


def test_bad_arguments_are_reported_not_fatal(S):
    L = S.lib()
    m = C.c_int()
    assert L.smle_gen_grid3d_shape(C.c_int(0), C.c_int(1), C.byref(m), C.byref(m), C.byref(m)) == -1
    assert L.smle_gen_grid3d_shape(C.c_int(2000), C.c_int(1), C.byref(m), C.byref(m), C.byref(m)) == -1
    assert L.smle_csr_dims(None, None, None, None, None) == -1
    assert b"null handle" in L.smle_last_error()


def test_no_cpu_fallback_without_device(S):
    if S.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    ro, ci, va = S.gen_wheel(4)
    with pytest.raises(S.SmleError, match="no CUDA device"):
        S.CsrMatrix(ro, ci, va)
    with pytest.raises(S.SmleError, match="no CUDA device"):
        S.merge_path_partition(ro, 2)
