"""ctypes binding of include/smle_b200.h.  See the package docstring."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

PKG_ROOT = Path(__file__).resolve().parents[2]           # sparse-matrix-linear-equations_b200/
REPO_ROOT = PKG_ROOT.parent
HEADER = REPO_ROOT / "include" / "smle_b200.h"

SIMPLE, MERGE, NONZERO_SPLIT = 0, 1, 2                    # work_2025/types.hpp:11-16

_I, _P, _D, _F = C.c_int, C.c_void_p, C.c_double, C.c_float


class SmleError(RuntimeError):
    """Raised for every non-zero status of the C ABI (and when the library is missing)."""


def lib_path() -> Path:
    return PKG_ROOT / "libsmle_b200.so"


def _declared_symbols():
    """Every function the header declares (used by the symbol-export test)."""
    if not HEADER.exists():
        return []
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(smle_[a-z0-9_]+)\s*\(", text)))


DECLARED_SYMBOLS = _declared_symbols()

_lib = None


def lib():
    """Load libsmle_b200.so; fail loudly when it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise SmleError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; "
                            f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(str(p))
        L.smle_last_error.restype = C.c_char_p
        L.smle_get_stream.restype = _P
        L.smle_launch_count.restype = C.c_longlong
        L.smle_driver_threshold_f64.restype = _D
        L.smle_dist_plan_request_size.restype = C.c_longlong
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise SmleError(f"smle status {rc}: {lib().smle_last_error().decode()}")


def _publish(dev: int) -> None:
    """Device outputs were written on the library's stream: unless the caller shares that stream
    (set_stream), wait for them so that torch ops on other streams see finished data."""
    if dev:
        import torch
        if torch.cuda.current_stream().cuda_stream != get_stream():
            sync()


def device_count() -> int:
    return int(lib().smle_device_count())


def init(device: int = 0) -> None:
    _check(lib().smle_init(_I(device)))


def set_stream(cuda_stream: int | None) -> None:
    """Launch on the caller's cudaStream_t (e.g. torch.cuda.Stream().cuda_stream)."""
    _check(lib().smle_set_stream(_P(cuda_stream or 0)))


def get_stream() -> int:
    return int(lib().smle_get_stream() or 0)


def sync() -> None:
    _check(lib().smle_sync())


def launch_count() -> int:
    return int(lib().smle_launch_count())


def sm_count() -> int:
    return int(lib().smle_sm_count())


# ---------------------------------------------------------------------------------------------
# argument marshalling: numpy arrays are host memory, torch CUDA tensors are device memory
# ---------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _arg(x, dtype, writable=False):
    """-> (void*, is_device, keepalive)"""
    if _is_torch(x):
        import torch
        want = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
        if x.dtype != want or not x.is_contiguous():
            raise SmleError(f"tensor must be contiguous {want}")
        if x.is_cuda and not writable:
            # inputs produced on torch's current stream must be complete before the library's
            # stream reads them (no-op when the caller shares its stream through set_stream)
            cur = torch.cuda.current_stream(x.device)
            if cur.cuda_stream != get_stream():
                cur.synchronize()
        return _P(x.data_ptr()), (1 if x.is_cuda else 0), x
    if isinstance(x, np.ndarray) and x.dtype.kind == "f" and x.dtype != np.dtype(dtype):
        raise SmleError(f"array of {x.dtype} passed where the handle computes in {np.dtype(dtype)}")
    a = np.ascontiguousarray(x, dtype=dtype)
    if writable and a is not x:
        raise SmleError("output array must be a contiguous numpy array of the value type")
    return a.ctypes.data_as(_P), 0, a


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise SmleError(f"unsupported value type {dtype}")


def _empty_like_arg(x, shape, dtype):
    if _is_torch(x):
        import torch
        return torch.empty(shape, dtype=x.dtype, device=x.device)
    return np.empty(shape, dtype=dtype)


# ---------------------------------------------------------------------------------------------
# merge path
# ---------------------------------------------------------------------------------------------
def merge_path_partition(row_offsets, num_parts: int, items_per_part: int = 0) -> np.ndarray:
    """(num_parts+1, 2) int32 (row, nnz) coordinates computed by the GPU search kernel;
    must equal MergePathSearch (merge_based.hpp:22-44) on the reference's share diagonals."""
    ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
    m, nnz = len(ro) - 1, int(ro[-1])
    row_end = np.ascontiguousarray(ro[1:])
    out = np.zeros((num_parts + 1, 2), dtype=np.int32)
    _check(lib().smle_merge_path_partition(row_end.ctypes.data_as(_P), _I(m), _I(nnz), _I(num_parts),
                                           _I(items_per_part), out.ctypes.data_as(_P)))
    return out


# ---------------------------------------------------------------------------------------------
# CsrMatrix handle
# ---------------------------------------------------------------------------------------------
class CsrMatrix:
    """Device-resident CsrMatrix<ValueT,int> (sparse_matrix.h:633-653)."""

    def __init__(self, row_offsets, column_indices, values, num_cols: int | None = None):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        ci = np.ascontiguousarray(column_indices, dtype=np.int32)
        va = np.ascontiguousarray(values)
        self.dtype = va.dtype
        s = _sfx(self.dtype)
        self.num_rows = len(ro) - 1
        self.num_nonzeros = len(ci)
        self.num_cols = int(num_cols) if num_cols is not None else self.num_rows
        h = _P()
        _check(getattr(lib(), f"smle_csr_create_{s}")(C.byref(h), _I(self.num_rows), _I(self.num_cols),
                                                      _I(self.num_nonzeros), ro.ctypes.data_as(_P),
                                                      ci.ctypes.data_as(_P), va.ctypes.data_as(_P)))
        self._h = h
        self._s = s

    def close(self):
        if getattr(self, "_h", None):
            lib().smle_csr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- partition actually used by the kernels ---------------------------------------------
    def tile_coords(self, k: int = 1):
        nt, ipt = _I(0), _I(0)
        _check(lib().smle_csr_tile_coords(self._h, _I(k), C.byref(nt), C.byref(ipt), None, _I(0)))
        out = np.zeros((nt.value + 1, 2), dtype=np.int32)
        _check(lib().smle_csr_tile_coords(self._h, _I(k), C.byref(nt), C.byref(ipt),
                                          out.ctypes.data_as(_P), _I(out.size)))
        return out, ipt.value

    # -- SpMV / SpMM ------------------------------------------------------------------------------
    def spmv(self, x, out=None):
        """y = A x  (OmpMergeCsrmv, cpu_spmv.cpp:360-421)."""
        px, dev, kx = _arg(x, self.dtype)
        y = out if out is not None else _empty_like_arg(x, (self.num_rows,), self.dtype)
        py, dev_y, ky = _arg(y, self.dtype, writable=True)
        if dev != dev_y:
            raise SmleError("x and y must both be host or both be device memory")
        _check(getattr(lib(), f"smle_spmv_{self._s}")(self._h, px, py, _I(dev)))
        _publish(dev)
        return y

    def spmm(self, X, out=None):
        """Y = A X, row-major n x k -> m x k  (OmpMergeCsrmm, merge_based.hpp:49-153)."""
        k = int(X.shape[1])
        pX, dev, kX = _arg(X, self.dtype)
        Y = out if out is not None else _empty_like_arg(X, (self.num_rows, k), self.dtype)
        pY, dev_y, kY = _arg(Y, self.dtype, writable=True)
        if dev != dev_y:
            raise SmleError("X and Y must both be host or both be device memory")
        _check(getattr(lib(), f"smle_spmm_{self._s}")(self._h, pX, pY, _I(k), _I(dev)))
        _publish(dev)
        return Y

    # -- CG -------------------------------------------------------------------------------------
    def cg_solve_single(self, b, max_iters: int, tolerance: float, out=None):
        """-> (iterations, x, final_rel_res)   CGSolveSingle (single_strategy.hpp:105-170)."""
        pb, dev, kb = _arg(b, self.dtype)
        x = out if out is not None else _empty_like_arg(b, (self.num_rows,), self.dtype)
        px, dev_x, kx = _arg(x, self.dtype, writable=True)
        if dev != dev_x:
            raise SmleError("b and x must both be host or both be device memory")
        it, rel = _I(0), _D(0)
        tol = _D(tolerance) if self._s == "f64" else _F(tolerance)
        _check(getattr(lib(), f"smle_cg_single_{self._s}")(self._h, pb, px, _I(max_iters), tol, _I(dev),
                                                           C.byref(it), C.byref(rel)))
        return it.value, x, rel.value

    def cg_solve_single_batch(self, b_vectors, max_iters: int, tolerance: float, out=None):
        """-> (iterations per vector, X)   the solve loop of TestCGSolveSingle (single_strategy.hpp:199-226):
        b_vectors is an (L, n) HOST array (numpy or pinned torch), vector v = row v; the host copies of
        neighbouring systems overlap with the solves."""
        pb, dev, _ = _arg(b_vectors, np.float64)
        L = int(b_vectors.shape[0])
        x = out if out is not None else _empty_like_arg(b_vectors, (L, self.num_rows), np.float64)
        px, dev_x, _ = _arg(x, np.float64, writable=True)
        if dev or dev_x:
            raise SmleError("cg_solve_single_batch takes host memory (device callers use cg_solve_single)")
        each = (C.c_int * max(L, 1))()
        total = C.c_longlong(0)
        _check(lib().smle_cg_single_batch_f64(self._h, pb, px, _I(L), _I(max_iters), _D(tolerance), each, C.byref(total)))
        return [each[i] for i in range(L)], x

    def cg_solve_multiple(self, B, max_iters: int, tolerance: float, kernel_type: int = MERGE,
                          out=None, want_history: bool = True):
        """-> (iterations, X, max_errors, final_rel_res)
        CGSolveMultiple (no_pretreatment.hpp:35-197); B, X row-major n x k."""
        k = int(B.shape[1])
        pB, dev, kB = _arg(B, self.dtype)
        X = out if out is not None else _empty_like_arg(B, (self.num_rows, k), self.dtype)
        pX, dev_x, kX = _arg(X, self.dtype, writable=True)
        if dev != dev_x:
            raise SmleError("B and X must both be host or both be device memory")
        cap = max(int(max_iters), 1) if want_history else 0
        hist = np.zeros(cap, dtype=np.float64) if want_history else None
        it, hl, rel = _I(0), _I(0), _D(0)
        tol = _D(tolerance) if self._s == "f64" else _F(tolerance)
        _check(getattr(lib(), f"smle_cg_multi_{self._s}")(self._h, pB, pX, _I(k), _I(max_iters), tol,
                                                          _I(kernel_type), _I(dev), C.byref(it),
                                                          hist.ctypes.data_as(_P) if want_history else None, _I(cap),
                                                          C.byref(hl), C.byref(rel)))
        return it.value, X, (hist[: hl.value].copy() if want_history else None), rel.value

    def pcg_spai_solve_multiple(self, M: "CsrMatrix", B, max_iters: int, tolerance: float, kernel_type: int = MERGE,
                                out=None, want_history: bool = True):
        """-> (iterations, X, max_errors, final_rel_res)
        SPAISolveMultiple (work_2025/main/sparse_approximate_inverse.hpp:31-230); M = handle of the SPAI matrix."""
        k = int(B.shape[1])
        pB, dev, kB = _arg(B, np.float64)
        X = out if out is not None else _empty_like_arg(B, (self.num_rows, k), np.float64)
        pX, dev_x, kX = _arg(X, np.float64, writable=True)
        if dev != dev_x:
            raise SmleError("B and X must both be host or both be device memory")
        cap = max(int(max_iters), 1) if want_history else 0
        hist = np.zeros(cap, dtype=np.float64) if want_history else None
        it, hl, rel = _I(0), _I(0), _D(0)
        _check(lib().smle_pcg_spai_multi_f64(self._h, M._h, pB, pX, _I(k), _I(max_iters), _D(tolerance), _I(kernel_type),
                                             _I(dev), C.byref(it), hist.ctypes.data_as(_P) if want_history else None,
                                             _I(cap), C.byref(hl), C.byref(rel)))
        return it.value, X, (hist[: hl.value].copy() if want_history else None), rel.value

    def cg_run_fixed(self, B, X, iters: int):
        """exactly `iters` iterations on device blocks (measurement helper)."""
        pB, dev, kB = _arg(B, np.float64)
        pX, dev_x, kX = _arg(X, np.float64, writable=True)
        if not (dev and dev_x):
            raise SmleError("cg_run_fixed needs device tensors")
        k = int(B.shape[1]) if B.dim() == 2 else 1
        _check(lib().smle_cg_run_fixed_f64(self._h, pB, pX, _I(k), _I(iters)))

    def cg_profile(self, B, X, iters: int):
        """mean ms of the three kernels of a CG iteration (CUDA events, no graph)."""
        pB, dev, kB = _arg(B, np.float64)
        pX, dev_x, kX = _arg(X, np.float64, writable=True)
        if not (dev and dev_x):
            raise SmleError("cg_profile needs device tensors")
        k = int(B.shape[1]) if B.dim() == 2 else 1
        out = (C.c_float * 3)()
        _check(lib().smle_cg_profile_f64(self._h, pB, pX, _I(k), _I(iters), out))
        return [float(v) for v in out]


# ---------------------------------------------------------------------------------------------
# generators (host side; CSR identical to reference generator + CsrMatrix::Init)
# ---------------------------------------------------------------------------------------------
def _gen(name, shape_args, gen_args, dtype):
    s = _sfx(dtype)
    m, n, nnz = _I(0), _I(0), _I(0)
    _check(getattr(lib(), f"smle_gen_{name}_shape")(*shape_args, C.byref(m), C.byref(n), C.byref(nnz)))
    ro = np.empty(m.value + 1, dtype=np.int32)
    ci = np.empty(nnz.value, dtype=np.int32)
    va = np.empty(nnz.value, dtype=dtype)
    _check(getattr(lib(), f"smle_gen_{name}_{s}")(*gen_args, ro.ctypes.data_as(_P), ci.ctypes.data_as(_P),
                                                  va.ctypes.data_as(_P)))
    return ro, ci, va


def _ct(dtype):
    return _D if np.dtype(dtype) == np.float64 else _F


def gen_grid2d(width, self_loop=True, diag=1.0, offd=1.0, dtype=np.float64):
    """InitGrid2d (sparse_matrix.h:458-527) -> CSR; diag/offd = 4/-1 gives the 2-D Poisson matrix."""
    ct = _ct(dtype)
    return _gen("grid2d", (_I(width), _I(int(self_loop))), (_I(width), _I(int(self_loop)), ct(diag), ct(offd)), dtype)


def gen_grid3d(width, self_loop=True, diag=1.0, offd=1.0, dtype=np.float64):
    """InitGrid3d (sparse_matrix.h:533-623) -> CSR; diag/offd = 6/-1 gives the 3-D Poisson matrix."""
    ct = _ct(dtype)
    return _gen("grid3d", (_I(width), _I(int(self_loop))), (_I(width), _I(int(self_loop)), ct(diag), ct(offd)), dtype)


def gen_grid3d_row_offsets(width, self_loop=True) -> np.ndarray:
    """row offsets of InitGrid3d(width, self_loop) alone (m+1 ints): what a rank of the row-partitioned
    solve needs to find its rows without building the matrix."""
    m, n, nnz = _I(0), _I(0), _I(0)
    _check(lib().smle_gen_grid3d_shape(_I(width), _I(int(self_loop)), C.byref(m), C.byref(n), C.byref(nnz)))
    ro = np.empty(m.value + 1, dtype=np.int32)
    _check(lib().smle_gen_grid3d_row_offsets(_I(width), _I(int(self_loop)), ro.ctypes.data_as(_P)))
    return ro


def gen_grid3d_rows(width, r0, r1, nnz_rows, self_loop=True, diag=1.0, offd=1.0):
    """rows [r0, r1) of the same matrix: (local row offsets from 0, GLOBAL column indices, values);
    nnz_rows = row_offsets[r1] - row_offsets[r0]."""
    lro = np.empty(r1 - r0 + 1, dtype=np.int32)
    ci = np.empty(nnz_rows, dtype=np.int32)
    va = np.empty(nnz_rows, dtype=np.float64)
    _check(lib().smle_gen_grid3d_rows_f64(_I(width), _I(int(self_loop)), _D(diag), _D(offd), _I(r0), _I(r1),
                                          lro.ctypes.data_as(_P), ci.ctypes.data_as(_P), va.ctypes.data_as(_P)))
    assert int(lro[-1]) == nnz_rows
    return lro, ci, va


def gen_wheel(spokes, value=1.0, dtype=np.float64):
    """InitWheel (sparse_matrix.h:417-450) -> CSR."""
    return _gen("wheel", (_I(spokes),), (_I(spokes), _ct(dtype)(value)), dtype)


def gen_dense(rows, cols, value=1.0, dtype=np.float64):
    """InitDense (sparse_matrix.h:385-412) -> CSR."""
    return _gen("dense", (_I(rows), _I(cols)), (_I(rows), _I(cols), _ct(dtype)(value)), dtype)


def gen_rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=42, unit_values=False, dtype=np.float64):
    """R-MAT power-law matrix standing in for the SuiteSparse set (no network here)."""
    return _gen("rmat", (_I(scale), _I(edge_factor)),
                (_I(scale), _I(edge_factor), _D(a), _D(b), _D(c), C.c_ulonglong(seed), _I(int(unit_values))), dtype)


def spai_build(row_offsets, column_indices, values) -> np.ndarray:
    """values of the SPAI preconditioner on A's pattern (SparseApproximateInversion,
    work_2025/cg/sparse_approximate_inversion.hpp:41-321); host code."""
    ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
    ci = np.ascontiguousarray(column_indices, dtype=np.int32)
    va = np.ascontiguousarray(values, dtype=np.float64)
    out = np.zeros(len(ci), dtype=np.float64)
    _check(lib().smle_spai_build_f64(_I(len(ro) - 1), _I(len(ci)), ro.ctypes.data_as(_P), ci.ctypes.data_as(_P),
                                     va.ctypes.data_as(_P), out.ctypes.data_as(_P)))
    return out


def gen_rhs_rand(seed: int, count: int) -> np.ndarray:
    """srand(seed); b[i] = rand()/RAND_MAX  (cpu_singlecg.cpp:88-90)."""
    out = np.empty(count, dtype=np.float64)
    _check(lib().smle_gen_rhs_rand_f64(C.c_uint(seed), C.c_longlong(count), out.ctypes.data_as(_P)))
    return out


def gen_rhs_rand_range(seed: int, first: int, count: int) -> np.ndarray:
    """entries [first, first+count) of the srand(seed) stream (the values before are drawn and dropped)."""
    out = np.empty(count, dtype=np.float64)
    _check(lib().smle_gen_rhs_rand_range_f64(C.c_uint(seed), C.c_longlong(first), C.c_longlong(count),
                                             out.ctypes.data_as(_P)))
    return out


def dist_bounds(row_offsets, world: int) -> np.ndarray:
    """first row of every part of the row partition (world+1 ints): MergePathSearch on the share
    diagonals g*ceil((m+nnz)/world), run on the GPU (smle_dist_bounds)."""
    ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
    out = np.zeros(world + 1, dtype=np.int32)
    _check(lib().smle_dist_bounds(ro.ctypes.data_as(_P), _I(len(ro) - 1), _I(world), out.ctypes.data_as(_P)))
    return out


def host_register(arr) -> None:
    """page-lock a numpy array for overlapped copies (smle_host_register)."""
    rc = lib().smle_host_register(_P(arr.ctypes.data), C.c_ulonglong(arr.nbytes))
    if rc < 0:
        _check(rc)


def host_unregister(arr) -> None:
    _check(lib().smle_host_unregister(_P(arr.ctypes.data)))


def driver_threshold(b: np.ndarray, n: int, tol: float) -> float:
    """||b[0:n]||_2 * tol -- the tolerance the reference drivers pass (cpu_singlecg.cpp:23-34,92)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().smle_driver_threshold_f64(b.ctypes.data_as(_P), _I(n), _D(tol)))
