"""ctypes loader for the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  It exposes two back-ends with one numpy-facing interface:

* ``port()``  -- oracle/libsmle_oracle.so, the plain-C restatement (oracle/smle_oracle.c)
* ``ref()``   -- oracle/_ref/libsmle_ref_{cg,spmv}.so, the UNMODIFIED reference sources
                 compiled from /root/reference by oracle/Makefile (None when not built)

All matrices are CSR triples (row_offsets int32[m+1], column_indices int32[nnz],
values float64/float32[nnz]) exactly as the reference's CsrMatrix holds them
(sparse_matrix.h:648-653).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_I = C.c_int
_P = C.c_void_p

SIMPLE, MERGE, NONZERO_SPLIT = 0, 1, 2  # work_2025/types.hpp:11-16


def build(quiet: bool = True) -> None:
    """Compile the checker (oracle always; oracle/_ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", str(HERE), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"unsupported value type {dtype}")


def _ct(dtype):
    return C.c_double if np.dtype(dtype) == np.float64 else C.c_float


class _Backend:
    """Common numpy interface over the orc_* (port) or ref_* (reference) symbols."""

    kind = "port"

    def __init__(self, cg_lib, spmv_lib, prefix, takes_n):
        self._cg = cg_lib
        self._spmv = spmv_lib
        self._p = prefix
        self._takes_n = takes_n  # the ref wrappers also take num_cols

    # -- helpers ---------------------------------------------------------
    def _f(self, lib, name):
        fn = getattr(lib, f"{self._p}_{name}")
        fn.restype = None
        return fn

    def _dims(self, m, n, nnz):
        return (_I(m), _I(n), _I(nnz)) if self._takes_n else (_I(m), _I(nnz))

    def set_threads(self, t: int) -> None:
        if self._p == "orc":
            self._cg.orc_set_threads(_I(t))
        else:
            self._cg.ref_set_threads(_I(t))
            if self._spmv is not None:
                self._spmv.ref_spmv_set_threads(_I(t))

    # -- merge path --------------------------------------------------------
    def merge_path_search(self, diagonal, row_end, nnz):
        row_end = np.ascontiguousarray(row_end, dtype=np.int32)
        out = np.zeros(2, dtype=np.int32)
        self._f(self._cg, "merge_path_search")(_I(diagonal), _ptr(row_end), _I(len(row_end)),
                                                _I(nnz), _ptr(out))
        return int(out[0]), int(out[1])

    def merge_partition(self, row_offsets, num_parts, items_per_part=0):
        """(num_parts+1, 2) int32 share boundaries, as the reference threads compute them."""
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        m, nnz = len(ro) - 1, int(ro[-1])
        total = m + nnz
        share = items_per_part if items_per_part > 0 else (total + num_parts - 1) // num_parts
        out = np.zeros((num_parts + 1, 2), dtype=np.int32)
        row_end = ro[1:]
        for t in range(num_parts + 1):
            out[t] = self.merge_path_search(min(share * t, total), row_end, nnz)
        return out

    # -- kernels -----------------------------------------------------------
    def spmv_gold(self, ro, ci, va, x, y_in=None, alpha=1.0, beta=0.0, n=None):
        s, ct = _sfx(va.dtype), _ct(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        n = n if n is not None else len(x)
        y_in = np.zeros(m, dtype=va.dtype) if y_in is None else y_in
        y = np.empty(m, dtype=va.dtype)
        dims = (_I(m), _I(n), _I(nnz)) if self._takes_n else (_I(m),)
        self._f(self._cg, f"spmv_gold_{s}")(*dims, _ptr(ro), _ptr(ci), _ptr(va), _ptr(x),
                                             _ptr(y_in), _ptr(y), ct(alpha), ct(beta))
        return y

    def merge_csrmv(self, T, ro, ci, va, x, n=None):
        s = _sfx(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        n = n if n is not None else len(x)
        y = np.full(m, -1.0, dtype=va.dtype)
        if self._p == "orc":
            fn = self._cg.orc_merge_csrmv_f64 if s == "f64" else self._cg.orc_merge_csrmv_f32
            fn.restype = _I
            rc = fn(_I(T), _I(m), _I(nnz), _ptr(ro[1:]), _ptr(ci), _ptr(va), _ptr(x), _ptr(y))
            if rc != 0:
                raise ValueError("reference OmpMergeCsrmv supports at most 256 threads")
        else:
            if self._spmv is None:
                raise RuntimeError("oracle/_ref/libsmle_ref_spmv.so not built")
            if T > 256:
                raise ValueError("reference OmpMergeCsrmv supports at most 256 threads")
            self._f(self._spmv, f"merge_csrmv_{s}")(_I(T), _I(m), _I(n), _I(nnz), _ptr(ro),
                                                    _ptr(ci), _ptr(va), _ptr(x), _ptr(y))
        return y

    def _csrmm(self, name, T, ro, ci, va, X, k, n=None):
        s = _sfx(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        n = n if n is not None else X.size // k
        Y = np.zeros((m, k), dtype=va.dtype)
        Xc = np.ascontiguousarray(X, dtype=va.dtype)
        if self._p == "orc":
            if name == "row_split_csrmm":
                self._f(self._cg, f"{name}_{s}")(_I(T), _I(m), _ptr(ro), _ptr(ci), _ptr(va),
                                                 _ptr(Xc), _ptr(Y), _I(k))
            else:
                self._f(self._cg, f"{name}_{s}")(_I(T), _I(m), _I(nnz), _ptr(ro[1:]), _ptr(ci),
                                                 _ptr(va), _ptr(Xc), _ptr(Y), _I(k))
        else:
            self._f(self._cg, f"{name}_{s}")(_I(T), _I(m), _I(n), _I(nnz), _ptr(ro), _ptr(ci),
                                             _ptr(va), _ptr(Xc), _ptr(Y), _I(k))
        return Y

    def merge_csrmm(self, T, ro, ci, va, X, k, n=None):
        return self._csrmm("merge_csrmm", T, ro, ci, va, X, k, n)

    def nonzero_split_csrmm(self, T, ro, ci, va, X, k, n=None):
        return self._csrmm("nonzero_split_csrmm", T, ro, ci, va, X, k, n)

    def row_split_csrmm(self, T, ro, ci, va, X, k, n=None):
        return self._csrmm("row_split_csrmm", T, ro, ci, va, X, k, n)

    # -- solvers -----------------------------------------------------------
    def cg_single(self, ro, ci, va, b, max_iters, tol):
        """returns (iterations, x) -- CGSolveSingle (single_strategy.hpp:105-170)."""
        s, ct = _sfx(va.dtype), _ct(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        x = np.empty(m, dtype=va.dtype)
        b = np.ascontiguousarray(b, dtype=va.dtype)
        fn = getattr(self._cg, f"{self._p}_cg_single_{s}")
        fn.restype = _I
        dims = (_I(m), _I(m), _I(nnz)) if self._takes_n else (_I(m),)
        it = fn(*dims, _ptr(ro), _ptr(ci), _ptr(va), _ptr(b), _ptr(x), _I(max_iters), ct(tol))
        return int(it), x

    def cg_multi(self, ro, ci, va, B, k, max_iters, tol, kernel=MERGE, T=8):
        """returns (iterations, X, max_error_history) -- CGSolveMultiple
        (no_pretreatment.hpp:35-197); B, X row-major n x k."""
        s, ct = _sfx(va.dtype), _ct(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        B = np.ascontiguousarray(B, dtype=va.dtype)
        X = np.empty((m, k), dtype=va.dtype)
        hist = np.zeros(max(max_iters, 1), dtype=np.float64)
        hl = _I(0)
        fn = getattr(self._cg, f"{self._p}_cg_multi_{s}")
        fn.restype = _I
        if self._p == "orc":
            it = fn(_I(T), _I(m), _I(nnz), _ptr(ro), _ptr(ci), _ptr(va), _ptr(B), _ptr(X), _I(k),
                    _I(max_iters), ct(tol), _I(kernel), _ptr(hist), C.byref(hl))
        else:
            self.set_threads(T)
            it = fn(_I(m), _I(m), _I(nnz), _ptr(ro), _ptr(ci), _ptr(va), _ptr(B), _ptr(X), _I(k),
                    _I(max_iters), ct(tol), _I(kernel), _ptr(hist), C.byref(hl))
        return int(it), X, hist[: hl.value].copy()

    # -- SPAI (N3) -----------------------------------------------------------
    def spai_build(self, ro, ci, va):
        """values of the SPAI preconditioner M on A's pattern -- SparseApproximateInversion
        (work_2025/cg/sparse_approximate_inversion.hpp:41-321)."""
        m, nnz = len(ro) - 1, len(ci)
        mv = np.zeros(nnz, dtype=np.float64)
        fn = getattr(self._cg, f"{self._p}_spai_build_f64")
        fn.restype = _I
        dims = (_I(m), _I(m), _I(nnz)) if self._takes_n else (_I(m), _I(nnz))
        rc = fn(*dims, _ptr(ro), _ptr(ci), _ptr(va), _ptr(mv))
        if rc != 0:
            raise RuntimeError("SPAI construction failed")
        return mv

    def spai_solve_multi(self, ro, ci, va, mv, B, k, max_iters, tol, kernel=MERGE, T=8):
        """returns (iterations, X, max_error_history) -- SPAISolveMultiple
        (work_2025/main/sparse_approximate_inverse.hpp:31-230)."""
        m, nnz = len(ro) - 1, len(ci)
        B = np.ascontiguousarray(B, dtype=np.float64)
        X = np.empty((m, k), dtype=np.float64)
        hist = np.zeros(max(max_iters, 1), dtype=np.float64)
        hl = _I(0)
        fn = getattr(self._cg, f"{self._p}_spai_solve_multi_f64")
        fn.restype = _I
        if self._p == "orc":
            it = fn(_I(T), _I(m), _I(nnz), _ptr(ro), _ptr(ci), _ptr(va), _ptr(mv), _ptr(B), _ptr(X), _I(k),
                    _I(max_iters), C.c_double(tol), _I(kernel), _ptr(hist), C.byref(hl))
        else:
            self.set_threads(T)
            it = fn(_I(m), _I(m), _I(nnz), _ptr(ro), _ptr(ci), _ptr(va), _ptr(mv), _ptr(B), _ptr(X), _I(k),
                    _I(max_iters), C.c_double(tol), _I(kernel), _ptr(hist), C.byref(hl))
        return int(it), X, hist[: hl.value].copy()

    # -- generators ----------------------------------------------------------
    def _shape(self, name, *args):
        m, n, nnz = _I(0), _I(0), _I(0)
        self._f(self._cg, name)(*[_I(a) for a in args], C.byref(m), C.byref(n), C.byref(nnz))
        return m.value, n.value, nnz.value

    def _alloc(self, m, nnz, dtype):
        return (np.empty(m + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32),
                np.empty(nnz, dtype=dtype))

    def gen_grid2d(self, w, self_loop=True, diag=1.0, offd=1.0, dtype=np.float64):
        s, ct = _sfx(dtype), _ct(dtype)
        m, _, nnz = self._shape("gen_grid2d_shape", w, int(self_loop))
        ro, ci, va = self._alloc(m, nnz, dtype)
        self._f(self._cg, f"gen_grid2d_{s}")(_I(w), _I(int(self_loop)), ct(diag), ct(offd),
                                              _ptr(ro), _ptr(ci), _ptr(va))
        return ro, ci, va

    def gen_grid3d(self, w, self_loop=True, diag=1.0, offd=1.0, dtype=np.float64):
        s, ct = _sfx(dtype), _ct(dtype)
        m, _, nnz = self._shape("gen_grid3d_shape", w, int(self_loop))
        ro, ci, va = self._alloc(m, nnz, dtype)
        self._f(self._cg, f"gen_grid3d_{s}")(_I(w), _I(int(self_loop)), ct(diag), ct(offd),
                                              _ptr(ro), _ptr(ci), _ptr(va))
        return ro, ci, va

    def gen_grid3d_sorted(self, w, self_loop=True, diag=1.0, offd=1.0, dtype=np.float64):
        """the CSR of gen_grid3d built directly in sorted order (port only; for the 300^3 bench system)"""
        s, ct = _sfx(dtype), _ct(dtype)
        m, _, nnz = self._shape("gen_grid3d_shape", w, int(self_loop))
        ro, ci, va = self._alloc(m, nnz, dtype)
        self._f(self._cg, f"gen_grid3d_sorted_{s}")(_I(w), _I(int(self_loop)), ct(diag), ct(offd),
                                                     _ptr(ro), _ptr(ci), _ptr(va))
        return ro, ci, va

    def gen_wheel(self, spokes, value=1.0, dtype=np.float64):
        s, ct = _sfx(dtype), _ct(dtype)
        ro, ci, va = self._alloc(spokes + 1, 2 * spokes, dtype)
        self._f(self._cg, f"gen_wheel_{s}")(_I(spokes), ct(value), _ptr(ro), _ptr(ci), _ptr(va))
        return ro, ci, va

    def gen_dense(self, rows, cols, value=1.0, dtype=np.float64):
        s, ct = _sfx(dtype), _ct(dtype)
        ro, ci, va = self._alloc(rows, rows * cols, dtype)
        self._f(self._cg, f"gen_dense_{s}")(_I(rows), _I(cols), ct(value), _ptr(ro), _ptr(ci),
                                             _ptr(va))
        return ro, ci, va


class _Port(_Backend):
    kind = "port"

    def rhs_rand(self, seed, count, dtype=np.float64):
        """srand(seed); b[i] = rand()/RAND_MAX (cpu_singlecg.cpp:88-90)."""
        out = np.empty(count, dtype=dtype)
        self._f(self._cg, f"rhs_rand_{_sfx(dtype)}")(C.c_uint(seed), C.c_longlong(count), _ptr(out))
        return out

    def driver_threshold(self, b, n, tol):
        s, ct = _sfx(b.dtype), _ct(b.dtype)
        fn = getattr(self._cg, f"orc_driver_threshold_{s}")
        fn.restype = ct
        return float(fn(_ptr(b), _I(n), ct(tol)))


class _Ref(_Backend):
    kind = "reference"

    def test_merge_csrmv(self, T, ro, ci, va, x, timing_iters):
        """TestOmpMergeCsrmv (cpu_spmv.cpp:429-475): mean ms over timing_iters after as many warm-ups."""
        s = _sfx(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        y = np.empty(m, dtype=va.dtype)
        yref = np.zeros(m, dtype=va.dtype)
        fn = getattr(self._spmv, f"ref_test_merge_csrmv_{s}")
        fn.restype = C.c_float
        ms = fn(_I(T), _I(m), _I(len(x)), _I(nnz), _ptr(ro), _ptr(ci), _ptr(va), _ptr(x),
                _ptr(yref), _ptr(y), _I(timing_iters))
        return float(ms), y

    def test_cg_single(self, ro, ci, va, b_vectors, num_vectors, max_iters, tol, timing_iters=1):
        """TestCGSolveSingle (single_strategy.hpp:179-240): (min_ms, total_iters, x)."""
        s, ct = _sfx(va.dtype), _ct(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        b = np.ascontiguousarray(b_vectors, dtype=va.dtype)
        x = np.empty_like(b)
        ms, it = C.c_double(0), C.c_double(0)
        self._f(self._cg, f"test_cg_single_{s}")(_I(m), _I(m), _I(nnz), _ptr(ro), _ptr(ci),
                                                 _ptr(va), _ptr(b), _ptr(x), _I(max_iters),
                                                 ct(tol), _I(num_vectors), _I(timing_iters),
                                                 C.byref(ms), C.byref(it))
        return ms.value, it.value, x

    def read_mtx(self, path):
        """CooMatrix::InitMarket + CsrMatrix::Init -> (ro, ci, va, num_cols)."""
        m, n, nnz = _I(0), _I(0), _I(0)
        fn = self._cg.ref_read_mtx_f64
        fn.restype = _I
        fn(str(path).encode(), C.byref(m), C.byref(n), C.byref(nnz), None, None, None)
        ro, ci, va = self._alloc(m.value, nnz.value, np.float64)
        fn(str(path).encode(), C.byref(m), C.byref(n), C.byref(nnz), _ptr(ro), _ptr(ci), _ptr(va))
        return ro, ci, va, n.value

    def test_cg_multi(self, ro, ci, va, B, k, max_iters, tol, kernel=MERGE, timing_iters=1):
        """TestCGMultipleRHS (no_pretreatment.hpp:205-256): (min_ms, iters, X)."""
        s, ct = _sfx(va.dtype), _ct(va.dtype)
        m, nnz = len(ro) - 1, len(ci)
        B = np.ascontiguousarray(B, dtype=va.dtype)
        X = np.empty_like(B)
        ms, it = C.c_double(0), C.c_double(0)
        self._f(self._cg, f"test_cg_multi_{s}")(_I(m), _I(m), _I(nnz), _ptr(ro), _ptr(ci),
                                                _ptr(va), _ptr(B), _ptr(X), _I(max_iters),
                                                ct(tol), _I(k), _I(timing_iters), _I(kernel),
                                                C.byref(ms), C.byref(it))
        return ms.value, it.value, X


_port = None
_ref = None
_ref_tried = False


def port() -> _Port:
    """The plain-C restatement (always available; built on demand)."""
    global _port
    if _port is None:
        so = HERE / "libsmle_oracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(HERE), "oracle"], check=True,
                           stdout=subprocess.DEVNULL)
        lib = C.CDLL(str(so))
        _port = _Port(lib, None, "orc", takes_n=False)
    return _port


def ref():
    """The compiled reference (None when oracle/_ref was never built)."""
    global _ref, _ref_tried
    if not _ref_tried:
        _ref_tried = True
        cg = HERE / "_ref" / "libsmle_ref_cg.so"
        sp = HERE / "_ref" / "libsmle_ref_spmv.so"
        if not cg.exists() and Path("/root/reference/sparse_matrix.h").exists():
            subprocess.run(["make", "-C", str(HERE), "ref"], check=False,
                           stdout=subprocess.DEVNULL)
        if cg.exists():
            _ref = _Ref(C.CDLL(str(cg)), C.CDLL(str(sp)) if sp.exists() else None, "ref",
                        takes_n=True)
    return _ref


def host_cores() -> int:
    return len(os.sched_getaffinity(0))
