"""smle_b200 -- Python face of libsmle_b200.so (include/smle_b200.h).

The product is the C-ABI CUDA library; this package is a thin ctypes binding that mirrors the
reference's function names for the hot path so tests and bench read like
the reference's own call sites:

    reference (C++ template)                      here
    ------------------------------------------    -----------------------------------------
    CsrMatrix<V,int>(coo)                          CsrMatrix(row_offsets, column_indices, values)
    MergePathSearch on the thread diagonals        merge_path_partition(row_offsets, parts)
    OmpMergeCsrmv(T, a, ..., x, y)                 a.spmv(x)
    OmpMergeCsrmm(T, a, ..., X, Y, k)              a.spmm(X)
    CGSolveSingle(a, b, x, max_iters, tol)         a.cg_solve_single(b, max_iters, tol)
    CGSolveMultiple(a, B, X, k, max_iters, tol..)  a.cg_solve_multiple(B, max_iters, tol)
    SparseApproximateInversion(a, m)               spai_build(row_offsets, column_indices, values)
    SPAISolveMultiple(a, m, B, X, k, ...)          a.pcg_spai_solve_multiple(M, B, max_iters, tol)

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible,
every compute call raises SmleError.
"""
from .capi import (SIMPLE, MERGE, NONZERO_SPLIT, CsrMatrix, SmleError, device_count, driver_threshold,
                   gen_dense, gen_grid2d, gen_grid3d, gen_grid3d_row_offsets, gen_grid3d_rows, gen_rhs_rand,
                   gen_rhs_rand_range, gen_rmat, gen_wheel, get_stream, dist_bounds, host_register, host_unregister,
                   init, launch_count, lib, lib_path, merge_path_partition, set_stream, sm_count, spai_build,
                   sync, DECLARED_SYMBOLS)

__all__ = [
    "SIMPLE", "MERGE", "NONZERO_SPLIT", "CsrMatrix", "SmleError", "device_count", "driver_threshold",
    "gen_dense", "gen_grid2d", "gen_grid3d", "gen_grid3d_row_offsets", "gen_grid3d_rows", "gen_rhs_rand",
    "gen_rhs_rand_range", "gen_rmat", "gen_wheel", "get_stream", "dist_bounds", "host_register", "host_unregister",
    "init", "launch_count", "lib", "lib_path", "merge_path_partition", "set_stream", "sm_count", "spai_build",
    "sync", "DECLARED_SYMBOLS",
]
