// smle_adapters.hpp -- the reference's C++ function signatures on top of the C ABI.
//
// Drop-in layer for the hot path of YuyaW-0118/Sparse-Matrix-Linear-Equations.  Every function
// below keeps the argument list of the reference template it replaces, so a reference driver
// switches to the GPU by changing the callee's name:
//
//   GpuMergeCsrmv        <- OmpMergeCsrmv          cpu_spmv.cpp:360-421
//   GpuMergeCsrmm        <- OmpMergeCsrmm          work_2025/spmm/merge_based.hpp:49-153
//   GpuNonzeroSplitCsrmm <- OmpNonzeroSplitCsrmm   work_2025/spmm/nonzero_splitting.hpp:52-150
//   GpuCsrSpmmT          <- OmpCsrSpmmT            work_2025/spmm/row_splitting.hpp:18-54
//   GpuMergePathSearch   <- MergePathSearch        work_2025/spmm/merge_based.hpp:22-44
//   GpuCGSolveSingle     <- CGSolveSingle          work_2025/main/single_strategy.hpp:105-170
//   GpuCGSolveMultiple   <- CGSolveMultiple        work_2025/main/no_pretreatment.hpp:35-197
//   TestGpuCGSolveSingle <- TestCGSolveSingle      work_2025/main/single_strategy.hpp:179-240
//   TestGpuCGMultipleRHS <- TestCGMultipleRHS      work_2025/main/no_pretreatment.hpp:205-256
//   GpuSparseApproximateInversion <- SparseApproximateInversion  work_2025/cg/sparse_approximate_inversion.hpp:41-321
//   GpuSPAISolveMultiple <- SPAISolveMultiple     work_2025/main/sparse_approximate_inverse.hpp:31-230
//   TestGpuCGMultipleSPAI <- TestCGMultipleSPAI   work_2025/main/sparse_approximate_inverse.hpp:232-287
//
// The matrix argument is duck-typed: anything with the fields of the reference's
// CsrMatrix<ValueT,int> (sparse_matrix.h:648-653) works, including the reference type itself, so
// this header does not include (or copy) any reference header.  Ownership and error behaviour
// follow the reference: the caller owns all host memory, outputs are fully overwritten, fatal
// errors print to stderr and exit(1) (sparse_matrix.h:186-190) -- the library underneath never
// exits, the adapter does it to keep the reference's contract.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/smle_b200.h"

namespace smle_adapters {

inline void die(const char *what)
{
    fprintf(stderr, "%s: %s\n", what, smle_last_error());
    exit(1);
}

// Device handles are cached per (row_offsets, column_indices, values) so that repeated calls on
// the same CsrMatrix (every CG iteration count sweep, every timing iteration) upload A once.
struct HandleCache {
    std::map<std::tuple<const void *, const void *, const void *, int, int>, smle_csr_t> map;
    ~HandleCache() { for (auto &kv : map) smle_csr_destroy(kv.second); }
    static HandleCache &get() { static HandleCache c; return c; }
};

template <typename CsrT>
smle_csr_t handle_of(CsrT &a)
{
    using V = typename std::remove_pointer<decltype(a.values)>::type;
    auto key = std::make_tuple((const void *)a.row_offsets, (const void *)a.column_indices,
                               (const void *)a.values, (int)a.num_rows, (int)a.num_nonzeros);
    auto &m = HandleCache::get().map;
    auto it = m.find(key);
    if (it != m.end()) return it->second;
    smle_csr_t h = nullptr;
    int rc;
    if constexpr (sizeof(V) == 8)
        rc = smle_csr_create_f64(&h, a.num_rows, a.num_cols, a.num_nonzeros, a.row_offsets, a.column_indices,
                                 (const double *)a.values);
    else
        rc = smle_csr_create_f32(&h, a.num_rows, a.num_cols, a.num_nonzeros, a.row_offsets, a.column_indices,
                                 (const float *)a.values);
    if (rc) die("smle_csr_create");
    m[key] = h;
    return h;
}

// drop the cached device copy (call before the host arrays are freed or modified)
template <typename CsrT>
void GpuCsrRelease(CsrT &a)
{
    auto &m = HandleCache::get().map;
    for (auto it = m.begin(); it != m.end();)
        if (std::get<0>(it->first) == (const void *)a.row_offsets) { smle_csr_destroy(it->second); it = m.erase(it); }
        else ++it;
}

} // namespace smle_adapters

enum GpuSpmmKernel { GPU_SIMPLE = 0, GPU_MERGE = 1, GPU_NONZERO_SPLIT = 2 };   // == SpmmKernel (types.hpp:11-16)

// MergePathSearch on list A = row end-offsets, list B = counting iterator (the only
// instantiation the reference uses); path_coordinate needs .x and .y like the reference's int2.
template <typename OffsetT, typename CoordinateT>
inline void GpuMergePathSearch(OffsetT diagonal, const OffsetT *a, OffsetT a_len, OffsetT b_len, CoordinateT &path_coordinate)
{
    // one boundary: share = diagonal, parts = 1 gives coordinates for diagonals 0 and min(diagonal, total)
    int xy[4];
    if (smle_merge_path_partition(a, a_len, b_len, 1, diagonal > 0 ? diagonal : 1, xy)) smle_adapters::die("smle_merge_path_partition");
    path_coordinate.x = diagonal > 0 ? xy[2] : xy[0];
    path_coordinate.y = diagonal > 0 ? xy[3] : xy[1];
}

template <typename ValueT, typename OffsetT, typename CsrT>
inline void GpuMergeCsrmv(int /*num_threads*/, CsrT &a, OffsetT * /*row_end_offsets*/, OffsetT * /*column_indices*/,
                          ValueT * /*values*/, ValueT *vector_x, ValueT *vector_y_out)
{
    smle_csr_t h = smle_adapters::handle_of(a);
    int rc;
    if constexpr (sizeof(ValueT) == 8) rc = smle_spmv_f64(h, vector_x, vector_y_out, 0);
    else rc = smle_spmv_f32(h, vector_x, vector_y_out, 0);
    if (rc) smle_adapters::die("smle_spmv");
}

template <typename ValueT, typename OffsetT, typename CsrT>
inline void GpuMergeCsrmm(int /*num_threads*/, CsrT &a, OffsetT * /*row_end_offsets*/, OffsetT * /*column_indices*/,
                          ValueT * /*values*/, ValueT *vector_x, ValueT *vector_y_out, int num_vectors)
{
    smle_csr_t h = smle_adapters::handle_of(a);
    int rc;
    if constexpr (sizeof(ValueT) == 8) rc = smle_spmm_f64(h, vector_x, vector_y_out, num_vectors, 0);
    else rc = smle_spmm_f32(h, vector_x, vector_y_out, num_vectors, 0);
    if (rc) smle_adapters::die("smle_spmm");
}

// the split strategy is a CPU threading choice: on the GPU both map to the merge-path kernel
template <typename ValueT, typename OffsetT, typename CsrT>
inline void GpuNonzeroSplitCsrmm(int t, CsrT &a, OffsetT *re, OffsetT *ci, ValueT *va, ValueT *x, ValueT *y, int k)
{
    GpuMergeCsrmm(t, a, re, ci, va, x, y, k);
}

template <typename ValueT, typename CsrT>
inline void GpuCsrSpmmT(int t, CsrT &a, ValueT *x, ValueT *y, int k)
{
    GpuMergeCsrmm(t, a, a.row_offsets + 1, a.column_indices, a.values, x, y, k);
}

template <typename ValueT, typename CsrT>
inline int GpuCGSolveSingle(CsrT &a, const ValueT *b, ValueT *x, int max_iters, ValueT tolerance)
{
    int iters = 0, rc;
    if constexpr (sizeof(ValueT) == 8) rc = smle_cg_single_f64(smle_adapters::handle_of(a), b, x, max_iters, tolerance, 0, &iters, nullptr);
    else rc = smle_cg_single_f32(smle_adapters::handle_of(a), b, x, max_iters, tolerance, 0, &iters, nullptr);
    if (rc) smle_adapters::die("smle_cg_single");
    return iters;
}

template <typename ValueT, typename CsrT>
inline int GpuCGSolveMultiple(CsrT &a, const ValueT *B, ValueT *X, int num_vectors, int max_iters, ValueT tolerance,
                              int kernel_type, std::vector<double> *max_errors = nullptr)
{
    int iters = 0, hist_len = 0, rc;
    std::vector<double> hist(max_errors ? (size_t)(max_iters > 0 ? max_iters : 1) : 0);
    if constexpr (sizeof(ValueT) == 8)
        rc = smle_cg_multi_f64(smle_adapters::handle_of(a), B, X, num_vectors, max_iters, tolerance, kernel_type, 0, &iters,
                               max_errors ? hist.data() : nullptr, (int)hist.size(), &hist_len, nullptr);
    else
        rc = smle_cg_multi_f32(smle_adapters::handle_of(a), B, X, num_vectors, max_iters, tolerance, kernel_type, 0, &iters,
                               max_errors ? hist.data() : nullptr, (int)hist.size(), &hist_len, nullptr);
    if (rc) smle_adapters::die("smle_cg_multi");
    if (max_errors) max_errors->assign(hist.begin(), hist.begin() + hist_len);
    return iters;
}

// SparseApproximateInversion (work_2025/cg/sparse_approximate_inversion.hpp:41-321): fills `l` with A's
// pattern and the SPAI values.  `l` must be default-constructed; its arrays are allocated with new[]
// (the reference's non-MKL branch, :71-75), so CsrMatrix's destructor releases them.
template <typename CsrT>
inline bool GpuSparseApproximateInversion(const CsrT &a, CsrT &l)
{
    using V = typename std::remove_pointer<decltype(a.values)>::type;
    using O = typename std::remove_pointer<decltype(a.row_offsets)>::type;
    static_assert(sizeof(V) == 8, "the reference drivers build SPAI for <double,int> only");
    l.num_rows = a.num_rows; l.num_cols = a.num_cols; l.num_nonzeros = a.num_nonzeros;
    l.row_offsets = new O[a.num_rows + 1];
    l.column_indices = new O[a.num_nonzeros > 0 ? a.num_nonzeros : 1];
    l.values = new V[a.num_nonzeros > 0 ? a.num_nonzeros : 1];
    for (long long i = 0; i <= a.num_rows; ++i) l.row_offsets[i] = a.row_offsets[i];
    for (long long i = 0; i < a.num_nonzeros; ++i) l.column_indices[i] = a.column_indices[i];
    return smle_spai_build_f64(a.num_rows, a.num_nonzeros, a.row_offsets, a.column_indices, (const double *)a.values,
                               (double *)l.values) == 0;
}

// SPAISolveMultiple (work_2025/main/sparse_approximate_inverse.hpp:31-230)
template <typename ValueT, typename CsrT>
inline int GpuSPAISolveMultiple(CsrT &a, CsrT &m, const ValueT *B, ValueT *X, int num_vectors, int max_iters, ValueT tolerance,
                                int kernel_type, std::vector<double> *max_errors = nullptr)
{
    static_assert(sizeof(ValueT) == 8, "the reference instantiates the solvers for <double,int> only");
    int iters = 0, hist_len = 0;
    std::vector<double> hist(max_errors ? (size_t)(max_iters > 0 ? max_iters : 1) : 0);
    if (smle_pcg_spai_multi_f64(smle_adapters::handle_of(a), smle_adapters::handle_of(m), B, X, num_vectors, max_iters, tolerance,
                                kernel_type, 0, &iters, max_errors ? hist.data() : nullptr, (int)hist.size(), &hist_len, nullptr))
        smle_adapters::die("smle_pcg_spai_multi_f64");
    if (max_errors) max_errors->assign(hist.begin(), hist.begin() + hist_len);
    return iters;
}

// TestCGMultipleSPAI (sparse_approximate_inverse.hpp:232-287): no warm-up solves, min over timed solves
template <typename ValueT, typename CsrT>
inline void TestGpuCGMultipleSPAI(CsrT &a, CsrT &m, ValueT *b_vectors, ValueT *x_solutions, int max_iters, ValueT tolerance,
                                  int num_vectors, int timing_iterations, int kernel_type, double &min_ms,
                                  double &iters_of_min_ms, std::vector<double> *max_errors = nullptr)
{
    min_ms = std::numeric_limits<double>::max();
    iters_of_min_ms = 0;
    for (int it = 0; it < timing_iterations; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        int iters = GpuSPAISolveMultiple(a, m, b_vectors, x_solutions, num_vectors, max_iters, tolerance, kernel_type,
                                         it == 0 ? max_errors : nullptr);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < min_ms) { min_ms = ms; iters_of_min_ms = iters; }
    }
}

// TestCGSolveSingle: L vectors, vector v = b_vectors[v*n ...] (column-major), solved one after
// another; min over timing_iterations of the total wall time; iterations summed over vectors.
template <typename ValueT, typename CsrT>
inline void TestGpuCGSolveSingle(CsrT &a, ValueT *b_vectors, ValueT *x_solutions, int max_iters, ValueT tolerance,
                                 int num_vectors, int timing_iterations, double &min_ms, double &iters_of_min_ms)
{
    const long long n = a.num_rows;
    min_ms = std::numeric_limits<double>::max();
    iters_of_min_ms = 0;
    // The driver's blocks come from mkl_malloc / std::vector, i.e. pageable memory: page-lock them for the
    // duration of the timing loop (outside it, like the reference's own untimed setup) so that the copies of
    // the neighbouring vectors really overlap with the solves.  A failure to register is not an error.
    const unsigned long long block_bytes = sizeof(ValueT) * (unsigned long long)n * (unsigned long long)num_vectors;
    const bool reg_b = smle_host_register(b_vectors, block_bytes) == 0, reg_x = smle_host_register(x_solutions, block_bytes) == 0;
    struct Unregister {
        void *b, *x;
        ~Unregister() { if (b) smle_host_unregister(b); if (x) smle_host_unregister(x); }
    } unregister{reg_b ? (void *)b_vectors : nullptr, reg_x ? (void *)x_solutions : nullptr};
    for (int it = 0; it < timing_iterations; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        long long total = 0;
        if constexpr (std::is_same<ValueT, double>::value) {
            // one call: the copies of neighbouring vectors overlap with the solves
            if (smle_cg_single_batch_f64(smle_adapters::handle_of(a), b_vectors, x_solutions, num_vectors, max_iters, tolerance,
                                         nullptr, &total))
                smle_adapters::die("smle_cg_single_batch_f64");
        } else {
            for (int v = 0; v < num_vectors; ++v)
                total += GpuCGSolveSingle(a, &b_vectors[v * n], &x_solutions[v * n], max_iters, tolerance);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < min_ms) { min_ms = ms; iters_of_min_ms = (double)total; }
    }
}

// TestCGMultipleRHS: timing_iterations warm-up solves, then timed solves, min wall time.
template <typename ValueT, typename CsrT>
inline void TestGpuCGMultipleRHS(CsrT &a, ValueT *b_vectors, ValueT *x_solutions, int max_iters, ValueT tolerance,
                                 int num_vectors, int timing_iterations, int kernel_type, double &min_ms,
                                 double &iters_of_min_ms, std::vector<double> *max_errors = nullptr)
{
    for (int it = 0; it < timing_iterations; ++it)
        GpuCGSolveMultiple(a, b_vectors, x_solutions, num_vectors, max_iters, tolerance, kernel_type);
    min_ms = std::numeric_limits<double>::max();
    iters_of_min_ms = 0;
    for (int it = 0; it < timing_iterations; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        int iters = GpuCGSolveMultiple(a, b_vectors, x_solutions, num_vectors, max_iters, tolerance, kernel_type,
                                       it == 0 ? max_errors : nullptr);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < min_ms) { min_ms = ms; iters_of_min_ms = iters; }
    }
}
