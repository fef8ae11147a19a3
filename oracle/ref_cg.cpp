// oracle/ref_cg.cpp -- C-ABI wrappers around the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile against the sources where
// they lie under /root/reference (never copied into this repo) into
// oracle/_ref/libsmle_ref_cg.so.  Used to (a) validate oracle/smle_oracle.c,
// (b) generate tests/golden/*.json, (c) serve as bench.py's cpu_baseline
// kind="reference".  One TU only: work_2025/hyper_parameters.hpp defines its globals
// non-inline.
#include <omp.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sparse_matrix.h"
#include "utils.h"
#include "work_2025/hyper_parameters.hpp"
#include "work_2025/main/no_pretreatment.hpp"
#include "work_2025/main/single_strategy.hpp"
#include "work_2025/main/sparse_approximate_inverse.hpp"

namespace {

// Borrowed view: the wrapper never owns the caller's arrays, so the pointers are
// nulled before CsrMatrix::~CsrMatrix -> Clear() (sparse_matrix.h:738-769) can free them.
template <typename V>
struct CsrView {
    CsrMatrix<V, int> a;
    CsrView(int m, int n, int nnz, const int *ro, const int *ci, const V *va)
    {
        a.num_rows = m; a.num_cols = n; a.num_nonzeros = nnz;
        a.row_offsets = const_cast<int *>(ro);
        a.column_indices = const_cast<int *>(ci);
        a.values = const_cast<V *>(va);
    }
    ~CsrView() { a.row_offsets = NULL; a.column_indices = NULL; a.values = NULL; }
};

template <typename V>
void copy_out(CsrMatrix<V, int> &csr, int *ro, int *ci, V *va)
{
    for (int i = 0; i <= csr.num_rows; ++i) ro[i] = csr.row_offsets[i];
    for (int i = 0; i < csr.num_nonzeros; ++i) { ci[i] = csr.column_indices[i]; va[i] = csr.values[i]; }
}

// Poisson fill of SURVEY.md Appendix B: tuple value chosen before CsrMatrix::Init.
template <typename V>
void fill(CooMatrix<V, int> &coo, V diag, V offd)
{
    for (int i = 0; i < coo.num_nonzeros; ++i)
        coo.coo_tuples[i].val = (coo.coo_tuples[i].row == coo.coo_tuples[i].col) ? diag : offd;
}

} // namespace

extern "C" {

void ref_set_threads(int t) { g_omp_threads = t; omp_set_num_threads(t); g_quiet = true; }

void ref_merge_path_search(int diagonal, const int *a, int a_len, int b_len, int *out_xy)
{
    CountingInputIterator<int> b(0);
    int2 c;
    MergePathSearch(diagonal, const_cast<int *>(a), b, a_len, b_len, c);
    out_xy[0] = c.x; out_xy[1] = c.y;
}

#define REF_DEFINE(V, S)                                                                         \
    void ref_merge_csrmm_##S(int T, int m, int n, int nnz, const int *ro, const int *ci,         \
                             const V *va, const V *X, V *Y, int k)                               \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        OmpMergeCsrmm(T, v.a, v.a.row_offsets + 1, v.a.column_indices, v.a.values,               \
                      const_cast<V *>(X), Y, k);                                                 \
    }                                                                                            \
    void ref_nonzero_split_csrmm_##S(int T, int m, int n, int nnz, const int *ro, const int *ci, \
                                     const V *va, const V *X, V *Y, int k)                       \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        OmpNonzeroSplitCsrmm(T, v.a, v.a.row_offsets + 1, v.a.column_indices, v.a.values,        \
                             const_cast<V *>(X), Y, k);                                          \
    }                                                                                            \
    void ref_row_split_csrmm_##S(int T, int m, int n, int nnz, const int *ro, const int *ci,     \
                                 const V *va, const V *X, V *Y, int k)                           \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        OmpCsrSpmmT(T, v.a, const_cast<V *>(X), Y, k);                                           \
    }                                                                                            \
    void ref_spmv_gold_##S(int m, int n, int nnz, const int *ro, const int *ci, const V *va,     \
                           const V *x, const V *y_in, V *y_out, V alpha, V beta)                 \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        SpmvGold(v.a, const_cast<V *>(x), const_cast<V *>(y_in), y_out, alpha, beta);            \
    }                                                                                            \
    int ref_cg_single_##S(int m, int n, int nnz, const int *ro, const int *ci, const V *va,      \
                          const V *b, V *x, int max_iters, V tol)                                \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        return CGSolveSingle(v.a, b, x, max_iters, tol);                                         \
    }                                                                                            \
    int ref_cg_multi_##S(int m, int n, int nnz, const int *ro, const int *ci, const V *va,       \
                         const V *B, V *X, int k, int max_iters, V tol, int kernel,              \
                         double *hist, int *hist_len)                                            \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        std::vector<double> errs;                                                                \
        int it = CGSolveMultiple(v.a, B, X, k, max_iters, tol, (SpmmKernel)kernel,               \
                                 hist ? &errs : nullptr);                                        \
        if (hist) for (size_t i = 0; i < errs.size(); ++i) hist[i] = errs[i];                    \
        if (hist_len) *hist_len = (int)errs.size();                                              \
        return it;                                                                               \
    }                                                                                            \
    /* TestCGSolveSingle / TestCGMultipleRHS: the reference's own timing wrappers */             \
    void ref_test_cg_single_##S(int m, int n, int nnz, const int *ro, const int *ci,             \
                                const V *va, V *b, V *x, int max_iters, V tol, int num_vectors,  \
                                int timing_iters, double *min_ms, double *iters)                 \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        TestCGSolveSingle(v.a, b, x, max_iters, tol, num_vectors, timing_iters, *min_ms, *iters);\
    }                                                                                            \
    void ref_test_cg_multi_##S(int m, int n, int nnz, const int *ro, const int *ci, const V *va, \
                               V *B, V *X, int max_iters, V tol, int k, int timing_iters,        \
                               int kernel, double *min_ms, double *iters)                        \
    {                                                                                            \
        CsrView<V> v(m, n, nnz, ro, ci, va);                                                     \
        TestCGMultipleRHS(v.a, B, X, max_iters, tol, k, timing_iters, (SpmmKernel)kernel,        \
                          *min_ms, *iters, nullptr);                                             \
    }                                                                                            \
    /* generators through the reference's CooMatrix -> CsrMatrix::Init */                        \
    void ref_gen_grid2d_##S(int w, int self_loop, V diag, V offd, int *ro, int *ci, V *va)       \
    {                                                                                            \
        CooMatrix<V, int> coo; coo.InitGrid2d(w, self_loop != 0); fill(coo, diag, offd);         \
        CsrMatrix<V, int> csr(coo); copy_out(csr, ro, ci, va);                                   \
    }                                                                                            \
    void ref_gen_grid3d_##S(int w, int self_loop, V diag, V offd, int *ro, int *ci, V *va)       \
    {                                                                                            \
        CooMatrix<V, int> coo; coo.InitGrid3d(w, self_loop != 0); fill(coo, diag, offd);         \
        CsrMatrix<V, int> csr(coo); copy_out(csr, ro, ci, va);                                   \
    }                                                                                            \
    void ref_gen_wheel_##S(int spokes, V value, int *ro, int *ci, V *va)                         \
    {                                                                                            \
        CooMatrix<V, int> coo; coo.InitWheel(spokes, value);                                     \
        CsrMatrix<V, int> csr(coo); copy_out(csr, ro, ci, va);                                   \
    }                                                                                            \
    void ref_gen_dense_##S(int rows, int cols, V value, int *ro, int *ci, V *va)                 \
    {                                                                                            \
        CooMatrix<V, int> coo; coo.InitDense(rows, cols, value);                                 \
        CsrMatrix<V, int> csr(coo); copy_out(csr, ro, ci, va);                                   \
    }

REF_DEFINE(double, f64)
REF_DEFINE(float, f32)

// SPAI (rank "next" N3): SparseApproximateInversion (work_2025/cg/sparse_approximate_inversion.hpp:41-321,
// static pattern S_M = S_A, per-column least squares through the shim's LAPACKE_dgels, symmetrised)
// and SPAISolveMultiple (work_2025/main/sparse_approximate_inverse.hpp:31-230).
int ref_spai_build_f64(int m, int n, int nnz, const int *ro, const int *ci, const double *va, double *m_values)
{
    CsrView<double> v(m, n, nnz, ro, ci, va);
    CsrMatrix<double, int> l;
    bool ok = SparseApproximateInversion(v.a, l);
    if (ok) for (int i = 0; i < nnz; ++i) m_values[i] = l.values[i];
    return ok ? 0 : 1;
}

int ref_spai_solve_multi_f64(int m, int n, int nnz, const int *ro, const int *ci, const double *va,
                             const double *m_values, const double *B, double *X, int k, int max_iters, double tol,
                             int kernel, double *hist, int *hist_len)
{
    CsrView<double> a(m, n, nnz, ro, ci, va), pm(m, n, nnz, ro, ci, m_values);
    std::vector<double> errs;
    int it = SPAISolveMultiple(a.a, pm.a, B, X, k, max_iters, tol, (SpmmKernel)kernel, hist ? &errs : nullptr);
    if (hist) for (size_t i = 0; i < errs.size(); ++i) hist[i] = errs[i];
    if (hist_len) *hist_len = (int)errs.size();
    return it;
}

// CooMatrix::InitMarket (sparse_matrix.h:211-380) + CsrMatrix::Init; two-call protocol:
// ro == NULL returns the shape only.
int ref_read_mtx_f64(const char *path, int *m, int *n, int *nnz, int *ro, int *ci, double *va)
{
    CooMatrix<double, int> coo;
    coo.InitMarket(std::string(path), 1.0, false);
    CsrMatrix<double, int> csr(coo);
    *m = csr.num_rows; *n = csr.num_cols; *nnz = csr.num_nonzeros;
    if (ro) copy_out(csr, ro, ci, va);
    return 0;
}

void ref_gen_grid2d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    CooMatrix<float, int> coo; coo.InitGrid2d(w, self_loop != 0);
    *m = coo.num_rows; *n = coo.num_cols; *nnz = coo.num_nonzeros;
}
void ref_gen_grid3d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    CooMatrix<float, int> coo; coo.InitGrid3d(w, self_loop != 0);
    *m = coo.num_rows; *n = coo.num_cols; *nnz = coo.num_nonzeros;
}

} // extern "C"
