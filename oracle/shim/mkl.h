/* oracle/shim/mkl.h -- stand-in for Intel MKL (not installed in this image) so the
 * UNMODIFIED reference headers compile.  Only the allocation entry points the hot
 * path uses are real; the sparse-BLAS calls used by the reference's off-path MKL
 * comparison baseline (cpu_spmv.cpp:640,656) are plain CSR loops so that
 * cpu_spmv.cpp links; LAPACKE_?gels (the SPAI setup's least-squares solve,
 * work_2025/cg/sparse_approximate_inversion.hpp:22,32) is a Householder QR.  Test infrastructure only. */
#ifndef SMLE_SHIM_MKL_H
#define SMLE_SHIM_MKL_H
#include <stdlib.h>
static inline void *mkl_malloc(size_t bytes, int align)
{
    void *p = NULL;
    if (align < (int)sizeof(void *)) align = sizeof(void *);
    if (posix_memalign(&p, (size_t)align, bytes ? bytes : 1) != 0) return NULL;
    return p;
}
static inline void mkl_free(void *p) { free(p); }
static inline void mkl_set_num_threads(int) {}
#define SMLE_SHIM_CSRGEMV(NAME, T)                                                          \
    static inline void NAME(const char *, const int *m, const T *a, const int *ia,          \
                            const int *ja, const T *x, T *y)                                \
    {                                                                                       \
        for (int r = 0; r < *m; ++r) {                                                      \
            T s = 0;                                                                        \
            for (int o = ia[r]; o < ia[r + 1]; ++o) s += a[o] * x[ja[o]];                   \
            y[r] = s;                                                                       \
        }                                                                                   \
    }
SMLE_SHIM_CSRGEMV(mkl_cspblas_scsrgemv, float)
SMLE_SHIM_CSRGEMV(mkl_cspblas_dcsrgemv, double)

/* LAPACKE_?gels for the one way the reference calls it (sparse_approximate_inversion.hpp:212-222):
 * row-major, no transpose, one right-hand side, m >= n.  Householder QR of A, Q^T b, back
 * substitution; the solution overwrites b[0:n].  info > 0: a zero on the diagonal of R (rank
 * deficient), like LAPACK; m < n (never produced by a static pattern that contains the diagonal)
 * is reported the same way. */
typedef int lapack_int;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#include <math.h>
#define SMLE_SHIM_GELS(NAME, T)                                                                   \
    static inline lapack_int NAME(int layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, \
                                  T *a, lapack_int lda, T *b, lapack_int ldb)                     \
    {                                                                                             \
        if (layout != LAPACK_ROW_MAJOR || (trans != 'N' && trans != 'n') || nrhs != 1 || m < n) return n + 1; \
        for (lapack_int j = 0; j < n; ++j) {                                                      \
            double norm = 0.0;                                                                    \
            for (lapack_int i = j; i < m; ++i) norm += (double)a[i * lda + j] * (double)a[i * lda + j]; \
            norm = sqrt(norm);                                                                    \
            if (norm == 0.0) return j + 1;                                                        \
            const double ajj = (double)a[j * lda + j];                                            \
            const double alpha = ajj > 0.0 ? -norm : norm;                                        \
            /* v = x - alpha e1, stored in place below the diagonal, v0 kept apart */             \
            const double v0 = ajj - alpha;                                                        \
            double vnorm2 = v0 * v0;                                                              \
            for (lapack_int i = j + 1; i < m; ++i) vnorm2 += (double)a[i * lda + j] * (double)a[i * lda + j]; \
            if (vnorm2 > 0.0) {                                                                   \
                for (lapack_int c = j + 1; c < n; ++c) {                                          \
                    double dot = v0 * (double)a[j * lda + c];                                     \
                    for (lapack_int i = j + 1; i < m; ++i) dot += (double)a[i * lda + j] * (double)a[i * lda + c]; \
                    const double f = 2.0 * dot / vnorm2;                                          \
                    a[j * lda + c] = (T)((double)a[j * lda + c] - f * v0);                        \
                    for (lapack_int i = j + 1; i < m; ++i)                                        \
                        a[i * lda + c] = (T)((double)a[i * lda + c] - f * (double)a[i * lda + j]); \
                }                                                                                 \
                double dot = v0 * (double)b[j * ldb];                                             \
                for (lapack_int i = j + 1; i < m; ++i) dot += (double)a[i * lda + j] * (double)b[i * ldb]; \
                const double f = 2.0 * dot / vnorm2;                                              \
                b[j * ldb] = (T)((double)b[j * ldb] - f * v0);                                    \
                for (lapack_int i = j + 1; i < m; ++i) b[i * ldb] = (T)((double)b[i * ldb] - f * (double)a[i * lda + j]); \
            }                                                                                     \
            a[j * lda + j] = (T)alpha;                                                            \
        }                                                                                         \
        for (lapack_int j = n - 1; j >= 0; --j) {                                                 \
            double s = (double)b[j * ldb];                                                        \
            for (lapack_int c = j + 1; c < n; ++c) s -= (double)a[j * lda + c] * (double)b[c * ldb]; \
            b[j * ldb] = (T)(s / (double)a[j * lda + j]);                                         \
        }                                                                                         \
        return 0;                                                                                 \
    }
SMLE_SHIM_GELS(LAPACKE_dgels, double)
SMLE_SHIM_GELS(LAPACKE_sgels, float)
#endif
