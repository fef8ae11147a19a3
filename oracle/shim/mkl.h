/* oracle/shim/mkl.h -- stand-in for Intel MKL (not installed in this image) so the
 * UNMODIFIED reference headers compile.  Only the allocation entry points the hot
 * path uses are real; the sparse-BLAS calls used by the reference's off-path MKL
 * comparison baseline (cpu_spmv.cpp:640,656) are plain CSR loops so that
 * cpu_spmv.cpp links.  Test infrastructure only. */
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
#endif
