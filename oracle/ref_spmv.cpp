// oracle/ref_spmv.cpp -- C-ABI wrappers around the UNMODIFIED reference cpu_spmv.cpp.
//
// TEST INFRASTRUCTURE ONLY.  OmpMergeCsrmv (cpu_spmv.cpp:360-421) is not in a header, so
// the whole driver TU is included where it lies under /root/reference with its main()
// renamed; nothing is copied into this repo.  Built by oracle/Makefile into
// oracle/_ref/libsmle_ref_spmv.so (separate .so: cpu_spmv.cpp re-defines MergePathSearch,
// SpmvGold and the g_* globals that the work_2025 headers also define).
#define main ref_cpu_spmv_main
#include "cpu_spmv.cpp"
#undef main

extern "C" {

void ref_spmv_set_threads(int t) { g_omp_threads = t; omp_set_num_threads(t); g_quiet = true; }

#define REF_SPMV_DEFINE(V, S)                                                                    \
    void ref_merge_csrmv_##S(int T, int m, int n, int nnz, const int *ro, const int *ci,         \
                             const V *va, const V *x, V *y)                                      \
    {                                                                                            \
        CsrMatrix<V, int> a;                                                                     \
        a.num_rows = m; a.num_cols = n; a.num_nonzeros = nnz;                                    \
        a.row_offsets = const_cast<int *>(ro);                                                   \
        a.column_indices = const_cast<int *>(ci);                                                \
        a.values = const_cast<V *>(va);                                                          \
        OmpMergeCsrmv(T, a, a.row_offsets + 1, a.column_indices, a.values,                       \
                      const_cast<V *>(x), y);                                                    \
        a.row_offsets = NULL; a.column_indices = NULL; a.values = NULL;                          \
    }                                                                                            \
    /* TestOmpMergeCsrmv (cpu_spmv.cpp:429-475): the reference's own timing loop */             \
    float ref_test_merge_csrmv_##S(int T, int m, int n, int nnz, const int *ro, const int *ci,   \
                                   const V *va, V *x, V *y_ref, V *y, int timing_iters)          \
    {                                                                                            \
        CsrMatrix<V, int> a;                                                                     \
        a.num_rows = m; a.num_cols = n; a.num_nonzeros = nnz;                                    \
        a.row_offsets = const_cast<int *>(ro);                                                   \
        a.column_indices = const_cast<int *>(ci);                                                \
        a.values = const_cast<V *>(va);                                                          \
        g_omp_threads = T; g_quiet = true;                                                       \
        float setup_ms = 0.f;                                                                    \
        float ms = TestOmpMergeCsrmv(a, x, y_ref, y, timing_iters, setup_ms);                    \
        a.row_offsets = NULL; a.column_indices = NULL; a.values = NULL;                          \
        return ms;                                                                               \
    }

REF_SPMV_DEFINE(double, f64)
REF_SPMV_DEFINE(float, f32)

} // extern "C"
