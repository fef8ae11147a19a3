// smle_gen.cpp -- host-side matrix / right-hand-side generators of libsmle_b200.so.
//
// These produce the INPUTS of the hot path (they are not part of it and do no SpMV/CG work).
// Output CSR is identical to what the reference builds with its COO generator followed by
// CsrMatrix::Init (sparse_matrix.h:668-733: stable sort by (row, col), duplicates kept), but is
// constructed directly in sorted order, in parallel, so that the 300^3 grid (188 M nonzeros)
// and the scale-24 R-MAT (268 M nonzeros) are ready in seconds.  tests/ compare every array
// with the reference generators on small sizes.
#include "../../include/smle_b200.h"

#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace {

// nnz formulas of sparse_matrix.h:466-474 and :541-550
void grid2d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    long long ww = (long long)w * w;
    long long z = 4LL * (w - 2) * (w - 2) + 3LL * 4 * (w - 2) + 2LL * 4 + (self_loop ? ww : 0);
    if (w == 1) z = self_loop ? 1 : 0;
    *m = *n = (int)ww;
    *nnz = (int)z;
}

void grid3d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    long long www = (long long)w * w * w;
    long long i3 = (long long)(w - 2) * (w - 2) * (w - 2), f = 6LL * (w - 2) * (w - 2), e = 12LL * (w - 2);
    long long z = 6 * i3 + 5 * f + 4 * e + 3 * 8 + (self_loop ? www : 0);
    if (w == 1) z = self_loop ? 1 : 0;
    *m = *n = (int)www;
    *nnz = (int)z;
}

// 2-D grid: sorted neighbours of me=(j,k): me-w (N), me-1 (W), [me], me+1 (E), me+w (S)
// (sparse_matrix.h:458-527 emits W,E,N,S,self; Init's sort gives the order above)
template <typename V>
int gen_grid2d(int w, int self_loop, V diag, V offd, int *ro, int *ci, V *va)
{
    if (w < 1 || !ro || !ci || !va) return SMLE_ERR_ARG;
    const int m = w * w;
#pragma omp parallel for schedule(static)
    for (int me = 0; me < m; ++me) {
        int j = me / w, k = me % w;
        // nonzeros before row me: closed form = sum of degrees of rows < me
        // computed incrementally below instead; here only the degree
        ro[me + 1] = (j > 0) + (k > 0) + (k + 1 < w) + (j + 1 < w) + (self_loop ? 1 : 0);
    }
    ro[0] = 0;
    for (int r = 0; r < m; ++r) ro[r + 1] += ro[r];
#pragma omp parallel for schedule(static)
    for (int me = 0; me < m; ++me) {
        int j = me / w, k = me % w, z = ro[me];
        if (j > 0)      { ci[z] = me - w; va[z++] = offd; }
        if (k > 0)      { ci[z] = me - 1; va[z++] = offd; }
        if (self_loop)  { ci[z] = me;     va[z++] = diag; }
        if (k + 1 < w)  { ci[z] = me + 1; va[z++] = offd; }
        if (j + 1 < w)  { ci[z] = me + w; va[z++] = offd; }
    }
    return SMLE_OK;
}

// 3-D grid: me=(i,j,k); sorted neighbours me-w^2, me-w, me-1, [me], me+1, me+w, me+w^2
// (sparse_matrix.h:533-623 emits -k,+k,-j,+j,-i,+i,self)
inline int grid3d_degree(int w, int ww, int me, int self_loop)
{
    const int i = me / ww, j = (me / w) % w, k = me % w;
    return (i > 0) + (j > 0) + (k > 0) + (k + 1 < w) + (j + 1 < w) + (i + 1 < w) + (self_loop ? 1 : 0);
}

int gen_grid3d_row_offsets(int w, int self_loop, int *ro)
{
    if (w < 1 || !ro) return SMLE_ERR_ARG;
    const int ww = w * w, m = ww * w;
#pragma omp parallel for schedule(static)
    for (int me = 0; me < m; ++me) ro[me + 1] = grid3d_degree(w, ww, me, self_loop);
    ro[0] = 0;
    for (int r = 0; r < m; ++r) ro[r + 1] += ro[r];
    return SMLE_OK;
}

// rows [r0, r1): local offsets from 0, global column indices
template <typename V>
int gen_grid3d_rows(int w, int self_loop, V diag, V offd, int r0, int r1, int *lro, int *ci, V *va)
{
    if (w < 1 || !lro) return SMLE_ERR_ARG;
    const int ww = w * w, m = ww * w;
    if (r0 < 0 || r1 < r0 || r1 > m) return SMLE_ERR_ARG;
    const int n = r1 - r0;
#pragma omp parallel for schedule(static)
    for (int l = 0; l < n; ++l) lro[l + 1] = grid3d_degree(w, ww, r0 + l, self_loop);
    lro[0] = 0;
    for (int l = 0; l < n; ++l) lro[l + 1] += lro[l];
    if (lro[n] > 0 && (!ci || !va)) return SMLE_ERR_ARG;
#pragma omp parallel for schedule(static)
    for (int l = 0; l < n; ++l) {
        const int me = r0 + l;
        int i = me / ww, j = (me / w) % w, k = me % w, z = lro[l];
        if (i > 0)      { ci[z] = me - ww; va[z++] = offd; }
        if (j > 0)      { ci[z] = me - w;  va[z++] = offd; }
        if (k > 0)      { ci[z] = me - 1;  va[z++] = offd; }
        if (self_loop)  { ci[z] = me;      va[z++] = diag; }
        if (k + 1 < w)  { ci[z] = me + 1;  va[z++] = offd; }
        if (j + 1 < w)  { ci[z] = me + w;  va[z++] = offd; }
        if (i + 1 < w)  { ci[z] = me + ww; va[z++] = offd; }
    }
    return SMLE_OK;
}

template <typename V>
int gen_grid3d(int w, int self_loop, V diag, V offd, int *ro, int *ci, V *va)
{
    if (w < 1 || !ro || !ci || !va) return SMLE_ERR_ARG;
    return gen_grid3d_rows<V>(w, self_loop, diag, offd, 0, w * w * w, ro, ci, va);
}

// wheel (sparse_matrix.h:417-450): hub row 0 -> 1..s; rim row i+1 -> ((i+1) % s) + 1
template <typename V>
int gen_wheel(int s, V value, int *ro, int *ci, V *va)
{
    if (s < 1 || !ro || !ci || !va) return SMLE_ERR_ARG;
    ro[0] = 0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s; ++i) {
        ci[i] = i + 1; va[i] = value;
        ro[i + 1] = s + i;
        ci[s + i] = ((i + 1) % s) + 1; va[s + i] = value;
    }
    ro[s + 1] = 2 * s;
    return SMLE_OK;
}

template <typename V>
int gen_dense(int rows, int cols, V value, int *ro, int *ci, V *va)
{
    if (rows < 0 || cols < 0 || !ro) return SMLE_ERR_ARG;
    for (int r = 0; r <= rows; ++r) ro[r] = r * cols;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) { ci[(size_t)r * cols + c] = c; va[(size_t)r * cols + c] = value; }
    return SMLE_OK;
}

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// R-MAT (Chakrabarti et al.): every edge picks one quadrant per level from 16 random bits;
// counter-based hashing makes edge e independent of the thread that generates it.
template <typename V>
int gen_rmat(int scale, int ef, double pa, double pb, double pc, uint64_t seed, int unit, int *ro, int *ci, V *va)
{
    if (scale < 1 || scale > 30 || ef < 1 || !ro || !ci || !va) return SMLE_ERR_ARG;
    const int m = 1 << scale;
    const long long E = (long long)ef * m;
    if (E > INT32_MAX) return SMLE_ERR_RANGE;
    const uint32_t ta = (uint32_t)(pa * 65536.0), tb = ta + (uint32_t)(pb * 65536.0),
                   tc = tb + (uint32_t)(pc * 65536.0);
    std::vector<int> erow((size_t)E), ecol((size_t)E);
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < E; ++e) {
        int r = 0, c = 0;
        uint64_t bits = 0;
        for (int l = 0; l < scale; ++l) {
            if ((l & 3) == 0) bits = splitmix64(seed ^ ((uint64_t)e * 8 + (uint64_t)(l >> 2)) * 0xD1342543DE82EF95ull);
            uint32_t u = (uint32_t)(bits & 0xFFFF);
            bits >>= 16;
            int rb = u >= tb, cb = (u >= ta && u < tb) || u >= tc;
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        erow[(size_t)e] = r;
        ecol[(size_t)e] = c;
    }
    // rows: histogram -> offsets -> stable scatter (edge order kept inside a row)
    std::vector<int> cur((size_t)m + 1, 0);
    for (long long e = 0; e < E; ++e) ++cur[(size_t)erow[(size_t)e] + 1];
    ro[0] = 0;
    for (int r = 0; r < m; ++r) ro[r + 1] = ro[r] + cur[(size_t)r + 1];
    for (int r = 0; r < m; ++r) cur[(size_t)r] = ro[r];
    std::vector<long long> eid((size_t)E);
    for (long long e = 0; e < E; ++e) eid[(size_t)cur[(size_t)erow[(size_t)e]]++] = e;
    // per row: stable sort by column, duplicates kept (CsrMatrix::Init semantics)
#pragma omp parallel for schedule(dynamic, 1024)
    for (int r = 0; r < m; ++r) {
        long long *b = eid.data() + ro[r], *e = eid.data() + ro[r + 1];
        std::stable_sort(b, e, [&](long long x, long long y) { return ecol[(size_t)x] < ecol[(size_t)y]; });
        for (long long *p = b; p < e; ++p) {
            size_t z = (size_t)(p - eid.data());
            ci[z] = ecol[(size_t)*p];
            if (unit) va[z] = (V)1.0;
            else {
                uint64_t h = splitmix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)*p + 0x1234567ull);
                va[z] = (V)(((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0));  // (0, 1]
            }
        }
    }
    return SMLE_OK;
}

} // namespace

extern "C" {

int smle_gen_grid2d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    if (w < 1 || !m || !n || !nnz || (long long)w * w * 5 > INT32_MAX) return SMLE_ERR_ARG;
    grid2d_shape(w, self_loop, m, n, nnz);
    return SMLE_OK;
}
int smle_gen_grid3d_shape(int w, int self_loop, int *m, int *n, int *nnz)
{
    if (w < 1 || !m || !n || !nnz || (long long)w * w * w * 7 > INT32_MAX) return SMLE_ERR_ARG;
    grid3d_shape(w, self_loop, m, n, nnz);
    return SMLE_OK;
}
int smle_gen_wheel_shape(int s, int *m, int *n, int *nnz)
{
    if (s < 1 || !m || !n || !nnz || s > INT32_MAX / 2 - 1) return SMLE_ERR_ARG;
    *m = *n = s + 1; *nnz = 2 * s;
    return SMLE_OK;
}
int smle_gen_dense_shape(int rows, int cols, int *m, int *n, int *nnz)
{
    if (rows < 0 || cols < 0 || !m || !n || !nnz || (long long)rows * cols > INT32_MAX) return SMLE_ERR_ARG;
    *m = rows; *n = cols; *nnz = rows * cols;
    return SMLE_OK;
}
int smle_gen_rmat_shape(int scale, int ef, int *m, int *n, int *nnz)
{
    if (scale < 1 || scale > 30 || ef < 1 || !m || !n || !nnz || ((long long)ef << scale) > INT32_MAX) return SMLE_ERR_ARG;
    *m = *n = 1 << scale; *nnz = ef << scale;
    return SMLE_OK;
}

int smle_gen_grid2d_f64(int w, int sl, double d, double o, int *ro, int *ci, double *va) { return gen_grid2d<double>(w, sl, d, o, ro, ci, va); }
int smle_gen_grid2d_f32(int w, int sl, float d, float o, int *ro, int *ci, float *va) { return gen_grid2d<float>(w, sl, d, o, ro, ci, va); }
int smle_gen_grid3d_f64(int w, int sl, double d, double o, int *ro, int *ci, double *va) { return gen_grid3d<double>(w, sl, d, o, ro, ci, va); }
int smle_gen_grid3d_f32(int w, int sl, float d, float o, int *ro, int *ci, float *va) { return gen_grid3d<float>(w, sl, d, o, ro, ci, va); }
int smle_gen_grid3d_row_offsets(int w, int sl, int *ro)
{
    if (w < 1 || (long long)w * w * w * 7 > INT32_MAX) return SMLE_ERR_ARG;
    return gen_grid3d_row_offsets(w, sl, ro);
}
int smle_gen_grid3d_rows_f64(int w, int sl, double d, double o, int r0, int r1, int *lro, int *ci, double *va)
{
    if (w < 1 || (long long)w * w * w * 7 > INT32_MAX) return SMLE_ERR_ARG;
    return gen_grid3d_rows<double>(w, sl, d, o, r0, r1, lro, ci, va);
}
int smle_gen_wheel_f64(int s, double v, int *ro, int *ci, double *va) { return gen_wheel<double>(s, v, ro, ci, va); }
int smle_gen_wheel_f32(int s, float v, int *ro, int *ci, float *va) { return gen_wheel<float>(s, v, ro, ci, va); }
int smle_gen_dense_f64(int r, int c, double v, int *ro, int *ci, double *va) { return gen_dense<double>(r, c, v, ro, ci, va); }
int smle_gen_dense_f32(int r, int c, float v, int *ro, int *ci, float *va) { return gen_dense<float>(r, c, v, ro, ci, va); }
int smle_gen_rmat_f64(int s, int ef, double a, double b, double c, unsigned long long seed, int unit, int *ro, int *ci, double *va)
{
    return gen_rmat<double>(s, ef, a, b, c, seed, unit, ro, ci, va);
}
int smle_gen_rmat_f32(int s, int ef, double a, double b, double c, unsigned long long seed, int unit, int *ro, int *ci, float *va)
{
    return gen_rmat<float>(s, ef, a, b, c, seed, unit, ro, ci, va);
}

int smle_gen_rhs_rand_f64(unsigned seed, long long count, double *out)
{
    if (count < 0 || (count > 0 && !out)) return SMLE_ERR_ARG;
    srand(seed);
    for (long long i = 0; i < count; ++i) out[i] = (double)rand() / (double)RAND_MAX;
    return SMLE_OK;
}

int smle_gen_rhs_rand_range_f64(unsigned seed, long long first, long long count, double *out)
{
    if (first < 0 || count < 0 || (count > 0 && !out)) return SMLE_ERR_ARG;
    srand(seed);
    for (long long i = 0; i < first; ++i) (void)rand();
    for (long long i = 0; i < count; ++i) out[i] = (double)rand() / (double)RAND_MAX;
    return SMLE_OK;
}

double smle_driver_threshold_f64(const double *b, int n, double tol)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (int i = 0; i < n; ++i) s += b[i] * b[i];
    return sqrt(s) * tol;
}

} // extern "C"
