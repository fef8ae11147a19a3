/*
 * oracle/smle_oracle_impl.h -- value-type-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see smle_oracle.c header).  Included twice by
 * smle_oracle.c with
 *     VT   = double / float
 *     SFX  = f64 / f32
 * Each function cites the reference file:line (under /root/reference) whose
 * algorithm it restates.  Nothing here is copied; every loop is re-derived
 * from the behaviour described in SURVEY.md section 8(a).
 */

#define ORC_CAT_(a, b) a##_##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)
#define FN(name) ORC_CAT(name, SFX)

/* ------------------------------------------------------------------ */
/* Serial gold SpMV: y_out = beta*y_in + alpha*A*x                      */
/* restates work_2025/spmm/sample.hpp:11-34 (== cpu_spmv.cpp:245-265)   */
/* ------------------------------------------------------------------ */
void FN(orc_spmv_gold)(int m, const int *ro, const int *ci, const VT *va,
                       const VT *x, const VT *y_in, VT *y_out, VT alpha, VT beta)
{
    for (int r = 0; r < m; ++r) {
        VT acc = beta * y_in[r];
        for (int o = ro[r]; o < ro[r + 1]; ++o)
            acc += alpha * va[o] * x[ci[o]];
        y_out[r] = acc;
    }
}

/* ------------------------------------------------------------------ */
/* Merge-path SpMV with T shares + serial carry fix-up                  */
/* restates cpu_spmv.cpp:360-421 (OmpMergeCsrmv)                        */
/*  - shares of ceil((m+nnz)/T) merge items (:379-386)                  */
/*  - whole rows (:392-401), trailing partial row (:404-408)            */
/*  - carry arrays are fixed [256] in the reference (:370-371)          */
/*  - fix-up runs over tid < T-1 only (:416-420)                        */
/* returns 0, or -1 when T exceeds the reference's 256-entry carry array */
/* ------------------------------------------------------------------ */
int FN(orc_merge_csrmv)(int T, int m, int nnz, const int *row_end, const int *ci,
                        const VT *va, const VT *x, VT *y)
{
    if (T < 1 || T > 256) return -1;
    int carry_row[256];
    VT carry_val[256];

#pragma omp parallel for schedule(static) num_threads(T)
    for (int tid = 0; tid < T; ++tid) {
        int total = m + nnz;
        int share = (total + T - 1) / T;
        int d0 = share * tid < total ? share * tid : total;
        int d1 = d0 + share < total ? d0 + share : total;
        int c0[2], c1[2];
        orc_merge_path_search(d0, row_end, m, nnz, c0);
        orc_merge_path_search(d1, row_end, m, nnz, c1);
        int r = c0[0], z = c0[1];
        for (; r < c1[0]; ++r) {
            VT acc = 0.0;
            for (; z < row_end[r]; ++z)
                acc += va[z] * x[ci[z]];
            y[r] = acc;
        }
        VT tail = 0.0;
        for (; z < c1[1]; ++z)
            tail += va[z] * x[ci[z]];
        carry_row[tid] = c1[0];
        carry_val[tid] = tail;
    }
    for (int tid = 0; tid < T - 1; ++tid)
        if (carry_row[tid] < m)
            y[carry_row[tid]] += carry_val[tid];
    return 0;
}

/* ------------------------------------------------------------------ */
/* Merge-path SpMM, X (n x k) and Y (m x k) row-major                   */
/* restates work_2025/spmm/merge_based.hpp:49-153 (OmpMergeCsrmm)       */
/*  - heap carry-out T x k (:60-61); fix-up over ALL tid with           */
/*    row < m guard (:138-149)                                          */
/* ------------------------------------------------------------------ */
void FN(orc_merge_csrmm)(int T, int m, int nnz, const int *row_end, const int *ci,
                         const VT *va, const VT *X, VT *Y, int k)
{
    int *carry_row = (int *)malloc(sizeof(int) * (size_t)T);
    VT *carry_val = (VT *)malloc(sizeof(VT) * (size_t)T * (size_t)k);

#pragma omp parallel for schedule(static) num_threads(T)
    for (int tid = 0; tid < T; ++tid) {
        int total = m + nnz;
        int share = (total + T - 1) / T;
        int d0 = share * tid < total ? share * tid : total;
        int d1 = d0 + share < total ? d0 + share : total;
        int c0[2], c1[2];
        orc_merge_path_search(d0, row_end, m, nnz, c0);
        orc_merge_path_search(d1, row_end, m, nnz, c1);
        VT *acc = (VT *)malloc(sizeof(VT) * (size_t)k);
        for (int i = 0; i < k; ++i) acc[i] = 0.0;
        int r = c0[0], z = c0[1];
        for (; r < c1[0]; ++r) {
            for (; z < row_end[r]; ++z) {
                VT v = va[z];
                const VT *xr = X + ci[z] * k; /* int product, as merge_based.hpp:97 */
                for (int i = 0; i < k; ++i) acc[i] += v * xr[i];
            }
            VT *yr = Y + r * k; /* merge_based.hpp:106 */
            for (int i = 0; i < k; ++i) { yr[i] = acc[i]; acc[i] = 0.0; }
        }
        for (; z < c1[1]; ++z) {
            VT v = va[z];
            const VT *xr = X + ci[z] * k;
            for (int i = 0; i < k; ++i) acc[i] += v * xr[i];
        }
        carry_row[tid] = c1[0];
        for (int i = 0; i < k; ++i) carry_val[(size_t)tid * k + i] = acc[i];
        free(acc);
    }
    for (int tid = 0; tid < T; ++tid) {
        int r = carry_row[tid];
        if (r < m)
            for (int i = 0; i < k; ++i)
                Y[r * k + i] += carry_val[(size_t)tid * k + i];
    }
    free(carry_val);
    free(carry_row);
}

/* ------------------------------------------------------------------ */
/* Even-nnz split SpMM                                                  */
/* restates work_2025/spmm/nonzero_splitting.hpp:52-150 and its         */
/* RowPathSearch (:19-44): x = first row whose end offset > y-1         */
/* ------------------------------------------------------------------ */
static int FN(orc_row_path_search)(const int *row_end, int m, int y)
{
    if (y == 0) return 0;
    int lo = 0, hi = m;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row_end[mid] <= y - 1) lo = mid + 1; else hi = mid;
    }
    return lo < m ? lo : m;
}

void FN(orc_nonzero_split_csrmm)(int T, int m, int nnz, const int *row_end, const int *ci,
                                 const VT *va, const VT *X, VT *Y, int k)
{
    int *carry_row = (int *)malloc(sizeof(int) * (size_t)T);
    VT *carry_val = (VT *)malloc(sizeof(VT) * (size_t)T * (size_t)k);

#pragma omp parallel for schedule(static) num_threads(T)
    for (int tid = 0; tid < T; ++tid) {
        int share = (nnz + T - 1) / T;
        int z = share * tid < nnz ? share * tid : nnz;
        int z1 = z + share < nnz ? z + share : nnz;
        int r = FN(orc_row_path_search)(row_end, m, z);
        int r1 = FN(orc_row_path_search)(row_end, m, z1);
        VT *acc = (VT *)malloc(sizeof(VT) * (size_t)k);
        for (int i = 0; i < k; ++i) acc[i] = 0.0;
        for (; r < r1; ++r) {
            for (; z < row_end[r]; ++z) {
                VT v = va[z];
                const VT *xr = X + ci[z] * k;
                for (int i = 0; i < k; ++i) acc[i] += v * xr[i];
            }
            VT *yr = Y + r * k;
            for (int i = 0; i < k; ++i) { yr[i] = acc[i]; acc[i] = 0.0; }
        }
        for (; z < z1; ++z) {
            VT v = va[z];
            const VT *xr = X + ci[z] * k;
            for (int i = 0; i < k; ++i) acc[i] += v * xr[i];
        }
        carry_row[tid] = r1;
        for (int i = 0; i < k; ++i) carry_val[(size_t)tid * k + i] = acc[i];
        free(acc);
    }
    for (int tid = 0; tid < T; ++tid) {
        int r = carry_row[tid];
        if (r < m)
            for (int i = 0; i < k; ++i)
                Y[r * k + i] += carry_val[(size_t)tid * k + i];
    }
    free(carry_val);
    free(carry_row);
}

/* ------------------------------------------------------------------ */
/* Row-split SpMM  -- restates work_2025/spmm/row_splitting.hpp:18-54   */
/* ------------------------------------------------------------------ */
void FN(orc_row_split_csrmm)(int T, int m, const int *ro, const int *ci, const VT *va,
                             const VT *X, VT *Y, int k)
{
#pragma omp parallel for schedule(static) num_threads(T)
    for (int r = 0; r < m; ++r) {
        VT *acc = (VT *)alloca(sizeof(VT) * (size_t)k);
        for (int i = 0; i < k; ++i) acc[i] = 0.0;
        for (int o = ro[r]; o < ro[r + 1]; ++o) {
            VT v = va[o];
            const VT *xr = X + ci[o] * k;
            for (int i = 0; i < k; ++i) acc[i] += v * xr[i];
        }
        for (int i = 0; i < k; ++i) Y[r * k + i] = acc[i];
    }
}

/* ------------------------------------------------------------------ */
/* Single-RHS CG building blocks                                        */
/* restate work_2025/main/single_strategy.hpp:29-55 (row-parallel SpMV),*/
/* :61-70 (dot), :76-83 (axpy), :90-97 (p update)                       */
/* ------------------------------------------------------------------ */
static void FN(orc_row_spmv)(int m, const int *ro, const int *ci, const VT *va,
                             const VT *x, VT *y)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < m; ++r) {
        VT s = 0.0;
        for (int o = ro[r]; o < ro[r + 1]; ++o) s += va[o] * x[ci[o]];
        y[r] = s;
    }
}

static VT FN(orc_dot)(int n, const VT *a, const VT *b)
{
    VT s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static void FN(orc_axpy)(int n, VT a, const VT *x, VT *y)
{
#pragma omp parallel for
    for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

static void FN(orc_update_p)(int n, const VT *r, VT beta, VT *p)
{
#pragma omp parallel for
    for (int i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
}

/* ------------------------------------------------------------------ */
/* Single-RHS CG -- restates single_strategy.hpp:105-170                */
/*  x0 = 0, r = p = b; test sqrt(r.r)/||b|| < tol BEFORE the p update;  */
/*  returned count includes the converging iteration (:152-156);        */
/*  ||b|| == 0 is replaced by 1 (:130-131).                             */
/* ------------------------------------------------------------------ */
int FN(orc_cg_single)(int n, const int *ro, const int *ci, const VT *va,
                      const VT *b, VT *x, int max_iters, VT tol)
{
    VT *r = (VT *)malloc(sizeof(VT) * (size_t)n);
    VT *p = (VT *)malloc(sizeof(VT) * (size_t)n);
    VT *Ap = (VT *)malloc(sizeof(VT) * (size_t)n);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }

    VT rs_old = FN(orc_dot)(n, r, r);
    VT b_norm = (VT)sqrt((double)FN(orc_dot)(n, b, b));
    if (b_norm == 0.0) b_norm = 1.0;

    int it = 0;
    for (; it < max_iters; ++it) {
        FN(orc_row_spmv)(n, ro, ci, va, p, Ap);
        VT pAp = FN(orc_dot)(n, p, Ap);
        VT alpha = rs_old / pAp;
        FN(orc_axpy)(n, alpha, p, x);
        FN(orc_axpy)(n, -alpha, Ap, r);
        VT rs_new = FN(orc_dot)(n, r, r);
        if ((VT)sqrt((double)rs_new) / b_norm < tol) { ++it; break; }
        VT beta = rs_new / rs_old;
        FN(orc_update_p)(n, r, beta, p);
        rs_old = rs_new;
    }
    free(r); free(p); free(Ap);
    return it;
}

/* ------------------------------------------------------------------ */
/* Multi-RHS helpers -- restate work_2025/cg/utils_multiple.hpp:9-24,   */
/* :28-42, :45-59 (row-major n x k blocks, per-column scalars)          */
/* ------------------------------------------------------------------ */
static void FN(orc_dot_multi)(int n, int k, const VT *X, const VT *Y, VT *out)
{
    int nt = omp_get_max_threads();
    VT *part = (VT *)calloc((size_t)nt * (size_t)k, sizeof(VT));
#pragma omp parallel
    {
        VT *mine = part + (size_t)omp_get_thread_num() * k;
#pragma omp for schedule(static)
        for (int j = 0; j < n; ++j) {
            const VT *xr = X + (long long)j * k;
            const VT *yr = Y + (long long)j * k;
            for (int i = 0; i < k; ++i) mine[i] += xr[i] * yr[i];
        }
    }
    for (int i = 0; i < k; ++i) out[i] = 0.0;
    for (int t = 0; t < nt; ++t)
        for (int i = 0; i < k; ++i) out[i] += part[(size_t)t * k + i];
    free(part);
}

static void FN(orc_axpy_multi)(int n, int k, const VT *a, const VT *X, VT *Y)
{
#pragma omp parallel for
    for (int j = 0; j < n; ++j) {
        const VT *xr = X + (long long)j * k;
        VT *yr = Y + (long long)j * k;
        for (int i = 0; i < k; ++i) yr[i] += a[i] * xr[i];
    }
}

static void FN(orc_update_p_multi)(int n, int k, const VT *R, const VT *beta, VT *P)
{
#pragma omp parallel for
    for (int j = 0; j < n; ++j) {
        const VT *rr = R + (long long)j * k;
        VT *pr = P + (long long)j * k;
        for (int i = 0; i < k; ++i) pr[i] = rr[i] + beta[i] * pr[i];
    }
}

/* ------------------------------------------------------------------ */
/* Multi-RHS CG -- restates work_2025/main/no_pretreatment.hpp:35-197   */
/*  k independent recurrences in lock-step; converged[i] latches and    */
/*  forces alpha=beta=0 (:109-120, :165-176); stops when all latched;   */
/*  max relative error per iteration is recorded (:133-155).            */
/*  kernel: 0 SIMPLE(row split) 1 MERGE 2 NONZERO_SPLIT (types.hpp:11)  */
/*  T = g_omp_threads of the reference (hyper_parameters.hpp:11).       */
/*  hist (nullable) receives up to max_iters doubles, *hist_len count.  */
/* ------------------------------------------------------------------ */
int FN(orc_cg_multi)(int T, int n, int nnz, const int *ro, const int *ci, const VT *va,
                     const VT *B, VT *X, int k, int max_iters, VT tol, int kernel,
                     double *hist, int *hist_len)
{
    size_t nk = (size_t)n * (size_t)k;
    VT *R = (VT *)malloc(sizeof(VT) * nk);
    VT *P = (VT *)malloc(sizeof(VT) * nk);
    VT *AP = (VT *)malloc(sizeof(VT) * nk);
    VT *alpha = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *beta = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *rs_old = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *rs_new = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *pAp = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *bn = (VT *)malloc(sizeof(VT) * (size_t)k);
    char *done = (char *)calloc((size_t)k, 1);

#pragma omp parallel for
    for (long long i = 0; i < (long long)nk; ++i) { X[i] = 0.0; R[i] = B[i]; P[i] = B[i]; }

    FN(orc_dot_multi)(n, k, B, B, bn);
    for (int i = 0; i < k; ++i) {
        bn[i] = (VT)sqrt((double)bn[i]);
        if (bn[i] == 0.0) bn[i] = 1.0;
    }
    FN(orc_dot_multi)(n, k, R, R, rs_old);
    int nh = 0;

    int it;
    for (it = 0; it < max_iters; ++it) {
        memset(AP, 0, sizeof(VT) * nk); /* no_pretreatment.hpp:93 */
        if (kernel == 0)
            FN(orc_row_split_csrmm)(T, n, ro, ci, va, P, AP, k);
        else if (kernel == 1)
            FN(orc_merge_csrmm)(T, n, nnz, ro + 1, ci, va, P, AP, k);
        else
            FN(orc_nonzero_split_csrmm)(T, n, nnz, ro + 1, ci, va, P, AP, k);

        FN(orc_dot_multi)(n, k, P, AP, pAp);
        for (int i = 0; i < k; ++i) alpha[i] = done[i] ? (VT)0.0 : rs_old[i] / pAp[i];
        FN(orc_axpy_multi)(n, k, alpha, P, X);
        for (int i = 0; i < k; ++i) alpha[i] = -alpha[i];
        FN(orc_axpy_multi)(n, k, alpha, AP, R);
        FN(orc_dot_multi)(n, k, R, R, rs_new);

        int ndone = 0;
        double worst = 0.0;
        for (int i = 0; i < k; ++i) {
            double rel = sqrt((double)rs_new[i]) / (double)bn[i];
            if (rel > worst) worst = rel;
            if (!done[i] && rel < (double)tol) done[i] = 1;
            if (done[i]) ++ndone;
        }
        if (hist) hist[nh] = worst;
        ++nh;
        if (ndone == k) { ++it; break; }

        for (int i = 0; i < k; ++i) beta[i] = done[i] ? (VT)0.0 : rs_new[i] / rs_old[i];
        FN(orc_update_p_multi)(n, k, R, beta, P);
        for (int i = 0; i < k; ++i) rs_old[i] = rs_new[i];
    }
    if (hist_len) *hist_len = nh;
    free(R); free(P); free(AP); free(alpha); free(beta); free(rs_old); free(rs_new);
    free(pAp); free(bn); free(done);
    return it;
}

/* ------------------------------------------------------------------ */
/* SPAI preconditioner                        */
/* orc_spai_build restates SparseApproximateInversion                    */
/*  (work_2025/cg/sparse_approximate_inversion.hpp:41-321): static       */
/*  pattern S_M = S_A; for every column k, J = rows of A's column k,     */
/*  I = union of the rows of A's columns j in J (in first-seen order),   */
/*  minimise || A(I,J) m - e_k(I) ||_2 (:131-222; the reference calls    */
/*  LAPACKE_?gels = Householder QR, restated in orc_lstsq), write m into */
/*  M(J,k) through the CSC->CSR map (:226-238; zeros when the solve      */
/*  fails :241-247), then symmetrise M = (M + M^T)/2 over the upper      */
/*  triangle (:268-318).  Returns 0.                                     */
/* ------------------------------------------------------------------ */
static int FN(orc_lstsq)(int m, int n, VT *a, VT *b)
{
    /* row-major m x n (lda = n), m >= n; solution in b[0:n]; > 0: rank deficient */
    if (m < n) return n + 1;
    for (int j = 0; j < n; ++j) {
        double nrm = 0.0;
        for (int i = j; i < m; ++i) nrm += (double)a[i * n + j] * (double)a[i * n + j];
        nrm = sqrt(nrm);
        if (nrm == 0.0) return j + 1;
        double ajj = (double)a[j * n + j], alpha = ajj > 0.0 ? -nrm : nrm, v0 = ajj - alpha;
        double vv = v0 * v0;
        for (int i = j + 1; i < m; ++i) vv += (double)a[i * n + j] * (double)a[i * n + j];
        if (vv > 0.0) {
            for (int c = j + 1; c <= n; ++c) {   /* c == n: the right-hand side */
                double dot = v0 * (double)(c < n ? a[j * n + c] : b[j]);
                for (int i = j + 1; i < m; ++i) dot += (double)a[i * n + j] * (double)(c < n ? a[i * n + c] : b[i]);
                double f = 2.0 * dot / vv;
                if (c < n) {
                    a[j * n + c] = (VT)((double)a[j * n + c] - f * v0);
                    for (int i = j + 1; i < m; ++i) a[i * n + c] = (VT)((double)a[i * n + c] - f * (double)a[i * n + j]);
                } else {
                    b[j] = (VT)((double)b[j] - f * v0);
                    for (int i = j + 1; i < m; ++i) b[i] = (VT)((double)b[i] - f * (double)a[i * n + j]);
                }
            }
        }
        a[j * n + j] = (VT)alpha;
    }
    for (int j = n - 1; j >= 0; --j) {
        double sacc = (double)b[j];
        for (int c = j + 1; c < n; ++c) sacc -= (double)a[j * n + c] * (double)b[c];
        b[j] = (VT)(sacc / (double)a[j * n + j]);
    }
    return 0;
}

int FN(orc_spai_build)(int m, int nnz, const int *ro, const int *ci, const VT *va, VT *mv)
{
    /* CSC of A with the map back to CSR positions (:88-118) */
    int *co = (int *)calloc((size_t)m + 1, sizeof(int));
    int *cr = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    int *map = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    VT *cv = (VT *)malloc(sizeof(VT) * (size_t)(nnz ? nnz : 1));
    for (int z = 0; z < nnz; ++z) ++co[ci[z] + 1];
    for (int c = 0; c < m; ++c) co[c + 1] += co[c];
    {
        int *pos = (int *)malloc(sizeof(int) * ((size_t)m + 1));
        memcpy(pos, co, sizeof(int) * ((size_t)m + 1));
        for (int r = 0; r < m; ++r)
            for (int z = ro[r]; z < ro[r + 1]; ++z) {
                int d = pos[ci[z]]++;
                cr[d] = r; cv[d] = va[z]; map[d] = z;
            }
        free(pos);
    }
    for (int z = 0; z < nnz; ++z) mv[z] = 0;
#pragma omp parallel
    {
        int *g2l = (int *)malloc(sizeof(int) * (size_t)(m ? m : 1));
        for (int i = 0; i < m; ++i) g2l[i] = -1;
        int cap_rows = 64, cap_dense = 1024;
        int *rows = (int *)malloc(sizeof(int) * (size_t)cap_rows);
        VT *dense = (VT *)malloc(sizeof(VT) * (size_t)cap_dense);
        VT *rhs = (VT *)malloc(sizeof(VT) * (size_t)cap_rows);
#pragma omp for schedule(static)
        for (int k = 0; k < m; ++k) {
            int j0 = co[k], nv = co[k + 1] - j0, ne = 0;
            if (nv == 0) continue;
            for (int q = j0; q < j0 + nv; ++q) {
                int col = cr[q];
                for (int t = co[col]; t < co[col + 1]; ++t) {
                    int r = cr[t];
                    if (g2l[r] == -1) {
                        if (ne == cap_rows) {
                            cap_rows *= 2;
                            rows = (int *)realloc(rows, sizeof(int) * (size_t)cap_rows);
                            rhs = (VT *)realloc(rhs, sizeof(VT) * (size_t)cap_rows);
                        }
                        g2l[r] = ne; rows[ne++] = r;
                    }
                }
            }
            if (ne * nv > cap_dense) { cap_dense = ne * nv; dense = (VT *)realloc(dense, sizeof(VT) * (size_t)cap_dense); }
            for (int i = 0; i < ne * nv; ++i) dense[i] = 0;
            for (int i = 0; i < ne; ++i) rhs[i] = 0;
            if (g2l[k] != -1) rhs[g2l[k]] = 1;
            for (int jl = 0; jl < nv; ++jl) {
                int col = cr[j0 + jl];
                for (int t = co[col]; t < co[col + 1]; ++t) dense[g2l[cr[t]] * nv + jl] = cv[t];
            }
            int info = FN(orc_lstsq)(ne, nv, dense, rhs);
            for (int jl = 0; jl < nv; ++jl) mv[map[j0 + jl]] = info == 0 ? rhs[jl] : (VT)0;
            for (int i = 0; i < ne; ++i) g2l[rows[i]] = -1;
        }
        free(g2l); free(rows); free(dense); free(rhs);
    }
    /* symmetrise over the upper triangle (:268-318) */
#pragma omp parallel for schedule(static)
    for (int r = 0; r < m; ++r)
        for (int z = ro[r]; z < ro[r + 1]; ++z) {
            int c = ci[z];
            if (c <= r) continue;
            for (int t = ro[c]; t < ro[c + 1]; ++t)
                if (ci[t] == r) {
                    VT avg = (mv[z] + mv[t]) * (VT)0.5;
                    mv[z] = avg; mv[t] = avg;
                    break;
                }
        }
    free(co); free(cr); free(map); free(cv);
    return 0;
}

/* ------------------------------------------------------------------ */
/* SPAI-preconditioned multi-RHS CG -- restates SPAISolveMultiple        */
/*  (work_2025/main/sparse_approximate_inverse.hpp:31-230): z = M r by   */
/*  an SpMM with M, rs = r.z, alpha guarded against pAp == 0 (:121-127), */
/*  beta against rs_old == 0 (:186-192), convergence on sqrt(r.r)/||b||  */
/*  before the M step (:139-166).                                        */
/* ------------------------------------------------------------------ */
static void FN(orc_spmm_kernel)(int T, int kernel, int n, int nnz, const int *ro, const int *ci, const VT *va,
                                const VT *X, VT *Y, int k)
{
    if (kernel == 0) FN(orc_row_split_csrmm)(T, n, ro, ci, va, X, Y, k);
    else if (kernel == 1) FN(orc_merge_csrmm)(T, n, nnz, ro + 1, ci, va, X, Y, k);
    else FN(orc_nonzero_split_csrmm)(T, n, nnz, ro + 1, ci, va, X, Y, k);
}

int FN(orc_spai_solve_multi)(int T, int n, int nnz, const int *ro, const int *ci, const VT *va, const VT *mv,
                             const VT *B, VT *X, int k, int max_iters, VT tol, int kernel,
                             double *hist, int *hist_len)
{
    size_t nk = (size_t)n * (size_t)k;
    VT *R = (VT *)malloc(sizeof(VT) * nk), *P = (VT *)malloc(sizeof(VT) * nk);
    VT *AP = (VT *)malloc(sizeof(VT) * nk), *Z = (VT *)malloc(sizeof(VT) * nk);
    VT *alpha = (VT *)malloc(sizeof(VT) * (size_t)k), *beta = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *rs_old = (VT *)malloc(sizeof(VT) * (size_t)k), *rs_new = (VT *)malloc(sizeof(VT) * (size_t)k);
    VT *pAp = (VT *)malloc(sizeof(VT) * (size_t)k), *bn = (VT *)malloc(sizeof(VT) * (size_t)k);
    char *done = (char *)calloc((size_t)k, 1);
#pragma omp parallel for
    for (long long i = 0; i < (long long)nk; ++i) { X[i] = 0.0; R[i] = B[i]; P[i] = 0.0; Z[i] = 0.0; }
    FN(orc_dot_multi)(n, k, B, B, bn);
    for (int i = 0; i < k; ++i) {
        bn[i] = (VT)sqrt((double)bn[i]);
        if (bn[i] == 0.0) bn[i] = 1.0;
    }
    FN(orc_spmm_kernel)(T, kernel, n, nnz, ro, ci, mv, R, Z, k);
    memcpy(P, Z, sizeof(VT) * nk);
    FN(orc_dot_multi)(n, k, R, Z, rs_old);
    int nh = 0, it;
    for (it = 0; it < max_iters; ++it) {
        FN(orc_spmm_kernel)(T, kernel, n, nnz, ro, ci, va, P, AP, k);
        FN(orc_dot_multi)(n, k, P, AP, pAp);
        for (int i = 0; i < k; ++i) alpha[i] = (!done[i] && pAp[i] != 0.0) ? rs_old[i] / pAp[i] : (VT)0.0;
        FN(orc_axpy_multi)(n, k, alpha, P, X);
        for (int i = 0; i < k; ++i) alpha[i] = -alpha[i];
        FN(orc_axpy_multi)(n, k, alpha, AP, R);
        FN(orc_dot_multi)(n, k, R, R, pAp);
        int ndone = 0;
        double worst = 0.0;
        for (int i = 0; i < k; ++i) {
            double rel = sqrt((double)pAp[i]) / (double)bn[i];
            if (rel > worst) worst = rel;
            if (!done[i] && rel < (double)tol) done[i] = 1;
            if (done[i]) ++ndone;
        }
        if (hist) hist[nh] = worst;
        ++nh;
        if (ndone == k) { ++it; break; }
        FN(orc_spmm_kernel)(T, kernel, n, nnz, ro, ci, mv, R, Z, k);
        FN(orc_dot_multi)(n, k, R, Z, rs_new);
        for (int i = 0; i < k; ++i) {
            beta[i] = (!done[i] && rs_old[i] != 0.0) ? rs_new[i] / rs_old[i] : (VT)0.0;
            rs_old[i] = rs_new[i];
        }
        FN(orc_update_p_multi)(n, k, Z, beta, P);
    }
    if (hist_len) *hist_len = nh;
    free(R); free(P); free(AP); free(Z); free(alpha); free(beta); free(rs_old); free(rs_new);
    free(pAp); free(bn); free(done);
    return it;
}

/* ------------------------------------------------------------------ */
/* COO -> CSR, as CsrMatrix::Init (sparse_matrix.h:668-733):            */
/* stable sort by (row, col), duplicates kept, then offsets fill.       */
/* Input COO arrays are consumed (sorted in place via a scratch copy).  */
/* ------------------------------------------------------------------ */
typedef struct { int row, col; VT val; } FN(orc_coo);

static void FN(orc_coo_stable_sort)(FN(orc_coo) *t, size_t n)
{
    /* bottom-up merge sort: stable, same ordering as std::stable_sort with
       CooComparator (sparse_matrix.h:636-643) */
    FN(orc_coo) *tmp = (FN(orc_coo) *)malloc(sizeof(*tmp) * (n ? n : 1));
    FN(orc_coo) *src = t, *dst = tmp;
    for (size_t w = 1; w < n; w <<= 1) {
#pragma omp parallel for schedule(dynamic, 1)
        for (long long lo0 = 0; lo0 < (long long)n; lo0 += (long long)(2 * w)) {
            size_t lo = (size_t)lo0;
            size_t mid = lo + w < n ? lo + w : n;
            size_t hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                int take_right = (src[j].row < src[i].row) ||
                                 (src[j].row == src[i].row && src[j].col < src[i].col);
                dst[o++] = take_right ? src[j++] : src[i++];
            }
            while (i < mid) dst[o++] = src[i++];
            while (j < hi) dst[o++] = src[j++];
        }
        FN(orc_coo) *sw = src; src = dst; dst = sw;
    }
    if (src != t) memcpy(t, src, sizeof(*t) * n);
    free(tmp);
}

static void FN(orc_coo_to_csr)(FN(orc_coo) *t, int m, int nnz, int *ro, int *ci, VT *va)
{
    FN(orc_coo_stable_sort)(t, (size_t)nnz);
    int prev = -1;
    for (int z = 0; z < nnz; ++z) {
        int r = t[z].row;
        for (int q = prev + 1; q <= r; ++q) ro[q] = z;
        prev = r;
        ci[z] = t[z].col;
        va[z] = t[z].val;
    }
    for (int q = prev + 1; q <= m; ++q) ro[q] = nnz;
}

/* Generators: each returns CSR through caller buffers sized by the
   matching orc_gen_*_shape().  diag/offd let the caller apply the Poisson
   fill described in SURVEY.md Appendix B (value chosen per tuple BEFORE
   the sort); diag == offd == 1.0 reproduces the reference default. */

/* InitGrid2d -- sparse_matrix.h:458-527 (neighbour order W,E,N,S,self) */
void FN(orc_gen_grid2d)(int w, int self_loop, VT diag, VT offd, int *ro, int *ci, VT *va)
{
    int m, n, nnz;
    orc_gen_grid2d_shape(w, self_loop, &m, &n, &nnz);
    FN(orc_coo) *t = (FN(orc_coo) *)malloc(sizeof(*t) * (size_t)(nnz ? nnz : 1));
    int z = 0;
    for (int j = 0; j < w; ++j)
        for (int k = 0; k < w; ++k) {
            int me = j * w + k;
            if (k - 1 >= 0) { t[z].row = me; t[z].col = j * w + (k - 1); t[z].val = offd; ++z; }
            if (k + 1 < w)  { t[z].row = me; t[z].col = j * w + (k + 1); t[z].val = offd; ++z; }
            if (j - 1 >= 0) { t[z].row = me; t[z].col = (j - 1) * w + k; t[z].val = offd; ++z; }
            if (j + 1 < w)  { t[z].row = me; t[z].col = (j + 1) * w + k; t[z].val = offd; ++z; }
            if (self_loop)  { t[z].row = me; t[z].col = me; t[z].val = diag; ++z; }
        }
    FN(orc_coo_to_csr)(t, m, nnz, ro, ci, va);
    free(t);
}

/* InitGrid3d -- sparse_matrix.h:533-623 (order -k,+k,-j,+j,-i,+i,self) */
void FN(orc_gen_grid3d)(int w, int self_loop, VT diag, VT offd, int *ro, int *ci, VT *va)
{
    int m, n, nnz;
    orc_gen_grid3d_shape(w, self_loop, &m, &n, &nnz);
    FN(orc_coo) *t = (FN(orc_coo) *)malloc(sizeof(*t) * (size_t)(nnz ? nnz : 1));
    int z = 0, ww = w * w;
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < w; ++j)
            for (int k = 0; k < w; ++k) {
                int me = i * ww + j * w + k;
                if (k - 1 >= 0) { t[z].row = me; t[z].col = me - 1;  t[z].val = offd; ++z; }
                if (k + 1 < w)  { t[z].row = me; t[z].col = me + 1;  t[z].val = offd; ++z; }
                if (j - 1 >= 0) { t[z].row = me; t[z].col = me - w;  t[z].val = offd; ++z; }
                if (j + 1 < w)  { t[z].row = me; t[z].col = me + w;  t[z].val = offd; ++z; }
                if (i - 1 >= 0) { t[z].row = me; t[z].col = me - ww; t[z].val = offd; ++z; }
                if (i + 1 < w)  { t[z].row = me; t[z].col = me + ww; t[z].val = offd; ++z; }
                if (self_loop)  { t[z].row = me; t[z].col = me;      t[z].val = diag; ++z; }
            }
    FN(orc_coo_to_csr)(t, m, nnz, ro, ci, va);
    free(t);
}

/* The same CSR as orc_gen_grid3d, written directly in sorted order (me-w^2, me-w, me-1, [me],
   me+1, me+w, me+w^2) and in parallel: the bench's reference arm needs the 300^3 system
   (188 M nonzeros) in seconds, the COO sort above takes minutes there.  tests/ compare the two
   on small widths. */
void FN(orc_gen_grid3d_sorted)(int w, int self_loop, VT diag, VT offd, int *ro, int *ci, VT *va)
{
    int ww = w * w, m = ww * w;
    ro[0] = 0;
#pragma omp parallel for schedule(static)
    for (int me = 0; me < m; ++me) {
        int i = me / ww, j = (me / w) % w, k = me % w;
        ro[me + 1] = (i > 0) + (j > 0) + (k > 0) + (k + 1 < w) + (j + 1 < w) + (i + 1 < w) + (self_loop ? 1 : 0);
    }
    for (int r = 0; r < m; ++r) ro[r + 1] += ro[r];
#pragma omp parallel for schedule(static)
    for (int me = 0; me < m; ++me) {
        int i = me / ww, j = (me / w) % w, k = me % w, z = ro[me];
        if (i > 0)     { ci[z] = me - ww; va[z++] = offd; }
        if (j > 0)     { ci[z] = me - w;  va[z++] = offd; }
        if (k > 0)     { ci[z] = me - 1;  va[z++] = offd; }
        if (self_loop) { ci[z] = me;      va[z++] = diag; }
        if (k + 1 < w) { ci[z] = me + 1;  va[z++] = offd; }
        if (j + 1 < w) { ci[z] = me + w;  va[z++] = offd; }
        if (i + 1 < w) { ci[z] = me + ww; va[z++] = offd; }
    }
}

/* InitWheel -- sparse_matrix.h:417-450: hub row 0 -> 1..s, rim i+1 -> ((i+1)%s)+1 */
void FN(orc_gen_wheel)(int spokes, VT value, int *ro, int *ci, VT *va)
{
    int m = spokes + 1, nnz = 2 * spokes;
    FN(orc_coo) *t = (FN(orc_coo) *)malloc(sizeof(*t) * (size_t)(nnz ? nnz : 1));
    int z = 0;
    for (int i = 0; i < spokes; ++i) { t[z].row = 0; t[z].col = i + 1; t[z].val = value; ++z; }
    for (int i = 0; i < spokes; ++i) {
        t[z].row = i + 1; t[z].col = ((i + 1) % spokes) + 1; t[z].val = value; ++z;
    }
    FN(orc_coo_to_csr)(t, m, nnz, ro, ci, va);
    free(t);
}

/* InitDense -- sparse_matrix.h:385-412 */
void FN(orc_gen_dense)(int rows, int cols, VT value, int *ro, int *ci, VT *va)
{
    int nnz = rows * cols;
    FN(orc_coo) *t = (FN(orc_coo) *)malloc(sizeof(*t) * (size_t)(nnz ? nnz : 1));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            t[r * cols + c].row = r; t[r * cols + c].col = c; t[r * cols + c].val = value;
        }
    FN(orc_coo_to_csr)(t, rows, nnz, ro, ci, va);
    free(t);
}

/* RHS as the CG drivers make it: srand(seed); b[i] = rand()/RAND_MAX
   (cpu_singlecg.cpp:88-90, cpu_multicg.cpp:164-166; glibc rand) */
void FN(orc_rhs_rand)(unsigned seed, long long count, VT *out)
{
    srand(seed);
    for (long long i = 0; i < count; ++i) out[i] = (VT)rand() / (VT)RAND_MAX;
}

/* driver threshold quirk: ||b[0:n]||_2 * tol  (cpu_singlecg.cpp:23-34) */
VT FN(orc_driver_threshold)(const VT *b, int n, VT tol)
{
    VT s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (int i = 0; i < n; ++i) s += b[i] * b[i];
    return (VT)sqrt((double)s) * tol;
}

#undef FN
#undef ORC_CAT
#undef ORC_CAT_
