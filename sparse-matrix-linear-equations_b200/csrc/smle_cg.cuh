// smle_cg.cuh -- fused CG vector kernels for sm_100a (fp64).
//
// One CG iteration of the reference (single_strategy.hpp:134-163 / no_pretreatment.hpp:90-182)
// is SpMM + 3 dot sweeps + 2 axpy sweeps + 1 p-update sweep (+ a memset), i.e. ~14 passes over
// the n x k blocks.  Here it is three kernels and 10 passes, with every scalar on the device:
//
//   K1  merge_kernel<DOT>     AP = A P, pAp partials while rows are emitted; last CTA: alpha
//   K2  cg_update_r_kernel    R -= alpha AP, r.r partials; last CTA: rs_new, convergence
//                             latch, beta, error history, iteration counter, stop flag
//   K3  cg_update_xp_kernel   X += alpha P;  P = R + beta P
//
// The arithmetic per element is the reference's (same operations on the same operands); only
// the sweep boundaries moved.  Per-column dot products are reduced deterministically: each CTA
// writes one partial per column and the last CTA to finish adds them in CTA order.
//
// Thread mapping (shared with the SpMM kernel): a "worker" of G lanes covers KB = G*VEC
// consecutive columns of one row with VEC-wide (up to 128-bit) accesses; workers stride over
// rows, column blocks of KB columns are looped inside the kernel.
#pragma once
#include "smle_common.cuh"

namespace smle {

// V = value type of the blocks (double: the reference's solvers; float: the fp32 variant).  Scalars, dot
// partials and their reductions are double for both.
template <typename V>
struct CgVecArgsT {
    const V *__restrict__ B;
    V *__restrict__ X;
    V *__restrict__ R;
    V *__restrict__ P;
    V *__restrict__ AP;
    int n, k;
    double *part;          // [gridDim.x * k] per-CTA partials
    unsigned int *ticket;
    V *__restrict__ Z = nullptr;   // preconditioned residual z = M r (SPAI-PCG only)
};
using CgVecArgs = CgVecArgsT<double>;

// reduce VEC per-lane partials over the workers of a CTA and publish them for this CTA
template <int G, int VEC>
__device__ __forceinline__ void publish_partials(double (&s)[VEC], int cb, int k, double *part,
                                                 double (*s_w)[G * VEC])
{
    constexpr int KB = G * VEC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = tid % G, wl = lane / G;
#pragma unroll
    for (int d = G; d < 32; d <<= 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] += __shfl_xor_sync(0xffffffffu, s[v], d);
    }
    __syncthreads();
    if (wl == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) s_w[warp][li * VEC + v] = s[v];
    }
    __syncthreads();
    if (tid < KB) {
        double t = 0;
        for (int i = 0; i < kWarps; ++i) t += s_w[i][tid];
        int c = cb * KB + tid;
        if (c < k) part[(size_t)blockIdx.x * k + c] = t;
    }
}

// ---------------------------------------------------------------------------------------
// cg_init_kernel: X = 0, R = P = B, b.b partials  (no_pretreatment.hpp:63-81,
// single_strategy.hpp:120-131).  Last CTA: bnorm = sqrt(b.b) (0 -> 1), rs_old = b.b,
// latches cleared, control words reset.
// ---------------------------------------------------------------------------------------
template <typename V, int G, int VEC>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(CgVecArgsT<V> a, CgScalars cg, int max_iters, double tol, int seq_base)
{
    constexpr int W = kThreads / G, KB = G * VEC;
    __shared__ double s_w[kWarps][KB];
    __shared__ double s_red[kThreads];
    const int tid = threadIdx.x, w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;

    for (int cb = 0; cb * KB < a.k; ++cb) {
        const int c0 = cb * KB + li * VEC;
        double s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = 0;
        if (c0 < a.k) {
            for (int row = blockIdx.x * W + w; row < a.n; row += gridDim.x * W) {
                size_t off = (size_t)row * k + c0;
                V b[VEC], z[VEC];
                ldg_vec<V, VEC>(b, a.B + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { z[v] = 0; s[v] += (double)b[v] * (double)b[v]; }
                st_vec<V, VEC>(a.X + off, z);
                st_vec<V, VEC>(a.R + off, b);
                st_vec<V, VEC>(a.P + off, b);
            }
        }
        publish_partials<G, VEC>(s, cb, a.k, a.part, s_w);
    }

    if (!last_cta_election(a.ticket, gridDim.x)) return;
    cta_reduce_columns<double>(a.part, nullptr, gridDim.x, a.k, cg.rs_old, s_red);
    for (int c = tid; c < a.k; c += kThreads) {
        double nb = sqrt(cg.rs_old[c]);
        cg.bnorm[c] = nb == 0.0 ? 1.0 : nb;
        cg.conv[c] = 0;
    }
    if (tid == 0) {
        cg.ctrl[CTRL_ITER] = 0;
        cg.ctrl[CTRL_STOP] = max_iters <= 0 ? 1 : 0;
        cg.ctrl[CTRL_HALT] = max_iters <= 0 ? 1 : 0;
        cg.ctrl[CTRL_MAX_ITERS] = max_iters;
        cg.ctrl[CTRL_NCONV] = 0;
        cg.ctrl[CTRL_SEQ_BASE] = seq_base;
        *cg.last_rel = 0.0;
        *cg.tol = tol;
    }
}

// ---------------------------------------------------------------------------------------
// Epilogue of K2, run by the last CTA to finish: rs_new from the per-CTA partials, per-column
// relative residual and latch (no_pretreatment.hpp:133-155), beta (:165-176), rs_old <- rs_new
// (:179-181), error history, iteration count, stop flag (:157-161 or max_iters).
// ---------------------------------------------------------------------------------------
template <typename Args>
__device__ __forceinline__ void cg_finalize_r(const Args &a, const CgScalars &cg, double *s_red, int *s_cnt, bool pcg = false)
{
    const int tid = threadIdx.x;
    if (a.k == 1) {
        // single right-hand side: shuffle-tree reduction, bookkeeping by one thread (the generic path
        // below spends ~3 us of every iteration in block-wide trees over a single value)
        const double rn = cta_reduce_one<double>(a.part, nullptr, gridDim.x, s_red);
        if (tid == 0) {
            const double ro = cg.rs_old[0];
            const double rel = sqrt(rn) / cg.bnorm[0];
            int cv = cg.conv[0];
            if (!cv && rel < *cg.tol) { cv = 1; cg.conv[0] = 1; }
            cg.rs_new[0] = rn;
            if (!pcg) {   // (PCG: beta and rs_old come from r.z, formed by the M step)
                cg.beta[0] = cv ? 0.0 : rn / ro;
                cg.rs_old[0] = rn;
            }
            const int it = cg.ctrl[CTRL_ITER];
            if (cg.hist && it < cg.hist_cap) cg.hist[it] = rel;
            *cg.last_rel = rel;
            cg.ctrl[CTRL_ITER] = it + 1;
            cg.ctrl[CTRL_NCONV] = cv;
            if (cv || it + 1 >= cg.ctrl[CTRL_MAX_ITERS]) cg.ctrl[CTRL_STOP] = 1;
        }
        return;
    }
    cta_reduce_columns<double>(a.part, nullptr, gridDim.x, a.k, cg.rs_new, s_red);

    double worst = 0.0;
    int nconv = 0;
    const double tol = *cg.tol;
    for (int c = tid; c < a.k; c += kThreads) {
        const double rn = cg.rs_new[c], ro = cg.rs_old[c];
        const double rel = sqrt(rn) / cg.bnorm[c];
        worst = fmax(worst, rel);
        int cv = cg.conv[c];
        if (!cv && rel < tol) { cv = 1; cg.conv[c] = 1; }
        nconv += cv;
        if (!pcg) {
            cg.beta[c] = cv ? 0.0 : rn / ro;
            cg.rs_old[c] = rn;
        }
    }
    s_red[tid] = worst;
    s_cnt[tid] = nconv;
    __syncthreads();
    for (int d = kThreads / 2; d > 0; d >>= 1) {
        if (tid < d) {
            s_red[tid] = fmax(s_red[tid], s_red[tid + d]);
            s_cnt[tid] += s_cnt[tid + d];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int it = cg.ctrl[CTRL_ITER];
        if (cg.hist && it < cg.hist_cap) cg.hist[it] = s_red[0];
        *cg.last_rel = s_red[0];
        cg.ctrl[CTRL_ITER] = it + 1;
        cg.ctrl[CTRL_NCONV] = s_cnt[0];
        if (s_cnt[0] == a.k || it + 1 >= cg.ctrl[CTRL_MAX_ITERS]) cg.ctrl[CTRL_STOP] = 1;
    }
}

// ---------------------------------------------------------------------------------------
// K2: R -= alpha * AP and r.r  (no_pretreatment.hpp:125-130; single_strategy.hpp:147-149).
// Last CTA: rs_new, per-column relative residual and latch (:133-155), beta (:165-176),
// rs_old <- rs_new (:179-181), history, iteration count, stop flag (:157-161 or max_iters).
// ---------------------------------------------------------------------------------------
template <typename V, int G, int VEC>
__global__ void __launch_bounds__(kThreads)
cg_update_r_kernel(CgVecArgsT<V> a, CgScalars cg)
{
    constexpr int W = kThreads / G, KB = G * VEC;
    __shared__ double s_w[kWarps][KB];
    __shared__ double s_red[kThreads];
    __shared__ int s_cnt[kThreads];
    griddep_wait();
    griddep_launch_dependents();
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x, w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;

    for (int cb = 0; cb * KB < a.k; ++cb) {
        const int c0 = cb * KB + li * VEC;
        double s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = 0;
        if (c0 < a.k) {
            V na[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) na[v] = (V)(-cg.alpha[c0 + v]);
            for (int row = blockIdx.x * W + w; row < a.n; row += gridDim.x * W) {
                size_t off = (size_t)row * k + c0;
                V r[VEC], ap[VEC];
                ld_vec<V, VEC>(r, a.R + off);
                ld_vec<V, VEC>(ap, a.AP + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { r[v] += na[v] * ap[v]; s[v] += (double)r[v] * (double)r[v]; }
                st_vec<V, VEC>(a.R + off, r);
            }
        }
        publish_partials<G, VEC>(s, cb, a.k, a.part, s_w);
    }

    if (!last_cta_election(a.ticket, gridDim.x)) return;
    cg_finalize_r(a, cg, s_red, s_cnt);
}

// ---------------------------------------------------------------------------------------
// K3: X += alpha * P (no_pretreatment.hpp:123) and, unless this was the final iteration,
// P = R + beta * P (:177).  HALT is raised by the first K1 that sees STOP, i.e. after the
// final iteration's K3 has run.
// ---------------------------------------------------------------------------------------
template <typename V, int G, int VEC>
__global__ void __launch_bounds__(kThreads)
cg_update_xp_kernel(CgVecArgsT<V> a, CgScalars cg)
{
    constexpr int W = kThreads / G, KB = G * VEC;
    griddep_wait();
    griddep_launch_dependents();
    if (cg.ctrl[CTRL_HALT]) return;
    const bool final_iter = cg.ctrl[CTRL_STOP] != 0;
    const int tid = threadIdx.x, w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;

    for (int cb = 0; cb * KB < a.k; ++cb) {
        const int c0 = cb * KB + li * VEC;
        if (c0 >= a.k) continue;
        V al[VEC], be[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { al[v] = (V)cg.alpha[c0 + v]; be[v] = (V)cg.beta[c0 + v]; }
        for (int row = blockIdx.x * W + w; row < a.n; row += gridDim.x * W) {
            size_t off = (size_t)row * k + c0;
            V x[VEC], p[VEC];
            ld_vec<V, VEC>(x, a.X + off);
            ld_vec<V, VEC>(p, a.P + off);
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[v] += al[v] * p[v];
            st_vec<V, VEC>(a.X + off, x);
            if (!final_iter) {
                V r[VEC];
                ld_vec<V, VEC>(r, a.R + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) p[v] = r[v] + be[v] * p[v];
                st_vec<V, VEC>(a.P + off, p);
            }
        }
    }
}

// =======================================================================================
// SPAI-preconditioned CG (SPAISolveMultiple, work_2025/main/sparse_approximate_inverse.hpp:31-230).
// One iteration = SpMM<DOT>(A, P -> AP, alpha) | pcg_update_xr | SpMM<DOT>(M, R -> Z, beta) |
// pcg_update_p: 2 sparse products + 9 block passes (reference: 2 products + ~16 passes).
// =======================================================================================
// X += alpha P; R -= alpha AP; r.r (:129-137).  Last CTA: relative residual, latch, history,
// iteration count, stop flag (:139-166) -- the convergence test precedes the M step.
template <int G, int VEC>
__global__ void __launch_bounds__(kThreads)
pcg_update_xr_kernel(CgVecArgs a, CgScalars cg)
{
    constexpr int W = kThreads / G, KB = G * VEC;
    __shared__ double s_w[kWarps][KB];
    __shared__ double s_red[kThreads];
    __shared__ int s_cnt[kThreads];
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x, w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;
    for (int cb = 0; cb * KB < a.k; ++cb) {
        const int c0 = cb * KB + li * VEC;
        double s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = 0;
        if (c0 < a.k) {
            double al[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) al[v] = cg.alpha[c0 + v];
            for (int row = blockIdx.x * W + w; row < a.n; row += gridDim.x * W) {
                size_t off = (size_t)row * k + c0;
                double x[VEC], p[VEC], r[VEC], ap[VEC];
                ld_vec<double, VEC>(x, a.X + off);
                ld_vec<double, VEC>(p, a.P + off);
                ld_vec<double, VEC>(r, a.R + off);
                ld_vec<double, VEC>(ap, a.AP + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    x[v] += al[v] * p[v];
                    r[v] += -al[v] * ap[v];
                    s[v] += r[v] * r[v];
                }
                st_vec<double, VEC>(a.X + off, x);
                st_vec<double, VEC>(a.R + off, r);
            }
        }
        publish_partials<G, VEC>(s, cb, a.k, a.part, s_w);
    }
    if (!last_cta_election(a.ticket, gridDim.x)) return;
    cg_finalize_r(a, cg, s_red, s_cnt, /*pcg=*/true);
}

// P = Z + beta P (:194).  Skipped once the stop flag is up (the reference breaks before the M step).
template <int G, int VEC>
__global__ void __launch_bounds__(kThreads)
pcg_update_p_kernel(CgVecArgs a, CgScalars cg)
{
    constexpr int W = kThreads / G, KB = G * VEC;
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x, w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;
    for (int cb = 0; cb * KB < a.k; ++cb) {
        const int c0 = cb * KB + li * VEC;
        if (c0 >= a.k) continue;
        double be[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) be[v] = cg.beta[c0 + v];
        for (int row = blockIdx.x * W + w; row < a.n; row += gridDim.x * W) {
            size_t off = (size_t)row * k + c0;
            double z[VEC], p[VEC];
            ld_vec<double, VEC>(z, a.Z + off);
            ld_vec<double, VEC>(p, a.P + off);
#pragma unroll
            for (int v = 0; v < VEC; ++v) p[v] = z[v] + be[v] * p[v];
            st_vec<double, VEC>(a.P + off, p);
        }
    }
}

// =======================================================================================
// Single right-hand side (k = 1) specialisations of K2 / K3: 128-bit accesses, four
// independent vectors in flight per thread, explicit L2 eviction priorities.
// =======================================================================================
constexpr int kVecUnroll = 4;

__global__ void __launch_bounds__(kThreads)
cg1_update_r_kernel(CgVecArgs a, CgScalars cg)
{
    __shared__ double s_red[kThreads];
    __shared__ int s_cnt[kThreads];
    griddep_wait();
    griddep_launch_dependents();
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    const double na = -cg.alpha[0];
    const long long n2 = a.n >> 1;
    const long long stride = (long long)gridDim.x * kThreads;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kThreads + tid; i < n2; i += stride * kVecUnroll) {
        double2 r[kVecUnroll], ap[kVecUnroll];
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < n2) {
                r[u] = ld_f64x2_hint(a.R + 2 * j, pol_last);
                ap[u] = ld_f64x2_hint(a.AP + 2 * j, pol_first);   // AP is dead after this read
            }
        }
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < n2) {
                r[u].x += na * ap[u].x;
                r[u].y += na * ap[u].y;
                s += r[u].x * r[u].x;
                s += r[u].y * r[u].y;
                st_f64x2_hint(a.R + 2 * j, r[u], pol_last);       // K3 reads r next
            }
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && tid == 0) {
        const int j = a.n - 1;
        double r = a.R[j] + na * a.AP[j];
        a.R[j] = r;
        s += r * r;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((tid & 31) == 0) s_red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int i = 0; i < kWarps; ++i) t += s_red[i];
        a.part[blockIdx.x] = t;
    }
    if (!last_cta_election(a.ticket, gridDim.x)) return;
    cg_finalize_r(a, cg, s_red, s_cnt);
}

__global__ void __launch_bounds__(kThreads)
cg1_update_xp_kernel(CgVecArgs a, CgScalars cg)
{
    griddep_wait();
    griddep_launch_dependents();
    if (cg.ctrl[CTRL_HALT]) return;
    const bool final_iter = cg.ctrl[CTRL_STOP] != 0;
    const int tid = threadIdx.x;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    const double al = cg.alpha[0], be = cg.beta[0];
    const long long n2 = a.n >> 1;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + tid; i < n2; i += stride * kVecUnroll) {
        double2 x[kVecUnroll], p[kVecUnroll], r[kVecUnroll];
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < n2) {
                x[u] = ld_f64x2_hint(a.X + 2 * j, pol_first);      // x is touched once per iteration
                p[u] = ld_f64x2_hint(a.P + 2 * j, pol_last);
                if (!final_iter) r[u] = ld_f64x2_hint(a.R + 2 * j, pol_last);
            }
        }
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < n2) {
                x[u].x += al * p[u].x;
                x[u].y += al * p[u].y;
                st_f64x2_hint(a.X + 2 * j, x[u], pol_first);
                if (!final_iter) {
                    p[u].x = r[u].x + be * p[u].x;
                    p[u].y = r[u].y + be * p[u].y;
                    st_f64x2_hint(a.P + 2 * j, p[u], pol_last);    // the SpMV gathers p next
                }
            }
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && tid == 0) {
        const int j = a.n - 1;
        const double pj = a.P[j];
        a.X[j] += al * pj;
        if (!final_iter) a.P[j] = a.R[j] + be * pj;
    }
}

} // namespace smle
