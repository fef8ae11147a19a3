// smle_dist.cuh -- row-partitioned single-RHS CG over NVLink peer memory.
//
// The reference has no distributed code; this is the B200-native extension north_star asks for:
// one process per GPU, contiguous row blocks cut at the reference's merge-path coordinates
// (MergePathSearch on diagonals g*ceil((m+nnz)/G)), and NO library collective on the data path:
//
//   * halo exchange: after p is updated, every rank PUSHES the entries its neighbours need
//     straight into the halo tail of their p vector (peer pointers from CUDA IPC, stores over
//     NVLink), fences and publishes a sequence number in the neighbour's memory; only the BOUNDARY
//     TILES of the next SpMV wait for the incoming numbers, its interior tiles start at once;
//   * dot products: every rank posts its partial sum into a mailbox slot in EVERY peer's memory;
//     one warp of every CTA of the consuming kernel polls its local mailbox (lane q <- rank q) and
//     adds the G values in rank order, so all ranks obtain bit-identical alpha / beta /
//     convergence decisions with one NVLink write latency instead of a collective call
//     (flag-in-data slots: smle_distctl.cuh).
//
// Sequence numbers are derived from the device-resident iteration counter, so the whole
// iteration (3 or 4 launches) replays from one CUDA graph.  Mailbox slots are double-buffered by
// iteration parity; the reduction dependency chain keeps ranks within one step of each other,
// which makes the reuse safe.
#pragma once
#include "smle_cg.cuh"
#include "smle_common.cuh"
#include "smle_distctl.cuh"

namespace smle {

// One warp: lane q posts this rank's partial `*value` into peer q's mailbox.
__global__ void dist_post_kernel(DistCtl d, int kind, const double *value, const int *ctrl)
{
    if (kind != 2 && ctrl[CTRL_STOP]) return;
    dist_post(d, kind, *value, ctrl);
}

// Measurement aid (A/B against a library all-reduce): one all-reduce of a double through the mailboxes
// = dist_post_kernel + this kernel, which waits for the G partials, writes their rank-ordered sum and
// advances the iteration counter the sequence numbers derive from.
__global__ void dist_allreduce_wait_kernel(DistCtl d, double *out, int *ctrl)
{
    const int it = ctrl[CTRL_ITER];
    const double s = dist_wait_sum(d, 0, it & 1, dist_mail_seq(ctrl, 0, it));
    if (threadIdx.x == 0) { *out = s; ctrl[CTRL_ITER] = it + 1; }
}

// K2 (distributed): alpha from the all-reduced p.Ap, r -= alpha*Ap, local r.r posted to every rank.
// The first batch of r / Ap is requested BEFORE the mailbox wait, so the NVLink latency of the
// reduction overlaps the DRAM latency of the sweep's first loads.
__global__ void __launch_bounds__(kThreads)
cg1d_update_r_kernel(CgVecArgs a, CgScalars cg, DistCtl d)
{
    __shared__ double s_red[kThreads];
    __shared__ double s_alpha;
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    const long long n2 = a.n >> 1;
    const long long stride = (long long)gridDim.x * kThreads;
    long long i = (long long)blockIdx.x * kThreads + tid;
    double2 r[kVecUnroll], ap[kVecUnroll];
    auto load = [&](long long base) {
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = base + u * stride;
            if (j < n2) {
                r[u] = ld_f64x2_hint(a.R + 2 * j, pol_last);
                ap[u] = ld_f64x2_hint(a.AP + 2 * j, pol_first);
            }
        }
    };
    if (i < n2) load(i);
    if (tid < 32) {
        const int it = cg.ctrl[CTRL_ITER];
        const double pAp = dist_wait_sum(d, 0, it & 1, dist_mail_seq(cg.ctrl, 0, it));
        if (tid == 0) {
            s_alpha = cg.rs_old[0] / pAp;
            if (blockIdx.x == 0) cg.alpha[0] = s_alpha;
        }
    }
    __syncthreads();
    const double na = -s_alpha;
    double s = 0.0;
    while (i < n2) {
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < n2) {
                r[u].x += na * ap[u].x;
                r[u].y += na * ap[u].y;
                s += r[u].x * r[u].x;
                s += r[u].y * r[u].y;
                st_f64x2_hint(a.R + 2 * j, r[u], pol_last);
            }
        }
        i += stride * kVecUnroll;
        if (i < n2) load(i);
    }
    if ((a.n & 1) && blockIdx.x == 0 && tid == 0) {
        const int j = a.n - 1;
        double rj = a.R[j] + na * a.AP[j];
        a.R[j] = rj;
        s += rj * rj;
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dd);
    if ((tid & 31) == 0) s_red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int w = 0; w < kWarps; ++w) t += s_red[w];
        a.part[blockIdx.x] = t;
    }
    if (!last_cta_election(a.ticket, gridDim.x)) return;
    const double rr = cta_reduce_one<double>(a.part, nullptr, gridDim.x, s_red);   // local r.r
    if (tid == 0) cg.rs_new[0] = rr;
    dist_post(d, 1, rr, cg.ctrl);                                                   // -> every rank's mailbox
}

// K3 (distributed).  mode 0: beta / convergence from the all-reduced r.r, x += alpha p,
// p = r + beta p (grid-stride: all CTAs sweep one moving window, which keeps the five HBM streams
// page-friendly -- a per-CTA chunked sweep measured 20 % slower); rows on the slab boundary also go
// straight into the neighbours' halo tails (fused push).  The last CTA publishes the halo sequence
// numbers and advances the iteration state; it does NOT wait for the neighbours -- the next SpMV's
// boundary tiles do (dist_wait_halo), its interior tiles start at once.
// mode 1 (after cg_init_kernel): b.b all-reduce -> rs_old / bnorm.
__global__ void __launch_bounds__(kThreads)
cg1d_update_xp_kernel(CgVecArgs a, CgScalars cg, DistCtl d, int mode)
{
    __shared__ double s_rr, s_beta;
    __shared__ int s_final, s_it;
    if (cg.ctrl[CTRL_STOP]) return;
    const int tid = threadIdx.x;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    const long long n2 = a.n >> 1;
    const long long stride = (long long)gridDim.x * kThreads;
    const long long i0 = (long long)blockIdx.x * kThreads + tid, step = stride * kVecUnroll;
    // This thread's batches in the order 0, last, 1, 2, ...: the first and the last rows of the slab are
    // what the neighbours need, so both are pushed at the START of the sweep and the system-scope fence at
    // its end finds those NVLink writes long completed instead of waiting a round trip for the last ones.
    const long long nb = i0 < n2 ? (n2 - i0 + step - 1) / step : 0;
    auto batch_start = [&](long long t) { return i0 + (t == 0 ? 0 : (t == 1 ? nb - 1 : t - 1)) * step; };
    long long i = i0;
    bool pushed = false;
    double2 x[kVecUnroll], p[kVecUnroll], r[kVecUnroll];
    auto load = [&](long long base) {
#pragma unroll
        for (int u = 0; u < kVecUnroll; ++u) {
            const long long j = base + u * stride;
            if (j < n2) {
                x[u] = ld_f64x2_hint(a.X + 2 * j, pol_first);
                p[u] = ld_f64x2_hint(a.P + 2 * j, pol_last);
                r[u] = ld_f64x2_hint(a.R + 2 * j, pol_last);
            }
        }
    };
    if (mode == 0 && i < n2) load(i);   // in flight while the reduction arrives
    if (tid < 32) {
        const int it = cg.ctrl[CTRL_ITER];
        if (mode == 1) {
            const double bb = dist_wait_sum(d, 2, 0, dist_mail_seq(cg.ctrl, 2, 0));
            if (tid == 0) { s_it = it; s_rr = bb; s_beta = 0.0; s_final = 0; }
        } else {
            const double rr = dist_wait_sum(d, 1, it & 1, dist_mail_seq(cg.ctrl, 1, it));
            if (tid == 0) {
                const double rel = sqrt(rr) / cg.bnorm[0];
                s_it = it;
                s_rr = rr;
                s_beta = rr / cg.rs_old[0];
                s_final = (rel < *cg.tol) || (it + 1 >= cg.ctrl[CTRL_MAX_ITERS]);
            }
        }
    }
    __syncthreads();
    if (mode == 0) {
        const bool final_iter = s_final != 0;
        const double al = cg.alpha[0], be = s_beta;
        for (long long t = 0; t < nb; ++t) {
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
                        st_f64x2_hint(a.P + 2 * j, p[u], pol_last);
                        if (d.fused) {
                            // rows on this rank's boundary go straight into the neighbours' halo tails (NVLink)
                            for (int g = 0; g < d.npush; ++g) {
                                const int off = (int)(2 * j) - d.push_lo[g];
                                double *dst = d.peer_p[d.push_q[g]] + d.send_dst[d.push_q[g]];
                                if ((unsigned)off < (unsigned)d.push_cnt[g]) { dst[off] = p[u].x; pushed = true; }
                                if ((unsigned)(off + 1) < (unsigned)d.push_cnt[g]) { dst[off + 1] = p[u].y; pushed = true; }
                            }
                        }
                    }
                }
            }
            if (t + 1 < nb) { i = batch_start(t + 1); load(i); }
        }
        if ((a.n & 1) && blockIdx.x == 0 && tid == 0) {
            const int j = a.n - 1;
            const double pj = a.P[j];
            a.X[j] += al * pj;
            if (!final_iter) {
                const double pn = a.R[j] + be * pj;
                a.P[j] = pn;
                if (d.fused)
                    for (int g = 0; g < d.npush; ++g) {
                        const int off = j - d.push_lo[g];
                        if ((unsigned)off < (unsigned)d.push_cnt[g]) { d.peer_p[d.push_q[g]][d.send_dst[d.push_q[g]] + off] = pn; pushed = true; }
                    }
            }
        }
        if (pushed) __threadfence_system();   // this thread's peer stores are ordered before the sequence numbers
    }
    if (!last_cta_election(a.ticket, gridDim.x)) return;
    if (mode == 0 && d.fused && !s_final) {
        // the halo of the p for iteration it+1 has left: tell the neighbours
        const unsigned long long seq = (unsigned long long)cg.ctrl[CTRL_SEQ_BASE] + (unsigned long long)s_it + 2ull;
        const int q = tid;
        if (q < d.world && q != d.rank && d.send_off[q + 1] > d.send_off[q]) st_release_sys(&d.peer[q]->halo_seq[d.rank], seq);
    }
    if (tid == 0) {
        if (mode == 1) {
            const double nb = sqrt(s_rr);
            cg.rs_old[0] = s_rr;
            cg.bnorm[0] = nb == 0.0 ? 1.0 : nb;
        } else {
            const int it = s_it;
            const double rel = sqrt(s_rr) / cg.bnorm[0];
            if (cg.hist && it < cg.hist_cap) cg.hist[it] = rel;
            *cg.last_rel = rel;
            cg.rs_new[0] = s_rr;
            cg.rs_old[0] = s_rr;
            cg.ctrl[CTRL_ITER] = it + 1;
            if (s_final) cg.ctrl[CTRL_STOP] = 1;
        }
    }
}

// Halo push: P[send_idx] -> the halo tail of each neighbour's p vector (NVLink peer stores), then the
// sequence numbers.  The consumers of the halo (boundary tiles of the next SpMV) wait for them.
__global__ void __launch_bounds__(kThreads)
dist_halo_push_kernel(DistCtl d, const double *__restrict__ P, const int *ctrl)
{
    if (ctrl[CTRL_STOP]) return;
    const unsigned long long seq = (unsigned long long)ctrl[CTRL_SEQ_BASE] + (unsigned long long)ctrl[CTRL_ITER] + 1ull;
    const int total = d.send_off[d.world];
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        int q = 0;
        while (i >= d.send_off[q + 1]) ++q;
        d.peer_p[q][d.send_dst[q] + (i - d.send_off[q])] = P[d.send_idx[i]];
    }
    __threadfence_system();
    if (!last_cta_election(d.ticket, gridDim.x)) return;
    const int q = threadIdx.x;
    if (q < d.world && q != d.rank && d.send_off[q + 1] > d.send_off[q]) st_release_sys(&d.peer[q]->halo_seq[d.rank], seq);
}

} // namespace smle
