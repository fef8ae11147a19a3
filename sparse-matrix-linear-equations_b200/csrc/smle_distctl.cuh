// smle_distctl.cuh -- control structures and peer-memory primitives of the row-partitioned path
// (shared by the SpMV kernel, which posts its dot-product partial, and smle_dist.cuh).
#pragma once
#include "smle_common.cuh"

namespace smle {

constexpr int kMaxRanks = 8;
constexpr int kMailKinds = 3;                 // 0: p.Ap   1: r.r   2: b.b (initialisation)
constexpr size_t kDistCtlBytes = 4096;        // control block in front of the p vector

// layout of the control block (identical on every rank; peers address it through IPC)
struct DistBlock {
    unsigned long long halo_seq[kMaxRanks];                           // written by peer q: halo of step seq has landed
    unsigned long long mail_seq[kMailKinds * 2 * kMaxRanks];
    double mail_val[kMailKinds * 2 * kMaxRanks];
    int error;                                                         // spin-wait timeout seen
};
static_assert(sizeof(DistBlock) <= kDistCtlBytes, "control block too large");

struct DistCtl {
    int rank, world;
    DistBlock *self;                    // local control block
    DistBlock *peer[kMaxRanks];         // every rank's control block (peer[rank] == self)
    double *peer_p[kMaxRanks];          // every rank's extended p vector [n_local + n_halo]
    const int *send_idx;                // local indices of the entries to push, grouped by peer
    int send_off[kMaxRanks + 1];        // group boundaries
    int send_dst[kMaxRanks];            // element offset in the peer's p vector where my group lands
    int needs_from[kMaxRanks];          // 1 when this rank receives halo entries from peer q
    unsigned int *ticket;
    // fused halo push (every send group is one contiguous run of local rows, e.g. the boundary planes
    // of a slab): K3 stores the new p of those rows straight into the neighbours' halo tails
    int fused;                          // 1: K3 pushes, no separate push kernel inside the iteration
    int npush;                          // peers with a non-empty group
    int push_q[kMaxRanks];              // their ranks
    int push_lo[kMaxRanks];             // first local row of the group
    int push_cnt[kMaxRanks];            // rows in the group
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ bool dist_spin(const unsigned long long *flag, unsigned long long want, int *error)
{
    long long spins = 0;
    while (ld_acquire_sys(flag) < want) {
        if (++spins > (1ll << 28)) {   // ~ seconds: a peer died or the launch order is broken
            *error = 1;
            return false;
        }
    }
    return true;
}

// sum of the `kind` mailbox over ranks in rank order (called by one thread)
__device__ __forceinline__ double dist_wait_sum(const DistCtl &d, int kind, int parity, unsigned long long seq)
{
    double s = 0.0;
    for (int q = 0; q < d.world; ++q) {
        const int idx = (kind * 2 + parity) * kMaxRanks + q;
        dist_spin(&d.self->mail_seq[idx], seq, &d.self->error);
        s += *(volatile double *)&d.self->mail_val[idx];
    }
    return s;
}

// Post this rank's partial `value` of reduction `kind` for iteration `it` into every peer's
// mailbox (called by the threads q < world of one CTA).
__device__ __forceinline__ void dist_post(const DistCtl &d, int kind, double value, const int *ctrl)
{
    const int it = ctrl[CTRL_ITER];
    const unsigned long long seq = (unsigned long long)ctrl[CTRL_SEQ_BASE] + (kind == 2 ? 1ull : (unsigned long long)it + 1ull);
    const int parity = kind == 2 ? 0 : (it & 1);
    const int q = threadIdx.x;
    if (q < d.world) {
        const int idx = (kind * 2 + parity) * kMaxRanks + d.rank;
        *(volatile double *)&d.peer[q]->mail_val[idx] = value;
        __threadfence_system();
        st_release_sys(&d.peer[q]->mail_seq[idx], seq);
    }
}

} // namespace smle
