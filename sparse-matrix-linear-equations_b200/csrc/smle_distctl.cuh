// smle_distctl.cuh -- control structures and peer-memory primitives of the row-partitioned path
// (shared by the SpMV kernel, which posts its dot-product partial and waits for the halo in its
// boundary tiles, and smle_dist.cuh).
//
// Dot products travel through a MAILBOX in every rank's memory.  A slot is two 64-bit words, each
// carrying half of the double next to the 32-bit sequence number of the reduction it belongs to
// (the layout of NCCL's LL protocol): a naturally aligned 8-byte store is a single NVLink write, so
// a reader that sees the expected sequence number in both words has the whole value -- no fence on
// the writer, no acquire on the reader, one NVLink write latency end to end.  Slot reuse is safe
// because the reductions themselves keep ranks within one iteration of each other.
#pragma once
#include "smle_common.cuh"

namespace smle {

constexpr int kMaxRanks = 8;
constexpr int kMailKinds = 3;                 // 0: p.Ap   1: r.r   2: b.b (initialisation)
constexpr size_t kDistCtlBytes = 4096;        // control block in front of the p vector
constexpr unsigned long long kDistTimeoutNs = 4000000000ull;   // a peer that stays silent this long is dead

// layout of the control block (identical on every rank; peers address it through IPC)
struct DistBlock {
    unsigned long long halo_seq[kMaxRanks];                  // written by peer q: its halo push of step seq has landed
    unsigned long long mail[kMailKinds * 2 * kMaxRanks * 2]; // [kind][parity][rank]{lo, hi}: seq32 << 32 | half of the value
    int error;                                               // a wait timed out
};
static_assert(sizeof(DistBlock) <= kDistCtlBytes, "control block too large");

struct DistCtl {
    int rank, world;
    DistBlock *self;                    // local control block
    DistBlock *peer[kMaxRanks];         // every rank's control block (peer[rank] == self)
    double *peer_p[kMaxRanks];          // every rank's extended p vector [halo_base + n_halo]
    const int *send_idx;                // local indices of the entries to push, grouped by peer
    int send_off[kMaxRanks + 1];        // group boundaries
    int send_dst[kMaxRanks];            // element offset in the peer's p vector where my group lands
    int needs_from[kMaxRanks];          // 1 when this rank receives halo entries from peer q
    unsigned int *ticket;
    int *stop;                          // &ctrl[CTRL_STOP] of the solve: raised when a wait times out
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

__device__ __forceinline__ void ld_volatile_u64x2(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void st_volatile_u64x2(unsigned long long *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// A wait gives up when a peer stays silent for kDistTimeoutNs (it died, or the launch order is
// broken): the error flag goes up, STOP turns the kernels still queued into no-ops, and every
// later wait returns at once, so the host gets SMLE_ERR_COMM in seconds instead of hanging.
__device__ __forceinline__ void dist_give_up(const DistCtl &d)
{
    *(volatile int *)&d.self->error = 1;
    if (d.stop) *(volatile int *)d.stop = 1;
}

__device__ __forceinline__ bool dist_spin(const unsigned long long *flag, unsigned long long want, const DistCtl &d)
{
    if (ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = global_timer_ns();
    unsigned int polls = 0;
    while (ld_acquire_sys(flag) < want) {
        if ((++polls & 255u) == 0) {
            if (*(volatile int *)&d.self->error) return false;
            if (global_timer_ns() - t0 > kDistTimeoutNs) { dist_give_up(d); return false; }
        }
    }
    return true;
}

// Sum of the `kind` mailbox over ranks, added in rank order (bit-identical on every rank).  Called by
// ALL 32 lanes of one warp: lane q polls rank q's slot, so the G loads are in flight together.
__device__ __forceinline__ double dist_wait_sum(const DistCtl &d, int kind, int parity, unsigned int seq)
{
    const int lane = threadIdx.x & 31;
    double v = 0.0;
    if (lane < d.world) {
        const unsigned long long *slot = &d.self->mail[((kind * 2 + parity) * kMaxRanks + lane) * 2];
        unsigned long long lo, hi;
        ld_volatile_u64x2(slot, lo, hi);
        if ((unsigned int)(lo >> 32) != seq || (unsigned int)(hi >> 32) != seq) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int polls = 0;
            for (;;) {
                ld_volatile_u64x2(slot, lo, hi);
                if ((unsigned int)(lo >> 32) == seq && (unsigned int)(hi >> 32) == seq) break;
                if ((++polls & 255u) == 0) {
                    if (*(volatile int *)&d.self->error) break;
                    if (global_timer_ns() - t0 > kDistTimeoutNs) { dist_give_up(d); break; }
                }
            }
        }
        v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
    }
    double s = 0.0;
    for (int q = 0; q < d.world; ++q) s += __shfl_sync(0xffffffffu, v, q);
    return s;
}

// sequence number of reduction `kind` in iteration `it` of the solve that started at seq_base
__device__ __forceinline__ unsigned int dist_mail_seq(const int *ctrl, int kind, int it)
{
    return (unsigned int)ctrl[CTRL_SEQ_BASE] + (kind == 2 ? 1u : (unsigned int)it + 1u);
}

// Post this rank's partial `value` of reduction `kind` for the current iteration into every rank's
// mailbox (called by the threads q < world of one CTA; thread `rank` writes the local slot).
__device__ __forceinline__ void dist_post(const DistCtl &d, int kind, double value, const int *ctrl)
{
    const int it = ctrl[CTRL_ITER];
    const unsigned long long seq = dist_mail_seq(ctrl, kind, it);
    const int parity = kind == 2 ? 0 : (it & 1);
    const int q = threadIdx.x;
    if (q < d.world) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
        unsigned long long *slot = &d.peer[q]->mail[((kind * 2 + parity) * kMaxRanks + d.rank) * 2];
        st_volatile_u64x2(slot, (seq << 32) | (bits & 0xffffffffull), (seq << 32) | (bits >> 32));
    }
}

// Halo of the p vector this iteration's SpMV gathers: wait until every neighbour this rank receives
// from has published sequence number seq_base + it + 1.  Called by all lanes of a warp.
__device__ __forceinline__ void dist_wait_halo(const DistCtl &d, const int *ctrl)
{
    const int q = threadIdx.x & 31;
    if (q < d.world && q != d.rank && d.needs_from[q]) {
        const unsigned long long seq = (unsigned long long)ctrl[CTRL_SEQ_BASE] + (unsigned long long)ctrl[CTRL_ITER] + 1ull;
        dist_spin(&d.self->halo_seq[q], seq, d);
    }
    __syncwarp();
}

} // namespace smle
