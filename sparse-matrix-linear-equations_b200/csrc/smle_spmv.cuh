// smle_spmv.cuh -- single-vector merge-path SpMV for sm_100a (k = 1 fast path).
//
// Replaces OmpMergeCsrmv (reference cpu_spmv.cpp:360-421) and the SpMV inside CGSolveSingle
// (work_2025/main/single_strategy.hpp:137).  Same merge-path decomposition as smle_merge.cuh,
// but laid out for one right-hand side, where the kernel is a pure HBM stream of the CSR arrays
// (12 B per nonzero in fp64) plus an L1/L2-resident gather of x:
//
//   * persistent CTAs, each owning a contiguous run of tiles of TILE = THREADS*IPT merge items;
//   * the tile's column indices, values and row offsets are brought into shared memory by the
//     TMA engine (cp.async.bulk, 1-D, completion on an mbarrier) STAGES tiles ahead of the
//     compute, with an L2 evict-first policy -- the matrix is streamed once per SpMV and must
//     not push the CG vectors out of the 126 MB L2;
//   * regular tiles (no row segment longer than 32, every stencil): one thread per row straight from
//     the stage buffers -- adjacent lanes own adjacent rows, so a warp's gathers of x fall into 2-3
//     cache lines and y is written coalesced; no CTA-wide barrier, warps drift across tiles;
//   * general tiles (R-MAT hubs, the wheel's hub row): product-staged -- every thread requests the x of
//     IPT nonzeros at once, the products replace the staged values, and after one barrier the rows are
//     summed from shared memory: up to 32 nonzeros by a thread, up to 1024 by a warp taking rows from the
//     tile's queue, longer segments by the whole CTA (one more barrier); a tile that lies inside one row
//     keeps its products in registers.  The classic per-thread merge walk with its three barriers per
//     tile and its bank conflicts is gone;
//   * carries: per-tile -> per-CTA in shared memory; the row cut by a CTA boundary is finished
//     wait-free through one global slot per boundary (the party that arrives second adds owner
//     part + carry, merge_based.hpp:137-149 semantics), so a plain SpMV has no last-CTA epilogue;
//   * with DOT the products y[r]*x[r] (the p.Ap of CG) are accumulated from the same registers; the
//     dot product is linear in the parts of a row, so carries add their share where they are formed.
#pragma once
#include <type_traits>

#include "smle_common.cuh"
#include "smle_distctl.cuh"
#include "smle_merge.cuh"

namespace smle {

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                            uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename V>
struct SpmvArgs {
    const int *__restrict__ ro;        // row offsets (m + 1 entries, 16 B of slack behind)
    const int *__restrict__ ci;
    const V *__restrict__ va;
    const V *__restrict__ x;
    V *__restrict__ y;
    const int2 *__restrict__ tile_xy;  // num_tiles + 1 merge-path coordinates
    const int *__restrict__ tile_maxlen;  // longest in-tile row segment per tile (matrix property)
    int m, nnz;
    int num_tiles, tiles_per_cta;
    V *cta_slot;                       // [gridDim.x] carry slots of the CTA boundaries (sentinel when empty)
    V *dot_part;                       // [gridDim.x]  (DOT)
    unsigned int *ticket;
    int debug_flags;                   // 1: skip compute (stream-only ceiling of the TMA pipeline)
    int med_lo;                        // general tiles: segments longer than this are queued for a warp (default kRowPathMaxLen)
    int keep_l2;                       // 1: the whole system fits in L2 -- stream the matrix with the normal priority so that
                                       //    the next product finds it there (else evict-first: read once per product)
    const DistCtl *dist;               // row-partitioned solve: halo waits, p.Ap posted to every peer (else NULL)
    const unsigned char *tile_halo;    // row-partitioned solve: 1 for tiles that gather halo columns (else NULL)
};

template <typename V, int THREADS, int IPT>
struct SpmvSmem {
    static constexpr int TILE = THREADS * IPT;
    static constexpr int EPV = 16 / (int)sizeof(V);                  // values per 16 bytes
    static constexpr int COL_WORDS = TILE + 8;                       // staged column indices
    static constexpr int VAL_ELEMS = TILE + 2 * EPV;                 // staged values
    static constexpr int RO_WORDS = TILE + 8;                        // staged row offsets
    static constexpr size_t HDR_OFFSET =
        ((size_t)COL_WORDS * 4 + (size_t)VAL_ELEMS * sizeof(V) + (size_t)RO_WORDS * 4 + 15) / 16 * 16;
    // tile header written by the producer: {lo.x, lo.y, hi.x, hi.y, longest row segment, gathers halo columns}
    static constexpr size_t STAGE_BYTES = HDR_OFFSET + 32;
};

// longest in-tile row segment for which the fused thread-per-row path is used
constexpr int kRowPathMaxLen = 32;
// longest row segment one warp reduces by itself; longer ones are strided over by the whole CTA
constexpr int kWarpRowMax = 1024;
// capacity of the per-tile queue of such segments (a tile of 3840 nonzeros holds at most 116 rows longer than 32,
// 480 longer than 7: the threshold can be lowered to 8 for experiments)
constexpr int kLongCap = 512;
// segments longer than kWarpRowMax in one tile: at most TILE / (kWarpRowMax + 1) <= 7 for tiles of up to 8192 items
constexpr int kHugeCap = 8;

// One warp per tile: the longest run of nonzeros of a single row inside the tile (complete rows,
// the leading part of row x0 and the trailing part of row x1).  A property of the matrix and the
// tiling only, computed once per handle next to the merge-path coordinates.
__global__ void tile_maxlen_kernel(const int *__restrict__ ro, const int2 *__restrict__ tile_xy, int num_tiles,
                                   int *__restrict__ out)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= num_tiles) return;
    const int2 lo = tile_xy[t], hi = tile_xy[t + 1];
    const int rows = hi.x - lo.x;
    int mx = 0;
    for (int i = lane; i <= rows; i += 32) {
        const int beg = (i == 0) ? lo.y : __ldg(ro + lo.x + i);          // end of the previous row
        const int end = (i == rows) ? hi.y : __ldg(ro + lo.x + i + 1);
        mx = max(mx, end - beg);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) out[t] = mx;
}

// One warp per tile: does the tile gather a halo column (index >= halo_base)?  Row-partitioned
// handles only; the flagged tiles wait for the neighbours' halo push before they gather.
__global__ void tile_halo_kernel(const int *__restrict__ ci, const int2 *__restrict__ tile_xy, int num_tiles,
                                 int halo_base, unsigned char *__restrict__ out)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= num_tiles) return;
    const int lo = tile_xy[t].y, hi = tile_xy[t + 1].y;
    int any = 0;
    for (int z = lo + lane; z < hi; z += 32) any |= __ldg(ci + z) >= halo_base;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) out[t] = any ? 1 : 0;
}

// out[0] = max over v, out[1] = number of entries above `thresh` (the general tiles of a partition)
__global__ void int_max_kernel(const int *__restrict__ v, int n, int thresh, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int val = i < n ? v[i] : 0;
    const int mx = __reduce_max_sync(0xffffffffu, val);
    const unsigned above = __ballot_sync(0xffffffffu, val > thresh);
    if ((threadIdx.x & 31) == 0) {
        if (mx > 0) atomicMax(out, mx);
        if (above) atomicAdd(out + 1, __popc(above));
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// barrier among the THREADS consumer threads only (the producer warp never joins it)
template <int THREADS>
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// CTAs per SM the shared-memory footprint allows (227 KB usable, ~1 KB reserved per CTA)
template <typename V, int THREADS, int IPT, int STAGES>
constexpr int spmv_ctas_per_sm()
{
    constexpr size_t per_cta = SpmvSmem<V, THREADS, IPT>::STAGE_BYTES * STAGES + 4096;
    constexpr int by_smem = (int)((size_t)227 * 1024 / per_cta);
    constexpr int by_threads = 2048 / (THREADS + 32);
    constexpr int n = by_smem < by_threads ? by_smem : by_threads;
    return n < 1 ? 1 : (n > 8 ? 8 : n);
}

// ---------------------------------------------------------------------------------------
// Row path building blocks (regular tiles).  Thread i owns local row i: adjacent lanes own
// adjacent rows, so for banded matrices the x gathers of a warp fall into 2-3 cache lines
// (vs ~8 when lanes walk consecutive nonzeros) and y is written coalesced.  All gathers of a
// thread are issued before the first FMA; the value is read from shared memory only when the
// FMA needs it, which keeps the live registers at 2 per gather.
// ---------------------------------------------------------------------------------------
template <typename V, bool COH>
__device__ __forceinline__ V gather(const V *x)
{
    // COH: the halo tail of x is written by the peers while this kernel runs -> L2-coherent load
    if constexpr (COH) return __ldcg(x); else return __ldg(x);
}

template <typename V, bool COH>
__device__ __forceinline__ V row_sum(const V *__restrict__ x, const int *pc, const V *pv, int beg, int end)
{
    constexpr int UB = 8;
    V sum = 0;
    while (beg < end) {
        // unconditional gathers (slots behind the row repeat its last nonzero: same address, an L1
        // hit): no predicate keeps ptxas from issuing all UB requests before the first FMA
        V xa[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) xa[j] = gather<V, COH>(x + pc[min(beg + j, end - 1)]);
#pragma unroll
        for (int j = 0; j < UB; ++j)
            if (beg + j < end) sum += pv[beg + j] * xa[j];
        beg += UB;
    }
    return sum;
}

// Partial sum of one row segment [beg, end) over the positions beg + id, beg + id + STRIDE, ...
// (id < STRIDE): the lanes of a warp (STRIDE = 32) or the threads of the CTA stride over a long row.
// Four gathers are requested before the first FMA; slots past the end repeat the last nonzero.
template <typename V, bool COH, int STRIDE>
__device__ __forceinline__ V strided_sum(const V *__restrict__ x, const int *pc, const V *pv, int beg, int end, int id)
{
    constexpr int UB = 4;
    V sum = 0;
    for (int z = beg + id; z < end; z += STRIDE * UB) {
        V xa[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) xa[j] = gather<V, COH>(x + pc[min(z + j * STRIDE, end - 1)]);
#pragma unroll
        for (int j = 0; j < UB; ++j)
            if (z + j * STRIDE < end) sum += pv[z + j * STRIDE] * xa[j];
    }
    return sum;
}

// ---- general tiles, product-staged path ---------------------------------------------------------
// The gather of x with an L2 cache hint in a register: evict-normal by default; SMLE_SPMV_DEBUG bit 3 asks
// for evict-last gathers and evict-first y stores (R-MAT scale 24, x = 134 MB against 126 MB of L2: no
// effect measured, profiles/r02_spmv_general_tiles_staged_ab.jsonl).  Halo tiles (COH) keep the
// L2-coherent load.
template <typename V, bool COH>
__device__ __forceinline__ V gather_hint(const V *x, uint64_t pol)
{
    if constexpr (COH) {
        return __ldcg(x);
    } else if constexpr (sizeof(V) == 8) {
        double v;
        asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(x), "l"(pol));
        return (V)v;
    } else {
        float v;
        asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(x), "l"(pol));
        return (V)v;
    }
}

template <typename V>
__device__ __forceinline__ void store_hint(V *p, V v, uint64_t pol)
{
    if constexpr (sizeof(V) == 8)
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"((double)v), "l"(pol) : "memory");
    else
        asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"((float)v), "l"(pol) : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Sum of the staged products of one row segment, in nonzero order (the order of the per-row gather
// path, so both paths round alike).  Four loads are requested before the first add.
template <typename V>
__device__ __forceinline__ V seg_sum(const V *pv, int beg, int end)
{
    V sum = 0;
    for (; beg + 4 <= end; beg += 4) {
        const V p0 = pv[beg], p1 = pv[beg + 1], p2 = pv[beg + 2], p3 = pv[beg + 3];
        sum += p0; sum += p1; sum += p2; sum += p3;
    }
    for (; beg < end; ++beg) sum += pv[beg];
    return sum;
}

// The same over the positions beg + id, beg + id + STRIDE, ... (lanes of a warp / threads of the CTA); four
// independent partial sums, so a pass is bound by the shared-memory pipe and not by the add latency.
template <typename V, int STRIDE>
__device__ __forceinline__ V seg_sum_strided(const V *pv, int beg, int end, int id)
{
    V s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int z = beg + id;
    for (; z + 3 * STRIDE < end; z += 4 * STRIDE) {
        s0 += pv[z]; s1 += pv[z + STRIDE]; s2 += pv[z + 2 * STRIDE]; s3 += pv[z + 3 * STRIDE];
    }
    for (; z < end; z += STRIDE) s0 += pv[z];
    return (s0 + s1) + (s2 + s3);
}

// Sum of n <= 32 per-warp parts by one warp (fixed tree: the result does not depend on timing)
template <typename V>
__device__ __forceinline__ V warp_sum_parts(const V *parts, int n, int lane)
{
    V v = lane < n ? parts[lane] : V(0);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Publisher side of the CTA-boundary exchange.  val = sum of the parts of row R = (row in progress
// at the end of CTA c) that lie in CTAs <= c.  If the owner's part is already in slot c, finish the
// row (owner part + carry); if CTA c+1 lies entirely inside the row, add its part and move on.
template <typename V>
__device__ __noinline__ void cta_carry_publish(const int2 *__restrict__ tile_xy, int tiles_per_cta, int num_tiles, int m,
                                               V *cta_slot, V *y, int c, V val, int row, int next)
{
    // row / next: the rows in progress at the end of CTA c and of CTA c+1, read by the caller at kernel start
    // (the epilogue of a 20 us launch should not wait for two more dependent L2 round trips).  While the carry
    // is passed through CTAs that lie inside one row (the wheel's hub spans 49 of them), the row stays the
    // same and a link costs one atomic exchange: the end row of the CTA after next is requested before the
    // exchange, and the fence is needed only in front of the final store (64 -> ~17 us of serial tail).
    if (row >= m) return;
    for (;;) {
        V *slot = cta_slot + c;
        const int next2 = tile_xy[min((c + 3) * tiles_per_cta, num_tiles)].x;
        const V owner = slot_exchange<V>(slot, val);
        if (is_sentinel<V>(owner)) return;                 // the other party comes later and finishes
        slot_reset<V>(slot);
        if (next > row) {                                  // the row ends in CTA c+1
            __threadfence();                               // its earlier stores to y[row] come before ours
            y[row] = owner + val;
            return;
        }
        val = val + owner;                                 // CTA c+1 lies inside the row: pass on
        ++c;
        next = next2;
    }
}

// ---------------------------------------------------------------------------------------
// spmv_kernel: THREADS consumer threads + one producer warp per CTA.
//
//   producer warp (one elected lane): for every tile of the CTA, waits until the stage buffer
//       is released (mbarrier "empty", one arrival per consumer warp) and issues the three
//       bulk copies of the tile (column indices, values, row offsets) onto the stage's "full"
//       mbarrier.  It runs STAGES tiles ahead and never touches the data.
//   consumers, regular tile (longest in-tile row segment <= kRowPathMaxLen, the common case):
//       wait "full" -> thread-per-row gather + FMA straight from the stage buffers -> arrive
//       on "empty".  NO CTA-wide barrier: warps drift freely across tiles.
//   consumers, general tile (long or wildly uneven rows): the same pass; rows of 33..1024 nonzeros
//       are queued and dealt to the warps after one barrier, longer segments go to the whole CTA.
//   carries: every tile records (first row, has-complete-row, carry-out) in shared memory;
//       thread 0 chains them in tile order every kChainTiles tiles and the fix-ups are applied
//       in parallel (reference semantics: merge_based.hpp:137-149, carry added after the row's
//       owner wrote its partial sum).
// ---------------------------------------------------------------------------------------
constexpr int kChainTiles = 512;   // tiles between two carry-chain resolutions of a CTA

//   MAXB > 0 caps the CTAs per SM below what shared memory allows (registers follow: 64 at two CTAs of 512
//   threads, where three would get 40 and spill): the configuration for skewed matrices, 480x4x2 at two CTAs
//   per SM, leaves ~92 KB of L1 to the scattered x gathers instead of ~28 KB (R-MAT scale 24: 3.9 -> 2.1 ms)
template <typename V, int THREADS, int IPT, int STAGES, bool DOT, int MAXB = 0>
__global__ void __launch_bounds__(THREADS + 32, (MAXB > 0 && MAXB < spmv_ctas_per_sm<V, THREADS, IPT, STAGES>())
                                                    ? MAXB : spmv_ctas_per_sm<V, THREADS, IPT, STAGES>())
spmv_kernel(SpmvArgs<V> a, CgScalars cg)
{
    using SM = SpmvSmem<V, THREADS, IPT>;
    constexpr int NW = THREADS / 32;
    constexpr int EPV = SM::EPV;
    static_assert(SM::TILE / (kWarpRowMax + 1) < kHugeCap && NW >= kHugeCap - 1, "huge-segment bookkeeping of the general tiles");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_full[STAGES], s_empty[STAGES];
    __shared__ V s_wsum[NW];
    __shared__ V s_hub[2][kHugeCap][NW];            // per-warp parts of the (at most kHugeCap) huge segments of a general tile, or of a tile
                                             // that lies inside one row ([.][0]); double-buffered by general-tile parity
    __shared__ int s_hq[3][kHugeCap], s_nhq[3];     // the huge segments of a tile (product-staged path; triple-buffered like s_long)
    __shared__ int s_huge[kHugeCap], s_nhuge;       // row segments longer than kWarpRowMax in the current tile
    __shared__ int s_long[3][kLongCap], s_nlong[3], s_next[3];   // queued segments of 33..kWarpRowMax (balance mode)
    __shared__ V s_tcarry[kChainTiles];      // carry-out of each tile of the current chunk
    __shared__ int s_trow[kChainTiles];      // first row of the tile, or -1 when no row completes in it
    __shared__ V s_running;                  // carry chained so far (row in progress at the chunk start)
    __shared__ int s_edge[3];                // rows in progress at the start / end of this CTA and at the end of the next one
    __shared__ V s_red[THREADS + 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if constexpr (DOT) {
        // STOP was written two launches back (K2): complete by now, see griddep_wait().  HALT is read
        // by the K3 that may still be running, so it is raised only after that kernel has finished.
        if (cg.ctrl[CTRL_STOP]) {
            griddep_wait();
            if (blockIdx.x == 0 && tid == 0) cg.ctrl[CTRL_HALT] = 1;
            return;
        }
    }

    const int t0 = blockIdx.x * a.tiles_per_cta;
    const int t1 = min(t0 + a.tiles_per_cta, a.num_tiles);

    auto stage_col = [&](int s) { return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES); };
    auto stage_val = [&](int s) {
        return reinterpret_cast<V *>(smem_raw + (size_t)s * SM::STAGE_BYTES + (size_t)SM::COL_WORDS * 4);
    };
    auto stage_ro = [&](int s) {
        return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES + (size_t)SM::COL_WORDS * 4 +
                                       (size_t)SM::VAL_ELEMS * sizeof(V));
    };
    auto stage_hdr = [&](int s) { return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES + SM::HDR_OFFSET); };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_running = 0;
        s_nhuge = 0;
        for (int q = 0; q < 3; ++q) { s_nlong[q] = 0; s_next[q] = 0; s_nhq[q] = 0; }
    }
    __syncthreads();

    V dot = 0;

    if (warp == NW) {
        // =============================== producer warp ===========================================
        // All 32 lanes fetch the metadata of the next 32 tiles with one coalesced request per array (a
        // dependent L2 round trip per tile in the single issuing lane delayed every stage of a short launch:
        // grid2d 1000^2 has 11 tiles per CTA); lane 0 waits for the stage and issues the bulk copies.
        // the matrix is read once per SpMV: evict-first, unless matrix and vectors together stay in L2
        const uint64_t pol_stream = a.keep_l2 ? l2_policy_evict_normal() : l2_policy_evict_first();
        int2 lo = a.tile_xy[min(t0, a.num_tiles)];
        int2 hi_l = make_int2(0, 0);
        int ml_l = 0, halo_l = 0;
        for (int t = t0; t < t1; ++t) {
            const int it = t - t0, s = it % STAGES, q = it & 31;
            if (q == 0) {
                const int tt = min(t + lane, a.num_tiles - 1);
                hi_l = a.tile_xy[tt + 1];
                ml_l = a.tile_maxlen[tt];
                halo_l = a.tile_halo ? (int)a.tile_halo[tt] : 0;
            }
            const int2 hi = make_int2(__shfl_sync(0xffffffffu, hi_l.x, q), __shfl_sync(0xffffffffu, hi_l.y, q));
            const int tile_ml_p = __shfl_sync(0xffffffffu, ml_l, q), tile_halo_p = __shfl_sync(0xffffffffu, halo_l, q);
            if (lane == 0) {
                if (it >= STAGES) {
                    mbar_wait(&s_empty[s], (uint32_t)(it / STAGES - 1) & 1u);
                    fence_proxy_async();
                }
                // the tile's coordinates and class travel with its data (released by the arrive below):
                // the consumers read them from shared memory after the wait on "full"
                int *hdr = stage_hdr(s);
                hdr[0] = lo.x; hdr[1] = lo.y; hdr[2] = hi.x; hdr[3] = hi.y;
                hdr[4] = tile_ml_p;
                hdr[5] = tile_halo_p;
                const int yc = lo.y & ~3;                                   // 16 B aligned column start
                const int yv = lo.y & ~(EPV - 1);                           // 16 B aligned value start
                const int rb = (lo.x + 1) & ~3;                             // 16 B aligned row-offset start
                const uint32_t nb_col = (uint32_t)((hi.y - yc + 3) & ~3) * 4u;
                const uint32_t nb_val = (uint32_t)((hi.y - yv + EPV - 1) & ~(EPV - 1)) * (uint32_t)sizeof(V);
                const uint32_t nb_ro = (uint32_t)((hi.x + 2 - rb + 3) & ~3) * 4u;
                mbar_expect_tx(&s_full[s], nb_col + nb_val + nb_ro);        // nb_ro > 0 always
                if (nb_col) tma_load_1d(stage_col(s), a.ci + yc, nb_col, &s_full[s], pol_stream);
                if (nb_val) tma_load_1d(stage_val(s), a.va + yv, nb_val, &s_full[s], pol_stream);
                tma_load_1d(stage_ro(s), a.ro + rb, nb_ro, &s_full[s], pol_stream);
            }
            lo = hi;
        }
    } else {
        // =============================== consumer warps ==========================================
        // The producer above streams the (immutable) matrix right away; x and y belong to the previous
        // kernel until it has completed.
        if constexpr (DOT) { griddep_wait(); griddep_launch_dependents(); }
        if (tid == 0) {   // for the epilogue; the loads complete in the shadow of the first tile's arrival
            const int e0 = a.tile_xy[min(t0, a.num_tiles)].x, e1 = a.tile_xy[min(t1, a.num_tiles)].x;
            const int e2 = a.tile_xy[min(t1 + a.tiles_per_cta, a.num_tiles)].x;
            s_edge[0] = e0; s_edge[1] = e1; s_edge[2] = e2;
        }
        bool halo_ready = false;
        // general tiles: L2 priorities of the x gathers and the y stores (debug_flags bit 3: x evict-last, y evict-first)
        const uint64_t pol_x = (a.debug_flags & 8) ? l2_policy_evict_last() : l2_policy_evict_normal();
        const uint64_t pol_y = (a.debug_flags & 8) ? l2_policy_evict_first() : l2_policy_evict_normal();
        int gen_count = 0;   // general tiles processed by this CTA so far (same in every thread)
        for (int t = t0; t < t1; ++t) {
            const int it = t - t0, s = it % STAGES, slot = it % kChainTiles;
            const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&s_full[s], parity);
            const int *hdr = stage_hdr(s);
            const int2 lo = make_int2(hdr[0], hdr[1]), hi = make_int2(hdr[2], hdr[3]);
            const int tile_ml = hdr[4];
            const bool tile_halo = hdr[5] != 0;
            const int x0 = lo.x, y0 = lo.y;
            const int rows = hi.x - x0, nz = hi.y - y0;
            const int yc = y0 & ~3, yv = y0 & ~(EPV - 1), rb = (x0 + 1) & ~3;
            int *s_col = stage_col(s);
            V *s_val = stage_val(s);
            const int *s_re = stage_ro(s) + (x0 + 1 - rb);   // s_re[i] = end offset of local row i
            const int *pc = s_col + (y0 - yc);                 // pc[z] = column of local nonzero z
            V *pv = s_val + (y0 - yv);                         // pv[z] = value (or product) of local nonzero z

            if (tid == 0) s_trow[slot] = rows > 0 ? x0 : -1;

            // row-partitioned solve: only tiles that gather halo columns wait for the neighbours' push of
            // this iteration's p, and a warp waits ONCE per launch (the sys-scope acquire invalidates L1:
            // per tile it cost 100 us on the CTAs that own the boundary); interior tiles never look
            if (tile_halo && !halo_ready) { dist_wait_halo(*a.dist, cg.ctrl); halo_ready = true; }

            if (a.debug_flags & 1) {
                // measurement aid: stream the tile through shared memory and do nothing with it
                if (tid == 0) s_tcarry[slot] = 0;
            } else if (tile_ml <= kRowPathMaxLen) {
                // ---- regular tile: fused thread-per-row path, no barrier --------------------------------
                // Pseudo-row `rows` is the trailing part of row hi.x: its sum is the tile carry-out.
                auto row_path = [&](auto coh) {
                for (int i = tid; i <= rows; i += THREADS) {
                    const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                    const int end = (i == rows) ? nz : s_re[i] - y0;
                    const V sum = row_sum<V, decltype(coh)::value>(a.x, pc, pv, beg, end);
                    [[maybe_unused]] V xr = 0;
                    if constexpr (DOT) {
                        // after the gathers: x[row] was just fetched for the diagonal entry (an L1 hit) and the
                        // load does not take one of the few load slots while the gathers are in flight
                        if (i < rows) xr = __ldg(a.x + x0 + i);
                    }
                    if (i < rows) {
                        a.y[x0 + i] = sum;
                        if constexpr (DOT) dot += sum * xr;
                    } else {
                        s_tcarry[slot] = sum;
                    }
                }
                };
                if (tile_halo) row_path(std::true_type{}); else row_path(std::false_type{});
            } else if ((a.debug_flags & 4) == 0) {
                // ---- general tile, product-staged: one gather per THREAD-item, then reductions from shared memory
                // A power-law tile holds few rows (R-MAT x16: ~170 of 2880 items), so a thread per row leaves
                // two thirds of the CTA without a gather to issue and walks its row in dependent rounds of global
                // latency.  Here every thread requests the x of IPT nonzeros at once (coalesced over the staged
                // indices), the products replace the staged values, and after ONE barrier the rows are summed from
                // shared memory: <= med_lo by a thread, <= kWarpRowMax by a warp taking rows from the tile's queue,
                // longer segments by the whole CTA.  A tile that lies inside one row (a hub row spans many tiles)
                // never writes its products back: thread sums -> warp sums -> warp 0.
                const int qc = gen_count % 3, qn = (gen_count + 1) % 3, hp = gen_count & 1;
                ++gen_count;
                if (tid == 0) { s_nlong[qn] = 0; s_next[qn] = 0; s_nhq[qn] = 0; }
                auto emit = [&](int i, V sum) {
                    if (i < rows) {
                        store_hint<V>(a.y + x0 + i, sum, pol_y);
                        if constexpr (DOT) dot += sum * __ldg(a.x + x0 + i);
                    } else {
                        s_tcarry[slot] = sum;
                    }
                };
                auto staged = [&](auto coh) {
                    constexpr bool COH = decltype(coh)::value;
                    V xa[IPT];
#pragma unroll
                    for (int j = 0; j < IPT; ++j) xa[j] = gather_hint<V, COH>(a.x + pc[min(tid + j * THREADS, nz - 1)], pol_x);
                    if (rows == 0) {
                        V part = 0;
#pragma unroll
                        for (int j = 0; j < IPT; ++j)
                            if (tid + j * THREADS < nz) part += pv[tid + j * THREADS] * xa[j];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                        if (lane == 0) s_hub[hp][0][warp] = part;
                    } else {
#pragma unroll
                        for (int j = 0; j < IPT; ++j)
                            if (tid + j * THREADS < nz) pv[tid + j * THREADS] *= xa[j];
                        // queue the long segments while the gathers are in flight
                        for (int i = tid; i <= rows; i += THREADS) {
                            const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                            const int end = (i == rows) ? nz : s_re[i] - y0;
                            const int len = end - beg;
                            if (len > kWarpRowMax) s_hq[qc][atomicAdd(&s_nhq[qc], 1) & (kHugeCap - 1)] = i;
                            else if (len > a.med_lo) s_long[qc][atomicAdd(&s_nlong[qc], 1) & (kLongCap - 1)] = i;
                        }
                    }
                };
                if (tile_halo) staged(std::true_type{}); else staged(std::false_type{});
                consumer_sync<THREADS>();   // products and queues complete
                if (rows == 0) {
                    if (warp == 0) {
                        const V total = warp_sum_parts<V>(s_hub[hp][0], NW, lane);
                        if (lane == 0) s_tcarry[slot] = total;
                    }
                } else {
                    // huge segments first: their per-warp parts must be complete at the tile's second barrier
                    const bool has_huge = tile_ml > kWarpRowMax;   // uniform over the CTA
                    const int nh = has_huge ? min(s_nhq[qc], kHugeCap) : 0;
                    for (int h = 0; h < nh; ++h) {
                        const int i = s_hq[qc][h];
                        const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                        const int end = (i == rows) ? nz : s_re[i] - y0;
                        V part = seg_sum_strided<V, THREADS>(pv, beg, end, tid);
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                        if (lane == 0) s_hub[hp][h][warp] = part;
                    }
                    for (int i = tid; i <= rows; i += THREADS) {
                        const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                        const int end = (i == rows) ? nz : s_re[i] - y0;
                        if (end - beg <= a.med_lo) emit(i, seg_sum<V>(pv, beg, end));
                    }
                    const int nl = min(s_nlong[qc], kLongCap);
                    for (;;) {
                        int idx = 0;
                        if (lane == 0) idx = atomicAdd(&s_next[qc], 1);
                        idx = __shfl_sync(0xffffffffu, idx, 0);
                        if (idx >= nl) break;
                        const int i = s_long[qc][idx];
                        const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                        const int end = (i == rows) ? nz : s_re[i] - y0;
                        V part = seg_sum_strided<V, 32>(pv, beg, end, lane);
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                        if (lane == 0) emit(i, part);
                    }
                    if (has_huge) {
                        consumer_sync<THREADS>();   // per-warp parts of the huge segments complete
                        if (warp < nh) {
                            const V total = warp_sum_parts<V>(s_hub[hp][warp], NW, lane);
                            if (lane == 0) emit(s_hq[qc][warp], total);
                        }
                    }
                }
            } else {
                // ---- general tile, per-row gathers (SMLE_SPMV_DEBUG=4; the default before the product-staged path)
                //   <= kRowPathMaxLen  one thread per row segment (as above)
                //   <= kWarpRowMax     the warp that found it reduces it: lanes stride over its nonzeros,
                //                      shuffle tree; warp-private, no CTA barrier
                //   longer             (a hub row: at most two such segments fit in a tile) the whole CTA
                //                      strides over it; the only barriers of the kernel's data path
                auto emit = [&](int i, V sum) {
                    if (i < rows) {
                        a.y[x0 + i] = sum;
                        if constexpr (DOT) dot += sum * __ldg(a.x + x0 + i);
                    } else {
                        s_tcarry[slot] = sum;
                    }
                };
                // Segments of 33..kWarpRowMax are not reduced by the warp that found them but queued and dealt to
                // the warps dynamically after one barrier: on power-law matrices a few such rows per tile
                // otherwise keep one warp busy while 14 wait for the stage to drain (R-MAT scale 22 / 23:
                // 1007 -> 628 us, 2009 -> 1377 us; profiles/r02_spmv_general_tile_balance_ab.txt; debug_flags
                // bit 1 switches back to the warp-private reduction).  Queue counters are triple-buffered over
                // the general tiles of this CTA: the set for the NEXT general tile is re-armed before this
                // tile's barrier, when its last users are two barriers behind.
                const bool balance = (a.debug_flags & 2) == 0;
                const int qc = gen_count % 3, qn = (gen_count + 1) % 3;
                ++gen_count;
                if (balance && tid == 0) { s_nlong[qn] = 0; s_next[qn] = 0; }
                auto tiers = [&](auto coh) {
                    constexpr bool COH = decltype(coh)::value;
                    for (int base = 0; base <= rows; base += THREADS) {   // warp-uniform trip count (ballots inside)
                        const int i = base + tid;
                        const bool valid = i <= rows;
                        int beg = 0, end = 0;
                        if (valid) {
                            beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                            end = (i == rows) ? nz : s_re[i] - y0;
                        }
                        const int len = end - beg;
                        V sum = 0;
                        if (valid && len <= a.med_lo) sum = row_sum<V, COH>(a.x, pc, pv, beg, end);
                        const bool is_med = valid && len > a.med_lo && len <= kWarpRowMax;
                        if (balance && is_med) s_long[qc][atomicAdd(&s_nlong[qc], 1) & (kLongCap - 1)] = i;
                        unsigned med = __ballot_sync(0xffffffffu, is_med && !balance);
                        while (med) {
                            const int src = __ffs(med) - 1;
                            med &= med - 1;
                            const int b2 = __shfl_sync(0xffffffffu, beg, src), e2 = __shfl_sync(0xffffffffu, end, src);
                            V part = strided_sum<V, COH, 32>(a.x, pc, pv, b2, e2, lane);
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                            if (lane == src) sum = part;
                        }
                        if (valid && len > kWarpRowMax) s_huge[atomicAdd(&s_nhuge, 1) & (kHugeCap - 1)] = i;
                        else if (valid && !(balance && is_med)) emit(i, sum);
                    }
                    if (balance || tile_ml > kWarpRowMax) consumer_sync<THREADS>();   // queues complete (uniform over the CTA)
                    if (balance) {
                        const int nl = min(s_nlong[qc], kLongCap);
                        for (;;) {
                            int idx = 0;
                            if (lane == 0) idx = atomicAdd(&s_next[qc], 1);
                            idx = __shfl_sync(0xffffffffu, idx, 0);
                            if (idx >= nl) break;
                            const int i = s_long[qc][idx];
                            const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                            const int end = (i == rows) ? nz : s_re[i] - y0;
                            V part = strided_sum<V, COH, 32>(a.x, pc, pv, beg, end, lane);
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                            if (lane == 0) emit(i, part);
                        }
                    }
                    if (tile_ml > kWarpRowMax) {   // the tile's longest segment is known: uniform over the CTA
                        const int nh = min(s_nhuge, kHugeCap);
                        for (int h = 0; h < nh; ++h) {
                            const int i = s_huge[h];
                            const int beg = (i == 0) ? 0 : s_re[i - 1] - y0;
                            const int end = (i == rows) ? nz : s_re[i] - y0;
                            V part = strided_sum<V, COH, THREADS>(a.x, pc, pv, beg, end, tid);
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                            if (lane == 0) s_wsum[warp] = part;
                            consumer_sync<THREADS>();
                            if (tid == 0) {
                                V total = 0;
                                for (int wi = 0; wi < NW; ++wi) total += s_wsum[wi];
                                emit(i, total);
                                s_nhuge = 0;   // every thread has read it; the barrier below orders the reset
                            }
                            consumer_sync<THREADS>();
                        }
                    }
                };
                if (tile_halo) tiers(std::true_type{}); else tiers(std::false_type{});
            }

            // release the stage: generic-proxy accesses ordered before the next bulk copy, one
            // arrival per consumer warp
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);

            // ---- carry chain: every kChainTiles tiles (and after the last one) ------------------------
            if (slot == kChainTiles - 1 || t == t1 - 1) {
                consumer_sync<THREADS>();
                const int cnt = slot + 1;
                if (tid == 0) {
                    V run = s_running;
                    for (int j = 0; j < cnt; ++j) {
                        const V out = s_tcarry[j];
                        if (s_trow[j] >= 0) { s_tcarry[j] = run; run = out; }   // run completes row s_trow[j]
                        else run += out;                                        // row continues through tile j
                    }
                    s_running = run;
                }
                consumer_sync<THREADS>();
                for (int j = tid; j < cnt; j += THREADS) {
                    const int row = s_trow[j];
                    const V add = s_tcarry[j];
                    if (row >= 0 && add != V(0)) {
                        a.y[row] = __ldcg(a.y + row) + add;
                        if constexpr (DOT) dot += add * __ldg(a.x + row);
                    }
                }
                consumer_sync<THREADS>();
            }
        }
    }

    __syncthreads();   // all tiles done (producer warp included)

    // ---- rows cut by the CTA boundaries: wait-free exchange through one slot per boundary --------
    // (merge_based.hpp:137-149: the carry is added to what the row's owner stored.)  The dot product
    // is linear in the parts of a row, so every CTA adds carry-out * x[row] to its own partial and no
    // cross-CTA fix-up of p.Ap is needed.
    V dot_carry = 0;
    if (tid == 0 && t1 > t0 && !(a.debug_flags & 1)) {
        const int c = blockIdx.x;
        const int r0 = s_edge[0], r1 = s_edge[1], r2 = s_edge[2];   // rows in progress at the start / end of this CTA, end of the next
        const bool has_in = c > 0 && r0 < a.m;
        const V out = s_running;                                  // leading part of row r1 seen by this CTA
        if constexpr (DOT) { if (r1 < a.m) dot_carry = out * __ldg(a.x + r1); }
        __threadfence();   // this CTA's stores to y[r0] are ordered before its slot operations
        if (r1 > r0) {
            if (has_in) {  // this CTA owns the end of row r0: meet the carry of CTA c-1
                const V mine = __ldcg(a.y + r0);
                const V other = slot_exchange<V>(a.cta_slot + (c - 1), mine);
                if (!is_sentinel<V>(other)) { slot_reset<V>(a.cta_slot + (c - 1)); a.y[r0] = mine + other; }
            }
            cta_carry_publish<V>(a.tile_xy, a.tiles_per_cta, a.num_tiles, a.m, a.cta_slot, a.y, c, out, r1, r2);
        } else if (has_in) {   // the whole CTA lies inside row r0 == r1: pass the carry on
            const V other = slot_exchange<V>(a.cta_slot + (c - 1), out);
            if (!is_sentinel<V>(other)) { slot_reset<V>(a.cta_slot + (c - 1)); cta_carry_publish<V>(a.tile_xy, a.tiles_per_cta, a.num_tiles, a.m, a.cta_slot, a.y, c, other + out, r1, r2); }
        } else {
            cta_carry_publish<V>(a.tile_xy, a.tiles_per_cta, a.num_tiles, a.m, a.cta_slot, a.y, c, out, r1, r2);
        }
    }

    if constexpr (!DOT) return;

    if constexpr (DOT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        if (lane == 0 && warp < NW) s_wsum[warp] = dot;
        __syncthreads();
        if (tid == 0) {
            V sdot = dot_carry;
            for (int wi = 0; wi < NW; ++wi) sdot += s_wsum[wi];
            a.dot_part[blockIdx.x] = sdot;
        }

        // ---- last CTA done: p.Ap from the per-CTA partials in CTA order (deterministic) -------------
        if (!last_cta_election(a.ticket, gridDim.x)) return;
        const V pAp = cta_reduce_one<V>(a.dot_part, nullptr, gridDim.x, s_red);
        if (a.dist) {
            // this rank's partial -> every rank's mailbox; K2 adds the G partials in rank order
            if (tid == 0) cg.pAp[0] = (double)pAp;
            dist_post(*a.dist, 0, (double)pAp, cg.ctrl);
        } else if (tid == 0) {
            cg_dot_scalars(cg, 0, (double)pAp);
        }
    }
}

} // namespace smle
