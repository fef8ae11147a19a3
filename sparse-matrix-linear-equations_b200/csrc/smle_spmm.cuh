// smle_spmm.cuh -- TMA-staged tall-skinny SpMM (Y = A X, X n x k row-major), row-per-worker.
//
// Replaces OmpMergeCsrmm (reference work_2025/spmm/merge_based.hpp:49-153) for matrices without
// very long rows (every stencil, most FEM matrices); skewed matrices (wheel, R-MAT hubs) keep the
// merge-walk kernel of smle_merge.cuh.  The decomposition is still the reference's merge path:
// tiles are equal shares of (row ends + nonzeros), a tile owns the rows that END inside it plus
// the leading part of the row it stops in, whose partial sum is a carry-out.
//
// B200 mapping:
//   * persistent CTAs, a producer warp streaming column indices / values / row
//     offsets of each tile through shared memory with cp.async.bulk + mbarriers (L2 evict-first:
//     A is read once per product), consumer warps that never meet at a CTA-wide barrier;
//   * tiles are dealt to CTAs ROUND-ROBIN in chunks of `chunk` consecutive tiles, so at any
//     moment the whole GPU works on one compact window of rows.  The dense block does not fit in
//     L2 (2 GB at 200^3 x 32), but the window plus the off-diagonal planes it touches does, so
//     every dense row comes from DRAM once instead of once per stencil plane;
//   * when the matrix has a dominant far stride D (the w^2 plane offset of a 3-D stencil; detected
//     from a sample of rows when the handle first runs an SpMM), the deal becomes a SCHEDULE: a CTA
//     walks a chain of tiles D rows apart, so the dense rows it fetched as the +D neighbours of one
//     tile are the rows it owns in the next and the -D neighbours of the one after -- L1 hits
//     instead of two more trips to L2 per row.  Chains are cut into segments and dealt so that
//     the CTAs working at the same time still share one compact window of rows (DRAM traffic stays
//     compulsory) and finish together;
//   * a WORKER of G lanes owns one row at a time and covers G*VEC = min(k, 32*VEC) columns with
//     VEC-wide (128-bit) loads of the dense rows: every nonzero (broadcast from shared memory) is
//     reused across all k right-hand sides.  The loads of a pass are unconditional (what lies
//     behind a row in the staged indices is always a valid column), which lets ptxas pipeline
//     them: ~4 requests of 512 B in flight per warp at 64 registers, 30 warps per SM.  Workers sit
//     on W consecutive rows, so the near-diagonal gathers hit L1 (kept at 192 KB);
//   * carries without fix-up passes and without waiting: the two tiles that share a cut row meet
//     in a global slot (one per tile and column).  Each swaps its part in with an atomic
//     exchange; the slot holds a signalling-NaN bit pattern that no arithmetic result can have
//     while it is empty, so the party that finds a value there knows it came second, stores the
//     finished row element (owner part + carry, the reference's order, merge_based.hpp:146 --
//     the same sum whoever finishes) and re-arms the slot.  No fences (they would invalidate
//     L1), no per-launch reset, no dependence on how tiles are scheduled;
//   * DOT adds the per-column p.Ap of CG while rows are stored; the last CTA reduces the per-CTA
//     partials in CTA order (deterministic) and forms alpha.
#pragma once
#include "smle_spmv.cuh"

namespace smle {

// Structure probe for the tile schedule: column offsets (col - row) of the first kSampleNnz nonzeros
// of `nsamples` evenly spaced rows (INT_MIN where the row is shorter).  Host code looks for a far
// offset that most rows share (smle_capi.cu, far_stride()).
constexpr int kSampleNnz = 16;

__global__ void sample_offsets_kernel(const int *__restrict__ ro, const int *__restrict__ ci, int m, int nsamples,
                                      int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsamples) return;
    const int row = (int)((long long)i * m / nsamples);
    const int beg = ro[row], end = ro[row + 1];
    for (int j = 0; j < kSampleNnz; ++j) out[i * kSampleNnz + j] = (beg + j < end) ? ci[beg + j] - row : INT_MIN;
}

// L2-coherent (volatile) vector access to a carry slot: VEC*sizeof(V) in {4, 8, 16} bytes
template <typename V, int VEC>
__device__ __forceinline__ void ld_volatile_vec(V (&out)[VEC], const V *p)
{
    if constexpr (sizeof(V) * VEC == 16) {
        uint4 t;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p) : "memory");
        memcpy(out, &t, 16);
    } else if constexpr (sizeof(V) * VEC == 8) {
        uint2 t;
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(p) : "memory");
        memcpy(out, &t, 8);
    } else {
        unsigned int t;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(t) : "l"(p) : "memory");
        memcpy(out, &t, 4);
    }
}

template <typename V, int VEC>
__device__ __forceinline__ void st_volatile_vec(V *p, const V (&in)[VEC])
{
    if constexpr (sizeof(V) * VEC == 16) {
        uint4 t;
        memcpy(&t, in, 16);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w)
                     : "memory");
    } else if constexpr (sizeof(V) * VEC == 8) {
        uint2 t;
        memcpy(&t, in, 8);
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(t.x), "r"(t.y) : "memory");
    } else {
        unsigned int t;
        memcpy(&t, in, 4);
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(t) : "memory");
    }
}

// streaming store of an output row (read next by a different kernel, never by this one)
template <typename V, int VEC>
__device__ __forceinline__ void st_stream_vec(V *p, const V (&in)[VEC], uint64_t pol)
{
    if constexpr (sizeof(V) * VEC == 16) {
        uint4 t;
        memcpy(&t, in, 16);
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z),
                     "r"(t.w), "l"(pol)
                     : "memory");
    } else if constexpr (sizeof(V) * VEC == 8) {
        uint2 t;
        memcpy(&t, in, 8);
        asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(t.x), "r"(t.y), "l"(pol) : "memory");
    } else {
        unsigned int t;
        memcpy(&t, in, 4);
        asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(t), "l"(pol) : "memory");
    }
}

// stage layout: like SpmvSmem, with room behind the column indices for the over-read of one pass
template <typename V, int TILE>
struct SpmmSmem {
    static constexpr int EPV = 16 / (int)sizeof(V);
    static constexpr int COL_WORDS = TILE + 24;
    static constexpr int VAL_ELEMS = TILE + 2 * EPV;
    static constexpr int RO_WORDS = TILE + 8;
    static constexpr size_t HDR_OFFSET =
        ((size_t)COL_WORDS * 4 + (size_t)VAL_ELEMS * sizeof(V) + (size_t)RO_WORDS * 4 + 15) / 16 * 16;
    // tile header written by the producer: {tile id (-1: no more tiles), lo.x, lo.y, hi.x, hi.y}
    static constexpr size_t STAGE_BYTES = HDR_OFFSET + 32;
};

constexpr int kSpmmRot = 17;   // worker rotation per tile

template <typename V>
struct SpmmArgs {
    const int *__restrict__ ro;
    const int *__restrict__ ci;
    const V *__restrict__ va;
    const V *__restrict__ X;
    V *__restrict__ Y;
    const int2 *__restrict__ tile_xy;
    int m, nnz, k;
    int num_tiles;
    int chunk;                         // consecutive tiles a CTA takes before the deal moves on
    const int *sched;                  // structure-aware tile schedule: the tiles of CTA b in processing order start at
    const int *sched_off;              //   sched[sched_off[b]] and end with the sentinel num_tiles; NULL -> the deal above
    V *tile_carry;                     // [num_tiles * k] carry slots, sentinel when empty
    V *dot_part;                       // [gridDim.x * k]  (DOT)
    V *dot_sum;                        // [k] reduced dot products before they become the solver's scalars (DOT)
    unsigned int *ticket;
    int band;                          // RING > 0: half-width of the window around the tile's rows kept in the ring
    int y_policy;                      // 1: stream Y with L2 evict-first
    int dot_late;                      // DOT: load X[row,:] after the gathers (1, default) or before them (0).  Kept as a
                                       // run-time switch on purpose: with the early load compiled out ptxas keeps only 2
                                       // gathers in flight instead of 4 (1.61 ms vs 1.42 ms; tests/test_sass_shape.py)
};


// val = sum of the parts of row tile_xy[tt+1].x that lie in tiles <= tt, column c.
// Returns the row element's share of the dot product X[row,c]*Y[row,c] when it finished the row.
template <typename V, bool DOT>
__device__ __noinline__ V carry_publish(const int2 *__restrict__ tile_xy, V *tile_carry, V *Y, const V *X, int m, int k,
                                        int tt, int c, V val)
{
    for (;;) {
        const int row = tile_xy[tt + 1].x;
        if (row >= m) return V(0);
        V *slot = tile_carry + (size_t)tt * (size_t)k + c;
        const V owner = slot_exchange<V>(slot, val);
        if (is_sentinel<V>(owner)) return V(0);            // the owner tile comes later and finishes
        slot_reset<V>(slot);
        if (tile_xy[tt + 2].x > row) {                     // the row ends in tile tt+1
            const size_t off = (size_t)row * (size_t)k + c;
            const V fin = owner + val;
            Y[off] = fin;
            if constexpr (DOT) return fin * __ldg(X + off);
            return V(0);
        }
        val = val + owner;                                 // tile tt+1 lies inside the row: pass on
        ++tt;
    }
}

// Template parameters
//   G, VEC, NV  a worker is G lanes; a lane holds NV vectors of VEC consecutive columns (VEC*sizeof(V)
//               <= 16 B), vector q at column cb*KB + q*G*VEC + li*VEC, so every load instruction of a
//               worker covers G*VEC*sizeof(V) contiguous bytes and a worker covers KB = G*VEC*NV columns
//   UB          nonzeros of a row whose dense rows are requested before the first FMA
//   THREADS     consumer threads (+ one producer warp);  TILE merge items per tile;  STAGES tiles
//               in flight;  MINB CTAs per SM
//   DOT         also accumulate X[r,:].Y[r,:] (p.Ap of CG); needs k <= KB (one column block)
//   RING        > 0: band-window variant.  The dense rows [first row - band, last row + band] of the tile
//               being processed live in a shared-memory ring of RING rows that the producer extends with one
//               bulk copy per tile (consecutive tiles of a chunk overlap in all but the new rows), so the
//               gathers of columns inside the band -- self, +-1, +-w of a stencil -- are shared-memory
//               reads and every dense row crosses L2 -> SM once per chunk instead of once per use.
//               Columns outside the window (+-w^2) keep the L1 path.  Needs k == KB (one column block).
template <typename V, int G, int VEC, int NV, int UB, int THREADS, int TILE, int STAGES, int MINB, bool DOT, int RING = 0>
__global__ void __launch_bounds__(THREADS + 32, MINB)
spmm_rows_kernel(SpmmArgs<V> a, CgScalars cg)
{
    using SM = SpmmSmem<V, TILE>;
    constexpr bool BAND = RING > 0;
    static_assert(UB <= 16, "over-read room behind the staged column indices");
    static_assert(!BAND || NV == 1, "the band-window variant covers one column block with one vector per lane");
    static_assert(THREADS % 32 == 0 && THREADS % G == 0 && G <= 32, "workers tile the consumer warps");
    constexpr int NW = THREADS / 32;
    constexpr int EPV = SM::EPV;
    constexpr int W = THREADS / G;         // workers per CTA
    constexpr int HB = G * VEC;            // columns per load instruction of a worker
    constexpr int KB = HB * NV;            // columns per column block

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_full[STAGES], s_empty[STAGES];
    // the reduction scratch of the DOT epilogue reuses the stage buffers (all tiles are consumed by
    // then): static shared memory stays at a few bytes, so that 2 stages of 1920 items fit the 64 KB
    // carve-out and L1 keeps 192 KB for the dense-row gathers
    static_assert(!DOT || (size_t)(NW * KB + THREADS + 32) * sizeof(V) <= SM::STAGE_BYTES * STAGES, "epilogue scratch");
    V(*s_wsum)[KB] = reinterpret_cast<V(*)[KB]>(smem_raw);
    V *s_red = reinterpret_cast<V *>(smem_raw) + NW * KB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = tid / G, li = tid % G;
    const size_t k = (size_t)a.k;

    if constexpr (DOT) {
        if (cg.ctrl[CTRL_STOP]) {      // see spmv_kernel: STOP is two launches old, HALT waits for K3
            griddep_wait();
            if (blockIdx.x == 0 && tid == 0) cg.ctrl[CTRL_HALT] = 1;
            return;
        }
    }

    // tile schedule: iteration `it` of CTA b -> tile ((it / chunk) * grid + b) * chunk + it % chunk
    const int chunk = a.chunk;
    const int stride = (int)gridDim.x * chunk;
    // (every CTA's list ends with the sentinel num_tiles, so no count has to stay live in a register)
    const int my_base = a.sched ? a.sched_off[blockIdx.x] : 0;
    auto tile_of = [&](int it) {
        if (a.sched) return __ldg(a.sched + my_base + it);
        return (it / chunk) * stride + (int)blockIdx.x * chunk + it % chunk;
    };

    auto stage_col = [&](int s) { return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES); };
    auto stage_val = [&](int s) {
        return reinterpret_cast<V *>(smem_raw + (size_t)s * SM::STAGE_BYTES + (size_t)SM::COL_WORDS * 4);
    };
    auto stage_ro = [&](int s) {
        return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES + (size_t)SM::COL_WORDS * 4 +
                                       (size_t)SM::VAL_ELEMS * sizeof(V));
    };
    auto stage_hdr = [&](int s) { return reinterpret_cast<int *>(smem_raw + (size_t)s * SM::STAGE_BYTES + SM::HDR_OFFSET); };
    // band-window variant: the ring of dense rows sits behind the stages; row g of the current chunk is in
    // slot (g - chunk origin) mod RING, KB values per slot
    [[maybe_unused]] V *ring = reinterpret_cast<V *>(smem_raw + SM::STAGE_BYTES * STAGES);
    constexpr unsigned kRingRowBytes = (unsigned)KB * (unsigned)sizeof(V);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // The gathers of a pass read UB column indices without looking at the row end; what lies behind
    // a tile's staged indices must therefore always be a valid column: zero now, indices of earlier
    // tiles later.
    for (int s = 0; s < STAGES; ++s)
        for (int i = tid; i < SM::COL_WORDS; i += blockDim.x) stage_col(s)[i] = 0;
    fence_proxy_async();
    __syncthreads();

    V dot[NV][VEC];
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int v = 0; v < VEC; ++v) dot[q][v] = 0;

    if (warp == NW) {
        // =============================== producer warp ===========================================
        if (lane == 0) {
            const uint64_t pol_stream = l2_policy_evict_first();
            [[maybe_unused]] const uint64_t pol_keep = l2_policy_evict_last();
            [[maybe_unused]] int prev_t = -2, ring_hi = 0, origin = 0;   // band window: last tile, rows loaded so far, chunk origin
            for (int it = 0;; ++it) {
                const int t = tile_of(it);
                const int s = it % STAGES;
                if (it >= STAGES) {
                    mbar_wait(&s_empty[s], (uint32_t)(it / STAGES - 1) & 1u);
                    fence_proxy_async();
                }
                if constexpr (BAND) {
                    // a new chunk reloads the window from slot 0: the tile before it (one stage back) must have
                    // been consumed as well, or its rows would be overwritten under the consumers
                    if (it >= 1 && t != prev_t + 1 && t < a.num_tiles) {
                        mbar_wait(&s_empty[(it - 1) % STAGES], (uint32_t)((it - 1) / STAGES) & 1u);
                        fence_proxy_async();
                    }
                }
                // The tile's identity travels with its data: the consumers read the header after the wait
                // on "full" (the arrive below releases these stores) and never touch the schedule or the
                // coordinate array themselves -- no global loads, no registers held across the tile.
                int *hdr = stage_hdr(s);
                if (t >= a.num_tiles) {
                    hdr[0] = -1;
                    mbar_arrive(&s_full[s]);
                    break;
                }
                const int2 lo = a.tile_xy[t], hi = a.tile_xy[t + 1];
                hdr[0] = t; hdr[1] = lo.x; hdr[2] = lo.y; hdr[3] = hi.x; hdr[4] = hi.y;
                const int yc = lo.y & ~3, yv = lo.y & ~(EPV - 1), rb = (lo.x + 1) & ~3;
                const uint32_t nb_col = (uint32_t)((hi.y - yc + 3) & ~3) * 4u;
                const uint32_t nb_val = (uint32_t)((hi.y - yv + EPV - 1) & ~(EPV - 1)) * (uint32_t)sizeof(V);
                const uint32_t nb_ro = (uint32_t)((hi.x + 2 - rb + 3) & ~3) * 4u;
                if constexpr (BAND) {
                    // window of this tile: rows [first - band, last + band]; consecutive tiles only add the new rows
                    // A tile and its successor must both find their windows in the ring: the half-width is
                    // clamped by the rows of this tile together with either neighbour (a few tiles on the
                    // faces of a grid hold more, shorter rows; their +-w then take the far path)
                    const int r_prev = hi.x - a.tile_xy[max(t - 1, 0)].x, r_next = a.tile_xy[min(t + 2, a.num_tiles)].x - lo.x;
                    const int half = max(min(a.band, (RING - 2 - max(r_prev, r_next)) / 2), 0);
                    const int last = min(hi.x, a.m - 1);
                    const int wlo = max(lo.x - half, 0), whi = min(last + half + 1, a.m);
                    int ld_lo;
                    if (t != prev_t + 1) { origin = wlo; ld_lo = wlo; }
                    else ld_lo = ring_hi;
                    const int ld_hi = max(whi, ld_lo);
                    prev_t = t; ring_hi = ld_hi;
                    hdr[5] = origin + ((wlo - origin) / RING) * RING;   // subtract this, then wrap once: ring slot of a row
                    hdr[6] = wlo; hdr[7] = whi;
                    const int len = ld_hi - ld_lo, s0 = (ld_lo - origin) % RING;
                    const int len0 = min(len, RING - s0), len1 = len - len0;
                    mbar_expect_tx(&s_full[s], nb_col + nb_val + nb_ro + (uint32_t)len * kRingRowBytes);
                    if (len0 > 0) tma_load_1d(ring + (size_t)s0 * KB, a.X + (size_t)ld_lo * KB, (uint32_t)len0 * kRingRowBytes, &s_full[s], pol_keep);
                    if (len1 > 0) tma_load_1d(ring, a.X + (size_t)(ld_lo + len0) * KB, (uint32_t)len1 * kRingRowBytes, &s_full[s], pol_keep);
                } else {
                    mbar_expect_tx(&s_full[s], nb_col + nb_val + nb_ro);
                }
                if (nb_col) tma_load_1d(stage_col(s), a.ci + yc, nb_col, &s_full[s], pol_stream);
                if (nb_val) tma_load_1d(stage_val(s), a.va + yv, nb_val, &s_full[s], pol_stream);
                tma_load_1d(stage_ro(s), a.ro + rb, nb_ro, &s_full[s], pol_stream);
            }
        }
    } else {
        // =============================== consumer warps ==========================================
        if constexpr (DOT) { griddep_wait(); griddep_launch_dependents(); }   // the producer already streams A
        const uint64_t pol_y = l2_policy_evict_first();
        const int num_cb = DOT ? 1 : (a.k + KB - 1) / KB;
        const unsigned kbytes = (unsigned)a.k * (unsigned)sizeof(V);   // bytes per dense row
        for (int it = 0;; ++it) {
            const int s = it % STAGES;
            const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&s_full[s], parity);
            const int *hdr = stage_hdr(s);
            const int t = hdr[0];
            if (t < 0) break;
            const int2 lo = make_int2(hdr[1], hdr[2]), hi = make_int2(hdr[3], hdr[4]);
            const int x0 = lo.x, y0 = lo.y;
            const int rows = hi.x - x0, nz = hi.y - y0;
            const int yc = y0 & ~3, yv = y0 & ~(EPV - 1), rb = (x0 + 1) & ~3;
            const int *s_re = stage_ro(s) + (x0 + 1 - rb);   // s_re[i] = end offset of local row i
            const int *pc = stage_col(s) + (y0 - yc);
            const V *pv = stage_val(s) + (y0 - yv);
            // Carries: pseudo-row `rows` is the part of row hi.x that lies in this tile (it continues
            // in tile t+1); local row 0 may have begun in tile t-1.  Both sides meet in slot t-1 / t.
            const bool has_in = t > 0 && x0 < a.m;
            [[maybe_unused]] int ring_sub = 0, wlo = 0, whi = 0;
            if constexpr (BAND) { ring_sub = hdr[5]; wlo = hdr[6]; whi = hdr[7]; }
            // ring slot of dense row g (wlo <= g < whi): g - ring_sub, wrapped once
            [[maybe_unused]] auto ring_row = [&](int g) {
                int d = g - ring_sub;
                d -= (d >= RING) ? RING : 0;
                return reinterpret_cast<const char *>(ring) + (size_t)(unsigned)d * kRingRowBytes;
            };

            // rotate the worker -> row map from tile to tile: with rows % W != 0 the same workers would
            // otherwise take the extra row of every tile
            const int wrot = (w + it * kSpmmRot) % W;

            for (int cb = 0; cb < num_cb; ++cb) {
                // columns of this lane; lanes past k read column block 0 instead and store nothing
                int coff[NV];
                bool ok[NV];
                const char *xlane[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    coff[q] = cb * KB + q * HB + li * VEC;
                    ok[q] = coff[q] < a.k;
                    if (!ok[q]) coff[q] = 0;
                    xlane[q] = reinterpret_cast<const char *>(a.X + coff[q]);
                }
                for (int i = wrot; i <= rows; i += W) {
                    const int beg0 = (i == 0) ? 0 : s_re[i - 1] - y0;
                    const int end = (i == rows) ? nz : s_re[i] - y0;
                    const bool is_out = (i == rows), is_in = (i == 0) && has_in;
                    if (is_out && hi.x >= a.m) continue;   // behind the last row: nothing to produce
                    V acc[NV][VEC];
#pragma unroll
                    for (int q = 0; q < NV; ++q)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[q][v] = 0;
                    V xr[NV][VEC];
                    if constexpr (DOT && !BAND) {
                        if (!is_out && !a.dot_late) {
#pragma unroll
                            for (int q = 0; q < NV; ++q)
                                ldg_vec<V, VEC>(xr[q], reinterpret_cast<const V *>(xlane[q] + (size_t)(unsigned)(x0 + i) * kbytes));
                        }
                    }
                    if constexpr (BAND) {
                        // columns are sorted: the nonzeros inside the window are one contiguous run [n0, n1)
                        int n0 = beg0, n1 = end;
                        while (n0 < end && pc[n0] < wlo) ++n0;
                        while (n1 > n0 && pc[n1 - 1] >= whi) --n1;
                        const int nlow = n0 - beg0, nfar = nlow + (end - n1);
                        auto far_index = [&](int j) { return j < nlow ? beg0 + j : n1 + (j - nlow); };
                        const unsigned loff = (unsigned)(li * VEC) * (unsigned)sizeof(V);   // (k == KB: coff[0] = li*VEC)
                        // far columns first: UB gathers in flight while the band is served from shared memory
                        V xv[UB][VEC];
                        int zf[UB];
                        if (nfar > 0) {
#pragma unroll
                            for (int u = 0; u < UB; ++u) {
                                zf[u] = far_index(min(u, nfar - 1));
                                ldg_vec<V, VEC>(xv[u], reinterpret_cast<const V *>(xlane[0] + (size_t)(unsigned)pc[zf[u]] * kbytes));
                            }
                        }
                        for (int z = n0; z < n1; ++z) {
                            V xs[VEC];
                            ld_vec<V, VEC>(xs, reinterpret_cast<const V *>(ring_row(pc[z]) + loff));
                            const V av = pv[z];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[0][v] += av * xs[v];
                        }
                        if (nfar > 0) {
#pragma unroll
                            for (int u = 0; u < UB; ++u)
                                if (u < nfar) {
                                    const V av = pv[zf[u]];
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) acc[0][v] += av * xv[u][v];
                                }
                            for (int j = UB; j < nfar; ++j) {   // rows with more than UB far columns (rare)
                                const int z = far_index(j);
                                V xg[VEC];
                                ldg_vec<V, VEC>(xg, reinterpret_cast<const V *>(xlane[0] + (size_t)(unsigned)pc[z] * kbytes));
                                const V av = pv[z];
#pragma unroll
                                for (int v = 0; v < VEC; ++v) acc[0][v] += av * xg[v];
                            }
                        }
                        if constexpr (DOT) {
                            if (!is_out) ld_vec<V, VEC>(xr[0], reinterpret_cast<const V *>(ring_row(x0 + i) + loff));
                        }
                    } else
                    for (int beg = beg0; beg < end; beg += UB) {
                        // UB dense rows requested per pass; ptxas keeps about as many loads in flight per
                        // warp as it has scoreboards, the rest of the latency is hidden by the other warps
                        const int *pcb = pc + beg;
                        const V *pvb = pv + beg;
                        const int cnt = end - beg;
                        V xv[UB][NV][VEC];
#pragma unroll
                        for (int u = 0; u < UB; ++u) {      // unconditional: slots behind the row read valid columns
                            const size_t xrow = (size_t)(unsigned)pcb[u] * kbytes;
#pragma unroll
                            for (int q = 0; q < NV; ++q)
                                ldg_vec<V, VEC>(xv[u][q], reinterpret_cast<const V *>(xlane[q] + xrow));
                        }
#pragma unroll
                        for (int u = 0; u < UB; ++u)
                            if (u < cnt) {
                                const V av = pvb[u];
#pragma unroll
                                for (int q = 0; q < NV; ++q)
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) acc[q][v] += av * xv[u][q][v];
                            }
                    }
                    if constexpr (DOT && !BAND) {
                        // the row's own dense row was just gathered for the diagonal entry: an L1 hit now,
                        // and it did not occupy one of the few load slots while the gathers were in flight
                        if (!is_out && a.dot_late) {
#pragma unroll
                            for (int q = 0; q < NV; ++q)
                                ldg_vec<V, VEC>(xr[q], reinterpret_cast<const V *>(xlane[q] + (size_t)(unsigned)(x0 + i) * kbytes));
                        }
                    }
                    if (is_in || is_out) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) {
                            if (!ok[q]) continue;
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                const int c = coff[q] + v;
                                V d = 0;
                                if (is_in) {
                                    // this row began in tile t-1: meet its carry in slot t-1
                                    V *slot = a.tile_carry + (size_t)(t - 1) * k + c;
                                    const V other = slot_exchange<V>(slot, acc[q][v]);
                                    if (!is_sentinel<V>(other)) {      // the carry was there first: finish here
                                        slot_reset<V>(slot);
                                        if (!is_out) {
                                            const V fin = acc[q][v] + other;
                                            a.Y[(size_t)(x0 + i) * k + c] = fin;
                                            if constexpr (DOT) d = fin * xr[q][v];
                                        } else {                       // no row ends in this tile: pass on
                                            d = carry_publish<V, DOT>(a.tile_xy, a.tile_carry, a.Y, a.X, a.m, a.k, t, c,
                                                                      other + acc[q][v]);
                                        }
                                    }
                                } else {
                                    d = carry_publish<V, DOT>(a.tile_xy, a.tile_carry, a.Y, a.X, a.m, a.k, t, c, acc[q][v]);
                                }
                                if constexpr (DOT) dot[q][v] += d;
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < NV; ++q) {
                            if (!ok[q]) continue;
                            V *dst = a.Y + (size_t)(x0 + i) * k + coff[q];
                            if (a.y_policy) st_stream_vec<V, VEC>(dst, acc[q], pol_y);
                            else st_vec<V, VEC>(dst, acc[q]);
                            if constexpr (DOT) {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) dot[q][v] += acc[q][v] * xr[q][v];
                            }
                        }
                    }
                }
            }

            // release the stage: one arrival per consumer warp
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
        }
    }

    if constexpr (!DOT) return;

    // ---- per-CTA dot partials, last CTA: pAp and alpha (no_pretreatment.hpp:107-120) --------------
    if constexpr (DOT) {
#pragma unroll
        for (int d = G; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < NV; ++q)
#pragma unroll
                for (int v = 0; v < VEC; ++v) dot[q][v] += __shfl_xor_sync(0xffffffffu, dot[q][v], d);
        }
        __syncthreads();
        if (warp < NW && lane / G == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q)
#pragma unroll
                for (int v = 0; v < VEC; ++v) s_wsum[warp][q * HB + li * VEC + v] = dot[q][v];
        }
        __syncthreads();
        if (tid < KB && tid < a.k) {
            V sdot = 0;
            for (int wi = 0; wi < NW; ++wi) sdot += s_wsum[wi][tid];
            a.dot_part[(size_t)blockIdx.x * k + tid] = sdot;
        }
        if (!last_cta_election(a.ticket, gridDim.x)) return;
        cta_reduce_columns<V>(a.dot_part, nullptr, gridDim.x, a.k, a.dot_sum, s_red);   // (V-typed scratch: the scalars are double)
        for (int c = tid; c < a.k; c += blockDim.x)
            cg_dot_scalars(cg, c, (double)a.dot_sum[c]);
    }
}

} // namespace smle
