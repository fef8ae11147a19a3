// smle_merge.cuh -- merge-path CSR SpMV / SpMM kernel for sm_100a.
//
// Replaces OmpMergeCsrmv (reference cpu_spmv.cpp:360-421) and OmpMergeCsrmm (reference
// work_2025/spmm/merge_based.hpp:49-153).  The decomposition is the reference's: the merge
// path over (row end-offsets, nonzero indices) is cut into equal shares of merge items, each
// share consumes whole rows, then a trailing partial row whose sum becomes a carry-out that a
// fix-up pass adds to the row's owner.  What is B200-specific is how the shares are mapped:
//
//   level 1  CTA      : a contiguous run of `tiles_per_cta` tiles (grid sized to the number of
//                       resident CTAs, so per-CTA carry-outs stay in the hundreds);
//   level 2  tile     : TILE = W*IPW merge items whose (row, nnz) start coordinates were
//                       computed once per matrix by merge_partition_kernel; its row-end
//                       offsets, column indices and values are staged in shared memory with
//                       coalesced loads;
//   level 3  worker   : G lanes (G*VEC = columns of the dense block handled at once) walk IPW
//                       merge items sequentially; every nonzero is broadcast from shared
//                       memory to the G lanes and multiplied with a 128-bit (VEC*sizeof(V))
//                       coalesced load of the dense row, so a nonzero is reused across all
//                       k right-hand sides.
//
// Carry-outs: worker -> (warp-shuffle segmented scan keyed by row) -> tile -> (shared memory)
// -> CTA -> (global, one entry per CTA) -> last CTA done applies them in CTA order, like the
// reference's serial fix-up loop (merge_based.hpp:137-149).  No atomics on the data path:
// results are deterministic run to run.
//
// DOT = true additionally accumulates the per-column dot products X[:,c] . Y[:,c] (the
// p.Ap of CG, no_pretreatment.hpp:107) while rows are emitted, and the last CTA turns them
// into alpha = rs_old / pAp (:109-120), saving two passes over the n x k blocks.
#pragma once
#include "smle_common.cuh"

namespace smle {

template <typename V>
struct MergeArgs {
    const int *__restrict__ row_end;   // row_offsets + 1 (m entries): merge list A
    const int *__restrict__ ci;        // column indices
    const V *__restrict__ va;          // values
    const V *__restrict__ X;           // n x k row-major
    V *__restrict__ Y;                 // m x k row-major
    const int2 *__restrict__ tile_xy;  // num_tiles + 1 merge-path coordinates (row, nnz)
    int m, nnz, k;
    int num_tiles, tiles_per_cta;
    int *carry_row;                    // [gridDim.x]   row each CTA stopped in
    V *carry_val;                      // [gridDim.x*k] its partial sum
    V *dot_part;                       // [gridDim.x*k] per-CTA dot partials        (DOT)
    V *fix_part;                       // [gridDim.x*k] dot share of the carry fix  (DOT)
    V *dot_sum;                        // [k] reduced dot products before they become the solver's scalars (DOT)
    unsigned int *ticket;
};

// ---------------------------------------------------------------------------------------
// merge_partition_kernel: one thread per share boundary runs the reference's 2-D diagonal
// binary search (merge_based.hpp:22-44) on list A = row end-offsets, list B = 0,1,2,...
// Must be bit-exact with the CPU function; tests compare every coordinate.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int2 merge_path_search(int diagonal, const int *__restrict__ a, int a_len,
                                                  int b_len)
{
    int lo = max(diagonal - b_len, 0);
    int hi = min(diagonal, a_len);
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= diagonal - mid - 1) lo = mid + 1; else hi = mid;
    }
    return make_int2(min(lo, a_len), diagonal - lo);
}

__global__ void merge_partition_kernel(const int *__restrict__ row_end, int m, int nnz,
                                       int items_per_part, int num_parts, int2 *__restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > num_parts) return;
    long long d = (long long)items_per_part * t;
    int total = m + nnz;
    int diag = d < total ? (int)d : total;
    out[t] = merge_path_search(diag, row_end, m, nnz);
}

// ---------------------------------------------------------------------------------------
// merge_kernel
//   G    lanes per worker (power of two, 1..32)
//   VEC  consecutive columns per lane (VEC*sizeof(V) <= 16)
//   IPW  merge items per worker per tile
//   U    nonzeros whose dense rows are requested before any is consumed (ILP)
// grid = (CTAs, column blocks of G*VEC columns)
// ---------------------------------------------------------------------------------------
template <typename V, int G, int VEC, int IPW, int U, bool DOT>
__global__ void __launch_bounds__(kThreads)
merge_kernel(MergeArgs<V> a, CgScalars cg)
{
    constexpr int W = kThreads / G;   // workers per CTA
    constexpr int TILE = W * IPW;     // merge items per tile
    constexpr int WPW = 32 / G;       // workers per warp
    constexpr int KB = G * VEC;       // columns per column block

    __shared__ int s_row_end[TILE + 1];
    __shared__ int s_col[TILE];
    __shared__ V s_val[TILE];
    __shared__ int s_wkey_last[kWarps], s_wkey_first[kWarps];
    __shared__ V s_wval[kWarps][KB];
    __shared__ V s_carry[KB];
    __shared__ V s_red[kThreads];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = tid / G;    // worker within CTA
    const int li = tid % G;   // lane within worker
    const int wl = lane / G;  // worker within warp

    if constexpr (DOT) {
        // CG graph: once the stop flag is up this launch is a no-op; the first SpMM after
        // the final x update raises HALT so the trailing update kernels no-op as well.
        if (cg.ctrl[CTRL_STOP]) {
            griddep_wait();
            if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) cg.ctrl[CTRL_HALT] = 1;
            return;
        }
        griddep_wait();
        griddep_launch_dependents();
    }

    const size_t k = (size_t)a.k;
    const int c0 = blockIdx.y * KB + li * VEC;  // first column of this lane
    const bool col_ok = c0 < a.k;

    const int t0 = blockIdx.x * a.tiles_per_cta;
    const int t1 = min(t0 + a.tiles_per_cta, a.num_tiles);

    V dot[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) dot[v] = 0;
    V acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0;
    int last_key = a.m;

    if (tid < KB) s_carry[tid] = 0;

    for (int t = t0; t < t1; ++t) {
        const int2 lo = a.tile_xy[t], hi = a.tile_xy[t + 1];
        const int x0 = lo.x, y0 = lo.y;
        const int rows = hi.x - x0, nz = hi.y - y0, items = rows + nz;

        __syncthreads();  // previous tile fully consumed; s_carry published
        for (int i = tid; i <= rows; i += kThreads) {
            int r = x0 + i;
            s_row_end[i] = r < a.m ? __ldg(a.row_end + r) : INT_MAX;
        }
        for (int i = tid; i < nz; i += kThreads) {
            s_col[i] = __ldg(a.ci + y0 + i);
            s_val[i] = __ldg(a.va + y0 + i);
        }
        __syncthreads();

        // ---- this worker's share of the tile: diagonals [d0, d1) ------------------------
        const int d0 = min(w * IPW, items), d1 = min(d0 + IPW, items);
        int r, r_end;
        {
            int l = max(d0 - nz, 0), h = min(d0, rows);
            while (l < h) {
                int mid = (l + h) >> 1;
                if (s_row_end[mid] - y0 <= d0 - mid - 1) l = mid + 1; else h = mid;
            }
            r = l;
            l = max(d1 - nz, 0), h = min(d1, rows);
            while (l < h) {
                int mid = (l + h) >> 1;
                if (s_row_end[mid] - y0 <= d1 - mid - 1) l = mid + 1; else h = mid;
            }
            r_end = l;
        }
        int z = d0 - r;             // local nonzero index
        const int z_end = d1 - r_end;

        // worker 0 continues the row the previous tile of this CTA stopped in
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = (w == 0) ? s_carry[li * VEC + v] : V(0);

        auto emit = [&](int rl) {
            if (col_ok) {
                size_t off = (size_t)(x0 + rl) * k + c0;
                st_vec<V, VEC>(a.Y + off, acc);
                if constexpr (DOT) {
                    V xr[VEC];
                    ldg_vec<V, VEC>(xr, a.X + off);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) dot[v] += acc[v] * xr[v];
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0;
        };

        int cur_end = s_row_end[r] - y0;  // local index one past the current row's last nonzero
        while (z < z_end) {
            V vv[U];
            V xv[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = z + u < z_end;
                const int zi = ok ? z + u : z;
                vv[u] = s_val[zi];
                const int c = s_col[zi];
                if (col_ok) ldg_vec<V, VEC>(xv[u], a.X + (size_t)c * k + c0);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xv[u][v] = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (z + u < z_end) {
                    while (cur_end <= z + u) {  // row ends before this nonzero: emit it
                        emit(r);
                        ++r;
                        cur_end = s_row_end[r] - y0;
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[v] += vv[u] * xv[u][v];
                }
            }
            z += U;
        }
        while (r < r_end) {  // row-end markers that follow the share's last nonzero
            emit(r);
            ++r;
        }

        // ---- carry-outs: segmented inclusive scan over workers keyed by the row in progress
        const int key = x0 + r_end;
#pragma unroll
        for (int d = 1; d < WPW; d <<= 1) {
            const int okey = __shfl_up_sync(0xffffffffu, key, d * G);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const V o = __shfl_up_sync(0xffffffffu, acc[v], d * G);
                if (wl >= d && okey == key) acc[v] += o;
            }
        }
        if (wl == WPW - 1) {
            if (li == 0) s_wkey_last[warp] = key;
#pragma unroll
            for (int v = 0; v < VEC; ++v) s_wval[warp][li * VEC + v] = acc[v];
        }
        if (wl == 0 && li == 0) s_wkey_first[warp] = key;
        __syncthreads();
        for (int i = warp - 1; i >= 0 && s_wkey_last[i] == key; --i) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] += s_wval[i][li * VEC + v];
        }
        bool tail;
        if constexpr (WPW > 1) {
            const int nkey = __shfl_down_sync(0xffffffffu, key, G);
            if (wl < WPW - 1) tail = nkey != key;
            else tail = (warp == kWarps - 1) || (s_wkey_first[warp + 1] != key);
        } else {
            tail = (warp == kWarps - 1) || (s_wkey_first[warp + 1] != key);
        }
        if (tail) {
            if (key < hi.x) {
                // the row was completed inside this tile by a later worker: add the chain
                if (col_ok) {
                    size_t off = (size_t)key * k + c0;
                    V yv[VEC];
                    ld_vec<V, VEC>(yv, a.Y + off);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) yv[v] += acc[v];
                    st_vec<V, VEC>(a.Y + off, yv);
                    if constexpr (DOT) {
                        V xr[VEC];
                        ldg_vec<V, VEC>(xr, a.X + off);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) dot[v] += acc[v] * xr[v];
                    }
                }
            } else {
                // key == hi.x: only the tile's last worker gets here -> tile carry-out
#pragma unroll
                for (int v = 0; v < VEC; ++v) s_carry[li * VEC + v] = acc[v];
                last_key = key;
            }
        }
    }

    // ---- CTA carry-out (the tile carry of its last tile) ---------------------------------
    if (w == W - 1) {
        if (t1 <= t0) last_key = a.m;
        if (li == 0 && blockIdx.y == 0) a.carry_row[blockIdx.x] = last_key;
        if (col_ok) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                a.carry_val[(size_t)blockIdx.x * k + c0 + v] = (t1 > t0) ? acc[v] : V(0);
        }
    }

    if constexpr (DOT) {
        // per-CTA dot partial for each column of this column block
#pragma unroll
        for (int d = G; d < 32; d <<= 1) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) dot[v] += __shfl_xor_sync(0xffffffffu, dot[v], d);
        }
        __syncthreads();
        if (wl == 0) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) s_wval[warp][li * VEC + v] = dot[v];
        }
        __syncthreads();
        if (tid < KB) {
            V s = 0;
            for (int i = 0; i < kWarps; ++i) s += s_wval[i][tid];
            int c = blockIdx.y * KB + tid;
            if (c < a.k) a.dot_part[(size_t)blockIdx.x * k + c] = s;
        }
    }

    // ---- last CTA done: serial-order carry fix-up (merge_based.hpp:137-149) ----------------
    if (!last_cta_election(a.ticket, gridDim.x * gridDim.y)) return;

    const int entries = gridDim.x;
    for (long long idx = tid; idx < (long long)entries * a.k; idx += kThreads) {
        const int e = (int)(idx / a.k), c = (int)(idx % a.k);
        const int row = __ldcg(a.carry_row + e);
        V fixdot = 0;
        if (row < a.m && (e == 0 || __ldcg(a.carry_row + e - 1) != row)) {
            V sum = 0;
            for (int e2 = e; e2 < entries && __ldcg(a.carry_row + e2) == row; ++e2)
                sum += __ldcg(a.carry_val + (size_t)e2 * k + c);
            size_t off = (size_t)row * k + c;
            a.Y[off] = __ldcg(a.Y + off) + sum;
            if constexpr (DOT) fixdot = sum * __ldg(a.X + off);
        }
        if constexpr (DOT) a.fix_part[idx] = fixdot;
    }

    if constexpr (DOT) {
        __syncthreads();
        // pAp[c] then alpha[c] = latched ? 0 : rs_old/pAp  (no_pretreatment.hpp:107-120)
        cta_reduce_columns<V>(a.dot_part, a.fix_part, entries, a.k, a.dot_sum, s_red);   // (V-typed scratch: the scalars are double)
        for (int c = tid; c < a.k; c += kThreads)
            cg_dot_scalars(cg, c, (double)a.dot_sum[c]);
    }
}

} // namespace smle
