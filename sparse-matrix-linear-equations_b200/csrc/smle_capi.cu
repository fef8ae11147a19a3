// smle_capi.cu -- C ABI (include/smle_b200.h) over the sm_100a kernels.
//
// Host-side runtime of the library: device-resident CSR handles with cached merge-path tile
// coordinates, kernel dispatch on (value type, k), the CG driver loop (CUDA-graph batches of
// iterations with device-side convergence control), and error reporting.  No CPU fallback:
// every compute entry point needs a CUDA device.
#include "../../include/smle_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <new>
#include <vector>

#include "smle_cg.cuh"
#include "smle_merge.cuh"
#include "smle_spmv.cuh"
#include "smle_spmm.cuh"
#include "smle_dist.cuh"
#include "smle_plan.h"

using namespace smle;

// =========================================================================================
// error handling / globals
// =========================================================================================
namespace {

thread_local char g_err[512] = "";
cudaStream_t g_own_stream = nullptr;   // created by smle_init
cudaStream_t g_stream = nullptr;       // the stream in use (own or caller's)
cudaStream_t g_copy_stream = nullptr;  // host <-> device copies that overlap a solve (batch solves)
int g_device = -1;
int g_sms = 0;
long long g_launches = 0;
const void *g_spmv_dist = nullptr;   // device DistCtl of the row-partitioned solve being launched

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(SMLE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

int ensure_init()
{
    if (g_device >= 0) return SMLE_OK;
    return smle_init(0);
}

cudaError_t g_launch_err = cudaSuccess;   // first failure of cudaLaunchKernelEx since the last check_launch()

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = g_launch_err;
    g_launch_err = cudaSuccess;
    if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SMLE_OK;
}

// Kernel launch on g_stream; with g_pdl set (the CG iteration loop) the launch carries the
// programmatic-stream-serialization attribute, so the kernel may be scheduled while its
// predecessor drains (smle_common.cuh, griddep_wait).  Captured into CUDA graphs as programmatic edges.
bool g_pdl = false;

bool pdl_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("SMLE_PDL"); on = e ? atoi(e) : 0; }   // measured slower: off
    return on != 0;
}

template <typename... KArgs, typename... Args>
void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = g_stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
    if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;   // reported by check_launch()
}

struct PdlScope {   // CG iteration launches inside this scope use PDL
    bool prev;
    PdlScope() : prev(g_pdl) { g_pdl = pdl_enabled(); }
    ~PdlScope() { g_pdl = prev; }
};

} // namespace

// =========================================================================================
// CSR handle
// =========================================================================================
struct CgWorkspace {
    int k = 0;
    size_t nk = 0;
    double *R = nullptr, *P = nullptr, *AP = nullptr;   // n x k blocks
    double *Bd = nullptr;                               // staging of B for host-pointer calls
    double *Bd2[2] = {nullptr, nullptr};                // batch solves: double-buffered uploads of b
    double *Xs = nullptr;                               // batch solves: x on its way to the host
    double *Xd = nullptr;                               // the iterate (graphs bake this pointer)
    double *scal = nullptr;                             // 6*k doubles + last_rel + tol
    int *conv = nullptr;
    int *ctrl = nullptr;
    double *hist = nullptr;
    int hist_cap = 0;
    double *part = nullptr;                             // vector-kernel partials [grid*k]
    int *ctrl_host = nullptr;                           // pinned mirror of ctrl
    cudaGraphExec_t graph = nullptr;                    // `graph_iters` iterations of K1,K2,K3
    int graph_iters = 0;
    unsigned graph_epoch = 0;                           // scratch_epoch of the handle when the graph was captured
    cudaGraphExec_t graph32 = nullptr;                  // the same for the fp32 solver (blocks reinterpreted as float)
    unsigned graph32_epoch = 0;
    // SPAI-preconditioned CG: z = M r, and the iteration graph (A step, x/r update, M step, p update)
    double *Z = nullptr;
    cudaGraphExec_t pcg_graph = nullptr;
    const void *pcg_m = nullptr;                        // preconditioner handle the graph was captured with
    unsigned pcg_epoch_a = 0, pcg_epoch_m = 0;
};

struct Partition {
    int items_per_tile = 0;
    int num_tiles = 0;
    int2 *xy = nullptr;   // device, num_tiles + 1
    int *maxlen = nullptr; // device, num_tiles: longest in-tile row segment
    int max_len = 0;       // max over maxlen[] (host copy): picks the SpMM kernel
    int general_tiles = 0; // tiles whose longest segment exceeds kRowPathMaxLen (host copy): picks the SpMV configuration
    unsigned char *halo = nullptr;   // device, num_tiles: tile gathers halo columns (row-partitioned handles only)
    int band = -1;                   // window half-width of the band-window SpMM for this tiling (0: none, -1: unknown)
    // structure-aware SpMM tile schedules, keyed by grid size (smle_spmm.cuh); sched == nullptr: none
    struct Sched { int *sched = nullptr, *off = nullptr; };
    std::map<int, Sched> scheds;
};

struct smle_csr_s {
    int m = 0, n = 0, nnz = 0, vbytes = 0;
    int *ro = nullptr, *ci = nullptr;
    void *va = nullptr;
    std::map<int, Partition> parts;   // keyed by items per tile
    // per-launch scratch of the merge kernel (sized for max grid)
    int max_ctas = 0;
    int *carry_row = nullptr;
    void *carry_val = nullptr;  int carry_k = 0;
    void *dot_part = nullptr, *fix_part = nullptr, *dot_sum = nullptr;
    unsigned int *ticket = nullptr;
    void *tile_carry = nullptr;       // carry slots of spmm_rows_kernel (sentinel-filled)
    size_t tile_carry_elems = 0;
    void *cta_slot = nullptr;         // carry slots of the CTA boundaries of spmv_kernel (sentinel-filled)
    // Captured CUDA graphs bake the scratch pointers above: every reallocation bumps the epoch and
    // graphs captured under an older epoch are rebuilt before their next launch.
    unsigned scratch_epoch = 0;
    int halo_base = -1;               // local block of a row partition: first halo column (else -1)
    int far_stride = -1;              // dominant far column offset in rows (0: none, -1: not probed yet)
    int spmv_cfg = 0;                 // configuration of the single-vector kernel picked for this matrix (0: not yet)
    std::vector<int> common_offsets;  // column offsets > 0 that >= 40 % of the sampled rows have, descending
    CgWorkspace ws;
};

namespace {

void free_workspace(CgWorkspace &w)
{
    if (w.graph) cudaGraphExecDestroy(w.graph);
    if (w.pcg_graph) cudaGraphExecDestroy(w.pcg_graph);
    if (w.graph32) cudaGraphExecDestroy(w.graph32);
    cudaFree(w.Z);
    cudaFree(w.R); cudaFree(w.P); cudaFree(w.AP); cudaFree(w.Bd); cudaFree(w.Xd);
    cudaFree(w.Bd2[0]); cudaFree(w.Bd2[1]); cudaFree(w.Xs);
    cudaFree(w.scal); cudaFree(w.conv); cudaFree(w.ctrl); cudaFree(w.hist); cudaFree(w.part);
    if (w.ctrl_host) cudaFreeHost(w.ctrl_host);
    w = CgWorkspace();
}

// ---- merge-path tiling --------------------------------------------------------------------
constexpr int kTileItems = 2048;   // merge items per tile = workers * items-per-worker

int get_partition(smle_csr_t a, int items_per_tile, Partition **out)
{
    auto it = a->parts.find(items_per_tile);
    if (it != a->parts.end()) { *out = &it->second; return SMLE_OK; }
    Partition p;
    p.items_per_tile = items_per_tile;
    long long total = (long long)a->m + a->nnz;
    p.num_tiles = (int)((total + items_per_tile - 1) / items_per_tile);
    if (p.num_tiles < 1) p.num_tiles = 1;
    CU(cudaMalloc(&p.xy, sizeof(int2) * (size_t)(p.num_tiles + 1)));
    int blocks = (p.num_tiles + 1 + 127) / 128;
    merge_partition_kernel<<<blocks, 128, 0, g_stream>>>(a->ro + 1, a->m, a->nnz, items_per_tile,
                                                         p.num_tiles, p.xy);
    ++g_launches;
    int rc = check_launch("merge_partition_kernel");
    if (rc) return rc;
    CU(cudaMalloc(&p.maxlen, sizeof(int) * (size_t)p.num_tiles));
    tile_maxlen_kernel<<<(p.num_tiles * 32 + 255) / 256, 256, 0, g_stream>>>(a->ro, p.xy, p.num_tiles, p.maxlen);
    ++g_launches;
    rc = check_launch("tile_maxlen_kernel");
    if (rc) return rc;
    {
        int *d_max = nullptr;
        int h_max[2] = {0, 0};
        CU(cudaMalloc(&d_max, 2 * sizeof(int)));
        CU(cudaMemsetAsync(d_max, 0, 2 * sizeof(int), g_stream));
        int_max_kernel<<<(p.num_tiles + 255) / 256, 256, 0, g_stream>>>(p.maxlen, p.num_tiles, kRowPathMaxLen, d_max);
        ++g_launches;
        CU(cudaMemcpyAsync(h_max, d_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        cudaFree(d_max);
        p.max_len = h_max[0];
        p.general_tiles = h_max[1];
    }
    if (a->halo_base >= 0) {
        CU(cudaMalloc(&p.halo, (size_t)p.num_tiles));
        tile_halo_kernel<<<(p.num_tiles * 32 + 255) / 256, 256, 0, g_stream>>>(a->ci, p.xy, p.num_tiles, a->halo_base, p.halo);
        ++g_launches;
        rc = check_launch("tile_halo_kernel");
        if (rc) return rc;
    }
    a->parts[items_per_tile] = p;
    *out = &a->parts[items_per_tile];
    return SMLE_OK;
}

int ensure_scratch(smle_csr_t a, int k)
{
    if (!a->carry_row) {
        a->max_ctas = g_sms * 8;
        CU(cudaMalloc(&a->carry_row, sizeof(int) * (size_t)a->max_ctas));
        CU(cudaMalloc(&a->ticket, sizeof(unsigned int) * 4));
        CU(cudaMemsetAsync(a->ticket, 0, sizeof(unsigned int) * 4, g_stream));
        CU(cudaMalloc(&a->cta_slot, 8 * (size_t)a->max_ctas));
        if (a->vbytes == 8) smle::fill_sentinel_kernel<double><<<4, 256, 0, g_stream>>>((double *)a->cta_slot, (size_t)a->max_ctas);
        else smle::fill_sentinel_kernel<float><<<4, 256, 0, g_stream>>>((float *)a->cta_slot, (size_t)a->max_ctas);
        ++g_launches;
        int rc = check_launch("fill_sentinel_kernel");
        if (rc) return rc;
    }
    if (k > a->carry_k) {
        if (g_stream) cudaStreamSynchronize(g_stream);   // nothing in flight may still use the old buffers
        ++a->scratch_epoch;
        cudaFree(a->carry_val); cudaFree(a->dot_part); cudaFree(a->fix_part); cudaFree(a->dot_sum);
        a->carry_val = a->dot_part = a->fix_part = a->dot_sum = nullptr;
        size_t bytes = (size_t)a->max_ctas * (size_t)k * 8;
        CU(cudaMalloc(&a->carry_val, bytes));
        CU(cudaMalloc(&a->dot_part, bytes));
        CU(cudaMalloc(&a->fix_part, bytes));
        CU(cudaMalloc(&a->dot_sum, (size_t)k * 8));
        a->carry_k = k;
    }
    return SMLE_OK;
}

// ---- kernel dispatch ------------------------------------------------------------------------
// (G, VEC) selection: VEC = widest vector that divides k (rows of the dense block must stay
// VEC-aligned), G = lanes needed to cover min(k, 32*VEC) columns, rounded up to a power of two.
template <typename V>
void pick_shape(int k, int *G, int *VEC)
{
    int maxvec = 16 / (int)sizeof(V);
    int vec = 1;
    for (int v = maxvec; v > 1; v >>= 1)
        if (k % v == 0) { vec = v; break; }
    int lanes = (k + vec - 1) / vec;
    int g = 1;
    while (g < lanes && g < 32) g <<= 1;
    *G = g; *VEC = vec;
}

template <typename V, int G, int VEC, bool DOT>
int launch_merge_t(smle_csr_t a, const V *X, V *Y, int k, const CgScalars &cg, bool dry)
{
    constexpr int IPW = kTileItems * G / kThreads;
    constexpr int U = (VEC * sizeof(V) >= 16) ? 4 : 8;
    Partition *p;
    int rc = get_partition(a, kTileItems, &p);
    if (rc) return rc;
    rc = ensure_scratch(a, k);
    if (rc) return rc;

    static int occ = 0;   // per instantiation
    if (!occ) {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, merge_kernel<V, G, VEC, IPW, U, DOT>,
                                                         kThreads, 0));
        if (occ < 1) occ = 1;
        if (occ > 8) occ = 8;
    }
    if (dry) return SMLE_OK;   // only the lazy setup (partition, scratch, occupancy) was wanted
    const int KB = G * VEC;
    const int col_blocks = (k + KB - 1) / KB;
    int max_ctas = g_sms * occ;
    if (max_ctas > a->max_ctas) max_ctas = a->max_ctas;
    int grid = p->num_tiles < max_ctas ? p->num_tiles : max_ctas;
    int tiles_per_cta = (p->num_tiles + grid - 1) / grid;
    grid = (p->num_tiles + tiles_per_cta - 1) / tiles_per_cta;

    MergeArgs<V> args;
    args.row_end = a->ro + 1;
    args.ci = a->ci;
    args.va = (const V *)a->va;
    args.X = X;
    args.Y = Y;
    args.tile_xy = p->xy;
    args.m = a->m; args.nnz = a->nnz; args.k = k;
    args.num_tiles = p->num_tiles; args.tiles_per_cta = tiles_per_cta;
    args.carry_row = a->carry_row;
    args.carry_val = (V *)a->carry_val;
    args.dot_part = (V *)a->dot_part;
    args.fix_part = (V *)a->fix_part;
    args.dot_sum = (V *)a->dot_sum;
    args.ticket = a->ticket;
    launch_kernel(merge_kernel<V, G, VEC, IPW, U, DOT>, dim3(grid, col_blocks), dim3(kThreads), 0, args, cg);
    ++g_launches;
    return check_launch("merge_kernel");
}


// k >= 2, no long rows: the TMA-staged row-per-worker kernel (smle_spmm.cuh)
//   THREADS consumer threads, TILE merge items per tile, STAGES tiles in flight, MINB CTAs per SM.
// SMLE_SPMM_CFG=<threads>x<tile>x<stages>x<minb> selects one of the instantiated configurations,
// SMLE_SPMM_CHUNK the number of consecutive tiles per deal (0 = contiguous runs per CTA).
constexpr int kSpmmRowsMaxLen = 64;    // longest in-tile row segment the row-per-worker kernel accepts

template <typename V>
int ensure_tile_carry(smle_csr_t a, size_t elems)
{
    if (elems <= a->tile_carry_elems) return SMLE_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    ++a->scratch_epoch;
    cudaFree(a->tile_carry);
    a->tile_carry = nullptr; a->tile_carry_elems = 0;
    CU(cudaMalloc(&a->tile_carry, sizeof(V) * elems));
    fill_sentinel_kernel<V><<<g_sms * 4, 256, 0, g_stream>>>((V *)a->tile_carry, elems);
    ++g_launches;
    a->tile_carry_elems = elems;
    return check_launch("fill_sentinel_kernel");
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---- structure-aware tile schedule of the row-per-worker SpMM kernel --------------------------------
// far_stride: the largest column offset D > 0 that at least 40 % of a sample of rows have a nonzero
// at (the w^2 plane offset of a 3-D stencil); 0 when there is none (R-MAT, wheel, 2-D grids whose far
// offset is within a tile or two).
int far_stride(smle_csr_t a, int *out)
{
    if (a->far_stride >= 0) { *out = a->far_stride; return SMLE_OK; }
    a->far_stride = 0;
    a->common_offsets.clear();
    *out = 0;
    if (a->m < 4096 || a->m != a->n) return SMLE_OK;
    const int ns = 2048;
    int *d_out = nullptr;
    CU(cudaMalloc(&d_out, sizeof(int) * ns * kSampleNnz));
    sample_offsets_kernel<<<(ns + 127) / 128, 128, 0, g_stream>>>(a->ro, a->ci, a->m, ns, d_out);
    ++g_launches;
    std::vector<int> h((size_t)ns * kSampleNnz);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_out, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "structure probe failed: %s", cudaGetErrorString(e));
    std::map<int, int> freq;   // offset -> rows of the sample that have it
    for (int i = 0; i < ns; ++i) {
        int seen[kSampleNnz], nseen = 0;
        for (int j = 0; j < kSampleNnz; ++j) {
            const int o = h[(size_t)i * kSampleNnz + j];
            if (o == INT_MIN || o <= 0) continue;
            bool dup = false;
            for (int q = 0; q < nseen; ++q) dup |= seen[q] == o;
            if (!dup) { seen[nseen++] = o; ++freq[o]; }
        }
    }
    for (auto it = freq.rbegin(); it != freq.rend(); ++it)
        if (it->second * 10 >= ns * 4) {
            if (!a->far_stride) a->far_stride = it->first;
            a->common_offsets.push_back(it->first);   // descending
        }
    *out = a->far_stride;
    return SMLE_OK;
}

// Chains of tiles D rows apart, cut into segments, dealt round-robin in order of their first tile.
int get_tile_sched(smle_csr_t a, Partition *p, int grid, const int **sched, const int **off)
{
    *sched = *off = nullptr;
    static int enabled = -1;
    if (enabled < 0) enabled = env_int("SMLE_SPMM_SCHED", 0);   // measured slower than the round-robin deal (profiles/r02_spmm_tile_schedule_ab.txt): off
    if (!enabled) return SMLE_OK;
    auto found = p->scheds.find(grid);
    if (found != p->scheds.end()) { *sched = found->second.sched; *off = found->second.off; return SMLE_OK; }
    Partition::Sched sc;
    int D = 0;
    int rc = far_stride(a, &D);
    if (rc) return rc;
    const int T = p->num_tiles;
    const double rows_per_tile = (double)a->m / (double)(T > 0 ? T : 1);
    if (D > 0 && (double)D >= 8.0 * rows_per_tile && T >= 4 * grid) {
        std::vector<int2> xy((size_t)T + 1);
        CU(cudaMemcpyAsync(xy.data(), p->xy, sizeof(int2) * xy.size(), cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        std::vector<int> first((size_t)T);
        for (int t = 0; t < T; ++t) first[(size_t)t] = xy[(size_t)t].x;
        // successor of tile t: the tile that owns row first[t] + D
        std::vector<char> used((size_t)T, 0);
        std::vector<std::vector<int>> chains;
        for (int t = 0; t < T; ++t) {
            if (used[(size_t)t]) continue;
            chains.emplace_back();
            int c = t;
            while (c < T && !used[(size_t)c]) {
                used[(size_t)c] = 1;
                chains.back().push_back(c);
                const long long want = (long long)first[(size_t)c] + D;
                if (want >= a->m) break;
                c = (int)(std::upper_bound(first.begin(), first.end(), (int)want) - first.begin()) - 1;
            }
        }
        // segment length: the deal must balance (units per CTA x tiles per unit close to T / grid)
        size_t longest = 0;
        for (auto &ch : chains) longest = std::max(longest, ch.size());
        // cost of a segment length = tiles on the busiest CTA + 0.3 tile-times per segment start (its first
        // tile finds none of its -D rows in L1)
        double best_cost = -1.0;
        int best_seg = 0;
        for (int seg = (int)longest; seg >= 4; --seg) {
            long long units = 0;
            for (auto &ch : chains) units += ((long long)ch.size() + seg - 1) / seg;
            const long long per_cta = (units + grid - 1) / grid;
            const double cost = (double)(per_cta * seg) + 0.3 * (double)per_cta;
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_seg = seg; }
        }
        if (best_seg > 0) {
            struct Unit { int first, chain, begin, len; };
            std::vector<Unit> units;
            for (int ci2 = 0; ci2 < (int)chains.size(); ++ci2)
                for (int b0 = 0; b0 < (int)chains[(size_t)ci2].size(); b0 += best_seg)
                    units.push_back({chains[(size_t)ci2][(size_t)b0], ci2, b0, std::min(best_seg, (int)chains[(size_t)ci2].size() - b0)});
            std::sort(units.begin(), units.end(), [](const Unit &x, const Unit &y) { return x.first < y.first; });
            std::vector<int> h_off((size_t)grid + 1, 0), h_sched;
            h_sched.reserve((size_t)T + grid);
            for (int b = 0; b < grid; ++b) {
                for (size_t u = (size_t)b; u < units.size(); u += (size_t)grid)
                    for (int i = 0; i < units[u].len; ++i) h_sched.push_back(chains[(size_t)units[u].chain][(size_t)(units[u].begin + i)]);
                h_sched.push_back(T);   // sentinel: end of this CTA's list
                h_off[(size_t)b + 1] = (int)h_sched.size();
            }
            if ((int)h_sched.size() == T + grid) {
                CU(cudaMalloc(&sc.sched, sizeof(int) * h_sched.size()));
                CU(cudaMalloc(&sc.off, sizeof(int) * ((size_t)grid + 1)));
                CU(cudaMemcpyAsync(sc.sched, h_sched.data(), sizeof(int) * h_sched.size(), cudaMemcpyHostToDevice, g_stream));
                CU(cudaMemcpyAsync(sc.off, h_off.data(), sizeof(int) * ((size_t)grid + 1), cudaMemcpyHostToDevice, g_stream));
                CU(cudaStreamSynchronize(g_stream));
            }
        }
    }
    p->scheds[grid] = sc;
    *sched = sc.sched; *off = sc.off;
    return SMLE_OK;
}

template <typename V, int G, int VEC, int NV, int UB, int THREADS, int TILE, int STAGES, int MINB, bool DOT>
int launch_spmm_rows_t(smle_csr_t a, const V *X, V *Y, int k, const CgScalars &cg, bool dry)
{
    using SM = SpmmSmem<V, TILE>;
    constexpr size_t smem = SM::STAGE_BYTES * STAGES;
    auto kern = spmm_rows_kernel<V, G, VEC, NV, UB, THREADS, TILE, STAGES, MINB, DOT>;
    Partition *p;
    int rc = get_partition(a, TILE, &p);
    if (rc) return rc;
    rc = ensure_scratch(a, k);
    if (!rc) rc = ensure_tile_carry<V>(a, (size_t)p->num_tiles * (size_t)k);
    if (rc) return rc;
    static int occ = 0;   // per instantiation
    if (!occ) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // leave the rest of the 256 KB to L1: the dense-row gathers live there
        // (the carve-out comes in steps; every CTA also takes 1 KB of reserved shared memory)
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, kern));
        const size_t need = (smem + fa.sharedSizeBytes + 1024) * MINB;
        int carve = 100;
        for (int kb : {8, 16, 32, 64, 100, 132, 164, 196, 228})
            if (need <= (size_t)kb * 1024) { carve = (kb * 100 + 227) / 228; break; }
        // SMLE_SPMM_CARVE: -1 (default) the smallest carve-out that fits, 0 leave the choice to the driver, else percent
        const int carve_env = env_int("SMLE_SPMM_CARVE", -1);
        if (carve_env > 0) carve = carve_env;
        if (carve_env != 0) CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS + 32, smem));
        if (occ < 1) return fail(SMLE_ERR_CUDA, "spmm_rows_kernel does not fit on an SM (%zu B smem)", smem);
        if (occ > MINB) occ = MINB;
    }
    int grid = g_sms * occ;
    if (grid > a->max_ctas) grid = a->max_ctas;
    if (grid > p->num_tiles) grid = p->num_tiles;
    const int *sched = nullptr, *sched_off = nullptr;
    rc = get_tile_sched(a, p, grid, &sched, &sched_off);   // built once per (partition, grid): part of the lazy setup
    if (rc) return rc;
    if (dry) return SMLE_OK;
    // consecutive tiles per deal: 1 everywhere except the 16-lane shape (k = 32 fp64 / 64 fp32), where 2
    // measured 5 % faster (profiles/r01_spmm_sweeps.txt); SMLE_SPMM_CHUNK overrides, 0 = contiguous runs
    static int chunk_env = -2;
    if (chunk_env == -2) chunk_env = env_int("SMLE_SPMM_CHUNK", -1);
    int chunk = chunk_env > 0 ? chunk_env : (chunk_env == 0 ? (p->num_tiles + grid - 1) / grid : (G == 16 ? 2 : 1));
    SpmmArgs<V> args;
    args.ro = a->ro; args.ci = a->ci; args.va = (const V *)a->va;
    args.X = X; args.Y = Y; args.tile_xy = p->xy;
    args.m = a->m; args.nnz = a->nnz; args.k = k;
    args.num_tiles = p->num_tiles; args.chunk = chunk;
    args.sched = sched; args.sched_off = sched_off;
    args.tile_carry = (V *)a->tile_carry;
    args.dot_part = (V *)a->dot_part;
    args.dot_sum = (V *)a->dot_sum;
    args.ticket = a->ticket;
    args.band = 0;
    static int ypol = -1;
    if (ypol < 0) ypol = env_int("SMLE_SPMM_YPOL", 1);
    args.y_policy = ypol;
    static int dot_late = -1;
    if (dot_late < 0) {
        dot_late = env_int("SMLE_SPMM_DOT_LATE", 1);
        if (env_int("SMLE_DEBUG_DISPATCH", 0))
            fprintf(stderr, "[smle] spmm k=%d: row-per-worker kernel, %d threads, %d tiles of %d items, chunk %d, grid %d, schedule %s\n",
                    k, THREADS, p->num_tiles, TILE, chunk, grid, sched ? "chains" : "round-robin");
    }
    args.dot_late = dot_late;
    launch_kernel(kern, dim3(grid), dim3(THREADS + 32), smem, args, cg);
    ++g_launches;
    return check_launch("spmm_rows_kernel");
}

// ---- band-window variant (k = 32 fp64) ----------------------------------------------------------------
// The ring holds RING dense rows of 256 B next to two stages of 960-item tiles (212 KB in all).  The
// window half-width is the largest column offset most rows share that still fits: a tile and its
// successor must both find their windows in the ring, 2*band + rows(t) + rows(t+1) + 2 <= RING.
constexpr int kSpmmThreads = 960, kSpmmStages = 2, kSpmmUB = 4;   // default configuration of the row-per-worker kernel
constexpr int kBandRing = 704, kBandTile = 960;

int band_halfwidth(smle_csr_t a, Partition *p, int *band)
{
    *band = 0;
    if (p->band >= 0) { *band = p->band; return SMLE_OK; }
    p->band = 0;
    int D = 0;
    int rc = far_stride(a, &D);
    if (rc || a->common_offsets.empty()) return rc;
    std::vector<int2> xy((size_t)p->num_tiles + 1);
    CU(cudaMemcpyAsync(xy.data(), p->xy, sizeof(int2) * xy.size(), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    // rows of two consecutive tiles, 90th percentile: the kernel clamps the window of the few tiles that hold
    // more (shorter) rows, so the common case decides the band
    std::vector<int> pair;
    for (int t = 0; t + 2 <= p->num_tiles; ++t) pair.push_back(xy[(size_t)t + 2].x - xy[(size_t)t].x);
    if (pair.empty()) pair.push_back(xy[1].x - xy[0].x);
    std::sort(pair.begin(), pair.end());
    const int pair_typ = pair[(pair.size() - 1) * 9 / 10];
    const int cap = (kBandRing - pair_typ - 4) / 2;
    for (int off : a->common_offsets)   // descending
        if (off <= cap) { p->band = off; break; }
    if (p->band < 8) p->band = 0;       // a band that narrow is what L1 already catches
    *band = p->band;
    return SMLE_OK;
}

template <bool DOT>
int launch_spmm_band(smle_csr_t a, const double *X, double *Y, int k, const CgScalars &cg, bool dry, bool *done)
{
    using V = double;
    constexpr int G = 16, VEC = 2, THREADS = 960, STAGES = 2;
    using SM = SpmmSmem<V, kBandTile>;
    constexpr size_t smem = SM::STAGE_BYTES * STAGES + (size_t)kBandRing * G * VEC * sizeof(V);
    auto kern = spmm_rows_kernel<V, G, VEC, 1, kSpmmUB, THREADS, kBandTile, STAGES, 1, DOT, kBandRing>;
    *done = false;
    Partition *p;
    int rc = get_partition(a, kBandTile, &p);
    if (rc) return rc;
    if (p->max_len > kSpmmRowsMaxLen) return SMLE_OK;
    int band = 0;
    rc = band_halfwidth(a, p, &band);
    if (rc || band == 0) return rc;
    rc = ensure_scratch(a, k);
    if (!rc) rc = ensure_tile_carry<V>(a, (size_t)p->num_tiles * (size_t)k);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr = true;
    }
    *done = true;
    if (dry) return SMLE_OK;
    int grid = g_sms;
    if (grid > p->num_tiles) grid = p->num_tiles;
    static int chunk = -1;
    if (chunk < 0) {
        chunk = env_int("SMLE_SPMM_BAND_CHUNK", 16);
        if (env_int("SMLE_DEBUG_DISPATCH", 0))
            fprintf(stderr, "[smle] spmm k=%d: band-window kernel, band %d rows, ring %d rows, %d tiles of %d items, chunk %d, %zu B smem\n",
                    k, band, kBandRing, p->num_tiles, kBandTile, chunk, smem);
    }
    SpmmArgs<V> args;
    args.ro = a->ro; args.ci = a->ci; args.va = (const V *)a->va;
    args.X = X; args.Y = Y; args.tile_xy = p->xy;
    args.m = a->m; args.nnz = a->nnz; args.k = k;
    args.num_tiles = p->num_tiles; args.chunk = chunk > 0 ? chunk : 1;
    args.sched = nullptr; args.sched_off = nullptr;
    args.tile_carry = (V *)a->tile_carry;
    args.dot_part = (V *)a->dot_part;
    args.dot_sum = (V *)a->dot_sum;
    args.ticket = a->ticket;
    args.band = band;
    args.y_policy = 1;
    args.dot_late = 1;
    launch_kernel(kern, dim3(grid), dim3(THREADS + 32), smem, args, cg);
    ++g_launches;
    return check_launch("spmm_rows_kernel (band window)");
}

// <threads>x<tile>x<stages>x<minb>x<ub>x<nv> as one integer
constexpr long long spmm_cfg_id(int th, int tl, int st, int mb, int ub, int nv)
{
    return ((((long long)th * 10000 + tl) * 10 + st) * 10 + mb) * 1000 + ub * 10 + nv;
}

long long spmm_cfg()
{
    static long long cfg = -1;
    if (cfg < 0) {
        cfg = 0;   // 0: the default of the shape
        const char *e = getenv("SMLE_SPMM_CFG");
        int th = 0, tl = 0, st = 0, mb = 0, ub = 0, nv = 0;
        if (e && sscanf(e, "%dx%dx%dx%dx%dx%d", &th, &tl, &st, &mb, &ub, &nv) == 6) cfg = spmm_cfg_id(th, tl, st, mb, ub, nv);
    }
    return cfg;
}

constexpr int kSpmmTile = 1920, kSpmmTileNarrow = 1440;

// G, VEC as picked by pick_shape.  Default configuration (sweeps in profiles/r01_spmm_sweeps.txt):
// ONE CTA of 30 consumer warps + producer per SM (64 registers per thread; ptxas keeps ~4 dense-row
// loads in flight per warp, the scoreboard count), tiles of 1920 merge items (~4 rows per worker
// for a 7-point stencil, 2 stages = 62 KB so the carve-out stays at 64 KB and L1 at 192 KB), dealt
// round-robin in chunks of 2.  Blocks wider than 32 lanes
// (k > 32*VEC) give every lane two vectors so that the fused p.Ap still sees all columns.

template <typename V, int G, int VEC, bool DOT>
int launch_spmm_rows(smle_csr_t a, const V *X, V *Y, int k, const CgScalars &cg, bool dry)
{
    const long long cfg = spmm_cfg();
    if constexpr (G == 16 && VEC == 2 && sizeof(V) == 8) {
        static int band_on = -1;
        if (band_on < 0) band_on = env_int("SMLE_SPMM_BAND", 0);
        if (band_on && cfg == 0 && k == 32 && a->m == a->n) {
            bool done = false;
            int rc = launch_spmm_band<DOT>(a, X, Y, k, cg, dry, &done);
            if (rc || done) return rc;
        }
    }
    if constexpr (G == 16 && VEC == 2 && sizeof(V) == 8) {   // tuning variants (k = 32 fp64 only)
        if (cfg != 0) {
#define SMLE_CFG(th, tl, st, mb, ub, nv) \
    if (cfg == spmm_cfg_id(th, tl, st, mb, ub, nv)) return launch_spmm_rows_t<V, G / nv, VEC, nv, ub, th, tl, st, mb, DOT>(a, X, Y, k, cg, dry);
            SMLE_CFG(960, 1920, 2, 1, 4, 1) SMLE_CFG(960, 2048, 2, 1, 4, 1) SMLE_CFG(960, 1920, 2, 1, 8, 1)
            SMLE_CFG(480, 2048, 2, 2, 4, 1) SMLE_CFG(224, 512, 2, 4, 4, 1) SMLE_CFG(960, 1920, 2, 1, 4, 2)
            SMLE_CFG(960, 1920, 2, 1, 2, 2) SMLE_CFG(960, 1920, 2, 1, 3, 2)
#undef SMLE_CFG
            return fail(SMLE_ERR_ARG, "unsupported SMLE_SPMM_CFG");
        }
    } else if (cfg != 0) {
        return fail(SMLE_ERR_ARG, "SMLE_SPMM_CFG variants exist for fp64 k = 32 only");
    }
    if (G == 32 && k > G * VEC)
        return launch_spmm_rows_t<V, G, VEC, 2, kSpmmUB, kSpmmThreads, kSpmmTile, kSpmmStages, 1, DOT>(a, X, Y, k, cg, dry);
    // narrow blocks: a tile of 1920 items has ~240 rows, so 960 threads of G <= 2 lanes per row would
    // mostly idle; smaller CTAs, more of them per SM, keep the same number of rows in flight
    if constexpr (G <= 2)
        return launch_spmm_rows_t<V, G, VEC, 1, kSpmmUB, 224, kSpmmTileNarrow, kSpmmStages, 4, DOT>(a, X, Y, k, cg, dry);
    if constexpr (G <= 8)
        return launch_spmm_rows_t<V, G, VEC, 1, kSpmmUB, 480, kSpmmTile, kSpmmStages, 2, DOT>(a, X, Y, k, cg, dry);
    return launch_spmm_rows_t<V, G, VEC, 1, kSpmmUB, kSpmmThreads, kSpmmTile, kSpmmStages, 1, DOT>(a, X, Y, k, cg, dry);
}

int spmm_tile_items(int G)
{
    const long long cfg = spmm_cfg();
    if (cfg && G == 16) return (int)((cfg / 100000) % 10000);
    return G <= 2 ? kSpmmTileNarrow : kSpmmTile;
}

// does the row-per-worker kernel take this (matrix, k)?   G, VEC: the shape pick_shape chose
int spmm_use_rows(smle_csr_t a, int G, int VEC, int k, bool dot, bool *use)
{
    *use = false;
    if (getenv("SMLE_SPMM_V1") != nullptr) return SMLE_OK;
    const int kb = G * VEC * (G == 32 ? 2 : 1);
    if (dot && k > kb) return SMLE_OK;   // the fused p.Ap needs all k columns in one block
    static int maxlen = -1;
    if (maxlen < 0) maxlen = env_int("SMLE_SPMM_ROWS_MAXLEN", kSpmmRowsMaxLen);
    Partition *p;
    int rc = get_partition(a, spmm_tile_items(G), &p);
    if (rc) return rc;
    *use = p->max_len <= maxlen;
    return SMLE_OK;
}

// k == 1: the TMA-staged single-vector kernel (smle_spmv.cuh)
//   THREADS threads per CTA
//   IPT     merge items per thread per tile (tile = THREADS*IPT items)
//   STAGES  tiles in flight per CTA
// The default was picked from the sweep in profiles/ (SMLE_SPMV_CFG=<threads>x<ipt>x<stages>
// overrides it for experiments).
template <typename V, int THREADS, int IPT, int STAGES, bool DOT, int MAXB = 0>
int launch_spmv_t(smle_csr_t a, const V *x, V *y, const CgScalars &cg, bool dry)
{
    using SM = SpmvSmem<V, THREADS, IPT>;
    constexpr size_t smem = SM::STAGE_BYTES * STAGES;
    auto kern = spmv_kernel<V, THREADS, IPT, STAGES, DOT, MAXB>;
    Partition *p;
    int rc = get_partition(a, SM::TILE, &p);
    if (rc) return rc;
    rc = ensure_scratch(a, 1);
    if (rc) return rc;
    static int occ = 0;   // per instantiation
    if (!occ) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS + 32, smem));
        if (occ < 1) return fail(SMLE_ERR_CUDA, "spmv_kernel does not fit on an SM (%zu B smem)", smem);
        if (occ > 8) occ = 8;
        if (MAXB > 0 && occ > MAXB) occ = MAXB;
        // (no carve-out preference: asking for the smallest carve-out that holds the resident CTAs -- more L1 --
        // measured slower on R-MAT, 551 vs 533 us at scale 22; profiles/r02_spmv_skewed_cfg.txt)
    }
    if (dry) return SMLE_OK;
    // SMLE_SPMV_WAVES=<w>: w times more CTAs than fit on the GPU at once, each with a shorter run of tiles; the
    // hardware block scheduler then evens out runs of unequal cost (skewed matrices)
    static int waves = 0;
    if (!waves) { waves = env_int("SMLE_SPMV_WAVES", 1); if (waves < 1) waves = 1; }
    int max_ctas = g_sms * occ * waves;
    if (max_ctas > a->max_ctas) max_ctas = a->max_ctas;
    int grid = p->num_tiles < max_ctas ? p->num_tiles : max_ctas;
    int tiles_per_cta = (p->num_tiles + grid - 1) / grid;
    grid = (p->num_tiles + tiles_per_cta - 1) / tiles_per_cta;
    SpmvArgs<V> args;
    args.ro = a->ro; args.ci = a->ci; args.va = (const V *)a->va;
    args.x = x; args.y = y; args.tile_xy = p->xy; args.tile_maxlen = p->maxlen;
    args.m = a->m; args.nnz = a->nnz;
    args.num_tiles = p->num_tiles; args.tiles_per_cta = tiles_per_cta;
    args.cta_slot = (V *)a->cta_slot;
    args.dot_part = (V *)a->dot_part;
    args.ticket = a->ticket;
    args.dist = (const DistCtl *)g_spmv_dist;
    args.tile_halo = g_spmv_dist ? p->halo : nullptr;
    { static int dbg = -1; if (dbg < 0) { const char *e = getenv("SMLE_SPMV_DEBUG"); dbg = e ? atoi(e) : 0; } args.debug_flags = dbg; }
    {
        // small systems (matrix + the vectors of the caller: x, y, and r, p, x of a CG iteration) stay in the 126 MB L2
        // from one product to the next when the matrix stream does not ask to be evicted first
        static int keep_mb = -1;
        if (keep_mb < 0) keep_mb = env_int("SMLE_SPMV_KEEP_MB", 96);
        const size_t foot = (size_t)a->nnz * (4 + sizeof(V)) + ((size_t)a->m + 1) * 4 +
                            (size_t)(DOT ? 5 : 2) * (size_t)(a->n > a->m ? a->n : a->m) * sizeof(V);
        args.keep_l2 = foot <= (size_t)keep_mb * 1024 * 1024;
    }
    { static int med = -1; if (med < 0) { med = env_int("SMLE_SPMV_MEDLO", kRowPathMaxLen); if (med < 8) med = 8; if (med > kRowPathMaxLen) med = kRowPathMaxLen; } args.med_lo = med; }
    if (args.med_lo < SM::TILE / kLongCap + 1) args.med_lo = SM::TILE / kLongCap + 1;   // the tile's queue holds kLongCap segments
    launch_kernel(kern, dim3(grid), dim3(THREADS + 32), smem, args, cg);   // + the producer warp
    ++g_launches;
    return check_launch("spmv_kernel");
}

constexpr int kSpmvThreads = 480, kSpmvIPT = 6, kSpmvStages = 2;   // default configuration (profiles/r01_spmv_sweeps.txt)

int spmv_cfg()   // threads*10000 + ipt*100 + stages (+ 10000000 * CTAs-per-SM cap, the optional 4th field)
{
    static int cfg = -1;
    if (cfg < 0) {
        cfg = kSpmvThreads * 10000 + kSpmvIPT * 100 + kSpmvStages;
        const char *e = getenv("SMLE_SPMV_CFG");
        int th = 0, i = 0, st = 0, mb = 0;
        if (e) {
            const int nf = sscanf(e, "%dx%dx%dx%d", &th, &i, &st, &mb);
            if (nf >= 3) cfg = th * 10000 + i * 100 + st + (nf == 4 ? mb : 0) * 10000000;
        }
    }
    return cfg;
}

bool spmv_cfg_forced() { return getenv("SMLE_SPMV_CFG") != nullptr; }

constexpr int kSpmvCfgSmall = 480 * 10000 + 4 * 100 + 3;                 // 480x4x3
constexpr int kSpmvCfgSkewed = 20000000 + 480 * 10000 + 4 * 100 + 2;     // 480x4x2, two CTAs per SM

// The configuration of the single-vector kernel for this matrix (cached on the handle):
//   * SMLE_SPMV_CFG when set;
//   * small problems (fewer than ~16 default tiles per CTA, e.g. grid2d 1000^2: 7): tiles of 1920 items in 3
//     stages start the first row sooner and drain faster (profiles/r02_spmv_grid2d_1000_cfg_sweep.jsonl);
//   * skewed fp64 matrices (half of the default tiles or more hold a row segment longer than kRowPathMaxLen:
//     R-MAT; the wheel, a third, is as fast on the default): two CTAs per SM with two stages of 1920 items each -- the scattered
//     x gathers of such matrices ask one L2 slice for the same few hot lines, and what L1 absorbs never gets
//     there: 76 KB of stages per CTA leave L1 ~92 KB where the default's 107 KB leave ~28 KB (R-MAT scale 22 /
//     23 / 24: 626 / 1400 / 3906 -> 508 / 907 / 2123 us; one CTA of 640 threads with the same L1: 536 / 1029 /
//     2521; profiles/r02_spmv_skewed_cfg.txt);
//   * else 480x6x2.
int spmv_pick(smle_csr_t a, int *cfg)
{
    if (a->spmv_cfg) { *cfg = a->spmv_cfg; return SMLE_OK; }
    int c = spmv_cfg();
    if (!spmv_cfg_forced()) {
        if ((long long)a->m + a->nnz < 16LL * kSpmvThreads * kSpmvIPT * 2 * g_sms) {
            c = kSpmvCfgSmall;
        } else {
            Partition *p;
            int rc = get_partition(a, kSpmvThreads * kSpmvIPT, &p);
            if (rc) return rc;
            static int skew_on = -1;
            if (skew_on < 0) skew_on = env_int("SMLE_SPMV_SKEWED_CFG", 1);
            // fp64 only: the fp32 stages are half as large, two CTAs already leave L1 ~64 KB, and one CTA per SM
            // measured slower there (R-MAT scale 24 fp32: 2.06 ms against 1.73)
            if (skew_on && a->vbytes == 8 && (long long)p->general_tiles * 2 >= p->num_tiles) c = kSpmvCfgSkewed;
        }
    }
    a->spmv_cfg = *cfg = c;
    return SMLE_OK;
}

int spmv_tile_items(smle_csr_t a, int *items)
{
    int c = 0;
    int rc = spmv_pick(a, &c);
    if (rc) return rc;
    *items = ((c / 10000) % 1000) * ((c / 100) % 100);
    return SMLE_OK;
}

template <typename V, bool DOT>
int launch_spmv(smle_csr_t a, const V *x, V *y, const CgScalars &cg, bool dry)
{
    int c = 0;
    int rc = spmv_pick(a, &c);
    if (rc) return rc;
    switch (c) {
#define SMLE_CFG(th, i, st) case th * 10000 + i * 100 + st: return launch_spmv_t<V, th, i, st, DOT>(a, x, y, cg, dry);
        SMLE_CFG(480, 6, 2) SMLE_CFG(480, 5, 2) SMLE_CFG(480, 7, 2) SMLE_CFG(256, 12, 2) SMLE_CFG(224, 8, 2) SMLE_CFG(960, 4, 2)
        SMLE_CFG(480, 4, 3) SMLE_CFG(480, 3, 4) SMLE_CFG(320, 6, 3) SMLE_CFG(640, 6, 2)
#undef SMLE_CFG
#define SMLE_CFG1(th, i, st) case 10000000 + th * 10000 + i * 100 + st: return launch_spmv_t<V, th, i, st, DOT, 1>(a, x, y, cg, dry);
        SMLE_CFG1(640, 6, 2) SMLE_CFG1(480, 6, 2) SMLE_CFG1(480, 8, 2)   // (640x9x2, 960x6x2, 320x12x2, 960x3x2, 960x2x2, 640x4x2 were measured and dropped: profiles/r02_spmv_skewed_cfg.txt)
#undef SMLE_CFG1
        case 20000000 + 480 * 10000 + 4 * 100 + 2: return launch_spmv_t<V, 480, 4, 2, DOT, 2>(a, x, y, cg, dry);   // 480x4x2x2   (480x3x2x2, stages of 1440 items and L1 ~124 KB: 980 / 2154 us at scale 23 / 24 against 907 / 2123 -- dropped)
    }
    return fail(SMLE_ERR_ARG, "unsupported SMLE_SPMV_CFG");
}

template <typename V, bool DOT>
int launch_merge(smle_csr_t a, const V *X, V *Y, int k, const CgScalars &cg, bool dry = false)
{
    if (k == 1 && getenv("SMLE_SPMV_V1") == nullptr) return launch_spmv<V, DOT>(a, X, Y, cg, dry);
    int G, VEC;
    pick_shape<V>(k, &G, &VEC);
    bool rows_kernel = false;
    int rc0 = spmm_use_rows(a, G, VEC, k, DOT, &rows_kernel);
    if (rc0) return rc0;
#define SMLE_CASE(g, v)                                                                 \
    if (G == g && VEC == v) {                                                           \
        if (rows_kernel) return launch_spmm_rows<V, g, v, DOT>(a, X, Y, k, cg, dry);    \
        return launch_merge_t<V, g, v, DOT>(a, X, Y, k, cg, dry);                       \
    }
    if constexpr (sizeof(V) == 8) {
        SMLE_CASE(1, 1) SMLE_CASE(2, 1) SMLE_CASE(4, 1) SMLE_CASE(8, 1) SMLE_CASE(16, 1) SMLE_CASE(32, 1)
        SMLE_CASE(1, 2) SMLE_CASE(2, 2) SMLE_CASE(4, 2) SMLE_CASE(8, 2) SMLE_CASE(16, 2) SMLE_CASE(32, 2)
    } else {
        SMLE_CASE(1, 1) SMLE_CASE(2, 1) SMLE_CASE(4, 1) SMLE_CASE(8, 1) SMLE_CASE(16, 1) SMLE_CASE(32, 1)
        SMLE_CASE(1, 2) SMLE_CASE(2, 2) SMLE_CASE(4, 2) SMLE_CASE(8, 2) SMLE_CASE(16, 2) SMLE_CASE(32, 2)
        SMLE_CASE(1, 4) SMLE_CASE(2, 4) SMLE_CASE(4, 4) SMLE_CASE(8, 4) SMLE_CASE(16, 4) SMLE_CASE(32, 4)
    }
#undef SMLE_CASE
    return fail(SMLE_ERR_ARG, "no kernel for k=%d", k);
}

template <typename V>
int csr_create(smle_csr_t *out, int m, int n, int nnz, const int *ro, const int *ci, const V *va)
{
    if (!out || m < 0 || n < 0 || nnz < 0 || !ro || (nnz > 0 && (!ci || !va)))
        return fail(SMLE_ERR_ARG, "smle_csr_create: bad argument");
    if ((long long)m + (long long)nnz > (long long)INT_MAX - kTileItems)
        return fail(SMLE_ERR_RANGE, "m + nnz = %lld exceeds the 32-bit merge path of the reference",
                    (long long)m + nnz);
    int rc = ensure_init();
    if (rc) return rc;
    smle_csr_s *a = new (std::nothrow) smle_csr_s();
    if (!a) return fail(SMLE_ERR_ALLOC, "out of host memory");
    a->m = m; a->n = n; a->nnz = nnz; a->vbytes = (int)sizeof(V);
    // 16 B of slack behind every array: tiles are staged with vector / bulk copies
    cudaError_t e;
    if ((e = cudaMalloc(&a->ro, sizeof(int) * ((size_t)m + 1) + 16)) != cudaSuccess ||
        (e = cudaMalloc(&a->ci, sizeof(int) * (size_t)nnz + 16)) != cudaSuccess ||
        (e = cudaMalloc(&a->va, sizeof(V) * (size_t)nnz + 16)) != cudaSuccess) {
        smle_csr_destroy(a);
        return fail(SMLE_ERR_ALLOC, "device allocation failed: %s", cudaGetErrorString(e));
    }
    cudaMemsetAsync((char *)a->ro + sizeof(int) * ((size_t)m + 1), 0, 16, g_stream);
    cudaMemsetAsync((char *)a->ci + sizeof(int) * (size_t)nnz, 0, 16, g_stream);
    cudaMemsetAsync((char *)a->va + sizeof(V) * (size_t)nnz, 0, 16, g_stream);
    if ((e = cudaMemcpyAsync(a->ro, ro, sizeof(int) * ((size_t)m + 1), cudaMemcpyHostToDevice, g_stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(a->ci, ci, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, g_stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(a->va, va, sizeof(V) * (size_t)nnz, cudaMemcpyHostToDevice, g_stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(g_stream)) != cudaSuccess) {
        smle_csr_destroy(a);
        return fail(SMLE_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    *out = a;
    return SMLE_OK;
}

template <typename V>
int spmm(smle_csr_t a, const V *X, V *Y, int k, int is_device_ptr)
{
    if (!a || !X || !Y || k < 1) return fail(SMLE_ERR_ARG, "smle_spmm: bad argument");
    if (a->vbytes != (int)sizeof(V)) return fail(SMLE_ERR_ARG, "value type of handle does not match call");
    int rc = ensure_init();
    if (rc) return rc;
    CgScalars none = {};
    if (is_device_ptr) return launch_merge<V, false>(a, X, Y, k, none);
    V *dX = nullptr, *dY = nullptr;
    size_t xb = sizeof(V) * (size_t)a->n * k, yb = sizeof(V) * (size_t)a->m * k;
    CU(cudaMalloc(&dX, xb ? xb : 16));
    cudaError_t e = cudaMalloc(&dY, yb ? yb : 16);
    if (e != cudaSuccess) { cudaFree(dX); return fail(SMLE_ERR_ALLOC, "device allocation failed"); }
    rc = SMLE_OK;
    if ((e = cudaMemcpyAsync(dX, X, xb, cudaMemcpyHostToDevice, g_stream)) != cudaSuccess)
        rc = fail(SMLE_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e));
    if (!rc) rc = launch_merge<V, false>(a, dX, dY, k, none);
    if (!rc && (e = cudaMemcpyAsync(Y, dY, yb, cudaMemcpyDeviceToHost, g_stream)) != cudaSuccess)
        rc = fail(SMLE_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(g_stream)) != cudaSuccess && !rc)
        rc = fail(SMLE_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e));
    cudaFree(dX); cudaFree(dY);
    return rc;
}

// ---- CG ---------------------------------------------------------------------------------------
constexpr int kGraphIters = 16;   // CG iterations per CUDA-graph launch

int ensure_workspace(smle_csr_t a, int k, int hist_cap)
{
    CgWorkspace &w = a->ws;
    if (w.k != k) {
        free_workspace(w);
        w.k = k;
        w.nk = (size_t)a->m * (size_t)k;
        size_t vb = sizeof(double) * (w.nk ? w.nk : 1);
        CU(cudaMalloc(&w.R, vb));
        CU(cudaMalloc(&w.P, vb));
        CU(cudaMalloc(&w.AP, vb));
        CU(cudaMalloc(&w.scal, sizeof(double) * (6 * (size_t)k + 2)));
        CU(cudaMalloc(&w.Xd, vb));
        CU(cudaMalloc(&w.conv, sizeof(int) * (size_t)k));
        CU(cudaMalloc(&w.ctrl, sizeof(int) * CTRL_WORDS));
        CU(cudaMalloc(&w.part, sizeof(double) * (size_t)g_sms * 8 * (size_t)k));
        CU(cudaMallocHost(&w.ctrl_host, sizeof(int) * CTRL_WORDS * 4));
    }
    if (hist_cap < 1) hist_cap = 1;
    if (hist_cap > w.hist_cap) {
        cudaFree(w.hist);
        w.hist = nullptr;
        CU(cudaMalloc(&w.hist, sizeof(double) * (size_t)hist_cap));
        w.hist_cap = hist_cap;
        if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; }
        if (w.pcg_graph) { cudaGraphExecDestroy(w.pcg_graph); w.pcg_graph = nullptr; }
        if (w.graph32) { cudaGraphExecDestroy(w.graph32); w.graph32 = nullptr; }
    }
    return SMLE_OK;
}

CgScalars make_scalars(CgWorkspace &w, int k)
{
    CgScalars s;
    s.rs_old = w.scal;
    s.rs_new = w.scal + k;
    s.pAp = w.scal + 2 * (size_t)k;
    s.alpha = w.scal + 3 * (size_t)k;
    s.beta = w.scal + 4 * (size_t)k;
    s.bnorm = w.scal + 5 * (size_t)k;
    s.last_rel = w.scal + 6 * (size_t)k;
    s.conv = w.conv;
    s.ctrl = w.ctrl;
    s.tol = w.scal + 6 * (size_t)k + 1;
    s.hist = w.hist;
    s.hist_cap = w.hist_cap;
    s.dot_mode = DOT_CG_ALPHA;
    return s;
}

template <typename V, int G, int VEC>
int launch_vec_t(int which, const CgVecArgsT<V> &va, const CgScalars &cg, int max_iters, double tol, int grid, int seq_base)
{
    if (which == 0) cg_init_kernel<V, G, VEC><<<grid, kThreads, 0, g_stream>>>(va, cg, max_iters, tol, seq_base);
    else if (which == 1) launch_kernel(cg_update_r_kernel<V, G, VEC>, dim3(grid), dim3(kThreads), 0, va, cg);
    else launch_kernel(cg_update_xp_kernel<V, G, VEC>, dim3(grid), dim3(kThreads), 0, va, cg);
    ++g_launches;
    return check_launch("cg vector kernel");
}

// fp32 blocks: the generic kernels only (the k = 1 specialisations and the multi-GPU / preconditioned loops are fp64)
int launch_vec(int which, const CgVecArgsT<float> &va, const CgScalars &cg, int max_iters = 0, double tol = 0.0, int seq_base = 0)
{
    int G, VEC;
    pick_shape<float>(va.k, &G, &VEC);
    const int W = kThreads / G;
    long long want = ((long long)va.n + W - 1) / W;
    int grid = (int)(want < (long long)g_sms * 8 ? want : (long long)g_sms * 8);
    if (grid < 1) grid = 1;
#define SMLE_CASE(g, v) if (G == g && VEC == v) return launch_vec_t<float, g, v>(which, va, cg, max_iters, tol, grid, seq_base);
    SMLE_CASE(1, 1) SMLE_CASE(2, 1) SMLE_CASE(4, 1) SMLE_CASE(8, 1) SMLE_CASE(16, 1) SMLE_CASE(32, 1)
    SMLE_CASE(1, 2) SMLE_CASE(2, 2) SMLE_CASE(4, 2) SMLE_CASE(8, 2) SMLE_CASE(16, 2) SMLE_CASE(32, 2)
    SMLE_CASE(1, 4) SMLE_CASE(2, 4) SMLE_CASE(4, 4) SMLE_CASE(8, 4) SMLE_CASE(16, 4) SMLE_CASE(32, 4)
#undef SMLE_CASE
    return fail(SMLE_ERR_ARG, "no vector kernel for k=%d", va.k);
}

int launch_vec(int which, const CgVecArgs &va, const CgScalars &cg, int max_iters = 0, double tol = 0.0, int seq_base = 0)
{
    if (va.k == 1 && which != 0 && getenv("SMLE_VEC_GENERIC") == nullptr) {
        // single right-hand side: 128-bit, unrolled kernels with L2 eviction priorities
        static int ctas_per_sm = -1;
        if (ctas_per_sm < 0) { const char *e = getenv("SMLE_VEC_CTAS"); ctas_per_sm = e ? atoi(e) : 3; }
        long long want = ((long long)(va.n >> 1) + kThreads * kVecUnroll - 1) / (kThreads * kVecUnroll);
        int grid = (int)(want < (long long)g_sms * ctas_per_sm ? want : (long long)g_sms * ctas_per_sm);
        if (grid < 1) grid = 1;
        // Under PDL these CTAs are scheduled while K1's CTAs drain one by one: without a limit the first
        // SMs to free up would swallow eight of them each.  An (unused) dynamic shared-memory request
        // pins the residency at ctas_per_sm, so the grid spreads evenly however early it arrives.
        static size_t pin = 0;
        if (!pin && g_pdl) {
            pin = ((size_t)227 * 1024 / (size_t)(ctas_per_sm > 0 ? ctas_per_sm : 1) - 8192) & ~(size_t)1023;
            CU(cudaFuncSetAttribute(cg1_update_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pin));
            CU(cudaFuncSetAttribute(cg1_update_xp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pin));
        }
        const size_t smem_pin = g_pdl ? pin : 0;
        if (which == 1) launch_kernel(cg1_update_r_kernel, dim3(grid), dim3(kThreads), smem_pin, va, cg);
        else launch_kernel(cg1_update_xp_kernel, dim3(grid), dim3(kThreads), smem_pin, va, cg);
        ++g_launches;
        return check_launch("cg1 vector kernel");
    }
    int G, VEC;
    pick_shape<double>(va.k, &G, &VEC);
    const int W = kThreads / G;
    long long want = ((long long)va.n + W - 1) / W;
    int grid = (int)(want < (long long)g_sms * 8 ? want : (long long)g_sms * 8);
    if (grid < 1) grid = 1;
#define SMLE_CASE(g, v) if (G == g && VEC == v) return launch_vec_t<double, g, v>(which, va, cg, max_iters, tol, grid, seq_base);
    SMLE_CASE(1, 1) SMLE_CASE(2, 1) SMLE_CASE(4, 1) SMLE_CASE(8, 1) SMLE_CASE(16, 1) SMLE_CASE(32, 1)
    SMLE_CASE(1, 2) SMLE_CASE(2, 2) SMLE_CASE(4, 2) SMLE_CASE(8, 2) SMLE_CASE(16, 2) SMLE_CASE(32, 2)
#undef SMLE_CASE
    return fail(SMLE_ERR_ARG, "no vector kernel for k=%d", va.k);
}

int launch_iteration(smle_csr_t a, const CgVecArgs &va, const CgScalars &cg)
{
    PdlScope pdl;
    int rc = launch_merge<double, true>(a, va.P, va.AP, va.k, cg);
    if (!rc) rc = launch_vec(1, va, cg);
    if (!rc) rc = launch_vec(2, va, cg);
    return rc;
}

// Batches of CG iterations with device-side convergence control.  `submit_batch()` queues `batch`
// iterations (a graph launch, or that many kernel triples); batch j+1 is queued before the control
// words of batch j are inspected, so the GPU never idles on the host round trip, and kernels queued
// behind a raised STOP flag are no-ops.
template <typename SubmitBatch>
int run_cg_batches(CgWorkspace &w, int max_iters, int batch, SubmitBatch submit_batch)
{
    if (max_iters <= 0) return SMLE_OK;
    cudaEvent_t ev[2];
    CU(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    auto submit = [&](int slot) -> int {
        int r2 = submit_batch();
        if (r2) return r2;
        CU(cudaMemcpyAsync(w.ctrl_host + slot * CTRL_WORDS, w.ctrl, sizeof(int) * CTRL_WORDS,
                           cudaMemcpyDeviceToHost, g_stream));
        CU(cudaEventRecord(ev[slot], g_stream));
        return SMLE_OK;
    };
    int launched = 0, j = 0;
    int rc = submit(0);
    launched += batch;
    while (!rc) {
        const bool more = launched < max_iters;
        if (more) {
            rc = submit((j + 1) & 1);
            launched += batch;
            if (rc) break;
        }
        cudaError_t e = cudaEventSynchronize(ev[j & 1]);
        if (e != cudaSuccess) { rc = fail(SMLE_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e)); break; }
        if (w.ctrl_host[(j & 1) * CTRL_WORDS + CTRL_STOP] || !more) break;
        ++j;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return rc;
}

// Solve A X = B.  B is a device pointer; the iterate lives in the workspace (w.Xd) so that the
// captured graph never depends on caller pointers; the result is copied to X_dev (device) or
// X_host (host) at the end.  tol < 0 never converges (fixed-count runs).
int cg_solve_device(smle_csr_t a, const double *B, double *X_dev, double *X_host, int k, int max_iters,
                    double tol, int *iters_out, double *hist_out, int hist_capacity, int *hist_len,
                    double *final_rel)
{
    const bool want_hist = hist_out != nullptr && hist_capacity > 0;
    int rc = ensure_workspace(a, k, want_hist ? (max_iters < hist_capacity ? max_iters : hist_capacity) : 0);
    if (rc) return rc;
    rc = ensure_scratch(a, k);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    CgScalars cg = make_scalars(w, k);
    CgVecArgs va;
    va.B = B; va.X = w.Xd; va.R = w.R; va.P = w.P; va.AP = w.AP;
    va.n = a->m; va.k = k; va.part = w.part; va.ticket = a->ticket + 1;

    rc = launch_vec(0, va, cg, max_iters, tol);
    if (rc) return rc;

    const bool use_graph = getenv("SMLE_NO_GRAPH") == nullptr;
    if (w.graph && w.graph_epoch != a->scratch_epoch) {   // scratch was reallocated since the capture
        cudaGraphExecDestroy(w.graph);
        w.graph = nullptr;
    }
    if (use_graph && !w.graph) {
        // lazy setup (partition kernel, occupancy query, scratch growth) must happen outside the capture
        rc = launch_merge<double, true>(a, va.P, va.AP, k, cg, /*dry=*/true);
        if (rc) return rc;
        cudaGraph_t graph;
        CU(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kGraphIters && !rc; ++i) rc = launch_iteration(a, va, cg);
        cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
        g_launches -= 3LL * kGraphIters;   // recorded, not launched
        if (rc) return rc;
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&w.graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        w.graph_iters = kGraphIters;
        w.graph_epoch = a->scratch_epoch;
    }

    const int batch = use_graph ? w.graph_iters : 4;
    rc = run_cg_batches(w, max_iters, batch, [&]() -> int {
        if (use_graph) {
            CU(cudaGraphLaunch(w.graph, g_stream));
            g_launches += 3LL * batch;
            return SMLE_OK;
        }
        for (int i = 0; i < batch; ++i) {
            int r2 = launch_iteration(a, va, cg);
            if (r2) return r2;
        }
        return SMLE_OK;
    });
    if (rc) return rc;

    // result + final state
    const size_t xb = sizeof(double) * w.nk;
    if (X_dev) CU(cudaMemcpyAsync(X_dev, w.Xd, xb, cudaMemcpyDeviceToDevice, g_stream));
    if (X_host) CU(cudaMemcpyAsync(X_host, w.Xd, xb, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host, w.ctrl, sizeof(int) * CTRL_WORDS, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host + CTRL_WORDS, cg.last_rel, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    int iters = w.ctrl_host[CTRL_ITER];
    if (iters_out) *iters_out = iters;
    if (final_rel) memcpy(final_rel, w.ctrl_host + CTRL_WORDS, sizeof(double));
    if (want_hist) {
        int nh = iters < w.hist_cap ? iters : w.hist_cap;
        if (nh > hist_capacity) nh = hist_capacity;
        CU(cudaMemcpyAsync(hist_out, w.hist, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        if (hist_len) *hist_len = nh;
    } else if (hist_len) {
        *hist_len = 0;
    }
    return SMLE_OK;
}

int cg_solve(smle_csr_t a, const double *B, double *X, int k, int max_iters, double tol,
             int is_device_ptr, int *iters_out, double *hist, int hist_capacity, int *hist_len,
             double *final_rel)
{
    if (!a || !B || !X || k < 1) return fail(SMLE_ERR_ARG, "smle_cg: bad argument");
    if (a->vbytes != 8) return fail(SMLE_ERR_ARG, "CG needs an fp64 handle (reference CG is <double,int>)");
    if (a->m != a->n) return fail(SMLE_ERR_ARG, "CG needs a square matrix");
    int rc = ensure_init();
    if (rc) return rc;
    if (is_device_ptr)
        return cg_solve_device(a, B, X, nullptr, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel);
    rc = ensure_workspace(a, k, 0);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    if (!w.Bd) CU(cudaMalloc(&w.Bd, sizeof(double) * (w.nk ? w.nk : 1)));
    CU(cudaMemcpyAsync(w.Bd, B, sizeof(double) * w.nk, cudaMemcpyHostToDevice, g_stream));
    return cg_solve_device(a, w.Bd, nullptr, X, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel);
}

// ---- fp32 CG: CGSolveSingle / CGSolveMultiple instantiated for <float,int> ----------
// Same three kernels per iteration with float blocks (the workspace buffers are reused, reinterpreted);
// scalars, dot partials and their reductions stay double.
int cg32_solve_device(smle_csr_t a, const float *B, float *X_dev, float *X_host, int k, int max_iters, double tol,
                      int *iters_out, double *hist_out, int hist_capacity, int *hist_len, double *final_rel)
{
    const bool want_hist = hist_out != nullptr && hist_capacity > 0;
    int rc = ensure_workspace(a, k, want_hist ? (max_iters < hist_capacity ? max_iters : hist_capacity) : 0);
    if (!rc) rc = ensure_scratch(a, k);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    CgScalars cg = make_scalars(w, k);
    CgVecArgsT<float> va;
    va.B = B; va.X = (float *)w.Xd; va.R = (float *)w.R; va.P = (float *)w.P; va.AP = (float *)w.AP;
    va.n = a->m; va.k = k; va.part = w.part; va.ticket = a->ticket + 1;
    rc = launch_vec(0, va, cg, max_iters, tol);
    if (rc) return rc;
    auto iteration = [&]() -> int {
        int r2 = launch_merge<float, true>(a, va.P, va.AP, k, cg);
        if (!r2) r2 = launch_vec(1, va, cg);
        if (!r2) r2 = launch_vec(2, va, cg);
        return r2;
    };
    const bool use_graph = getenv("SMLE_NO_GRAPH") == nullptr;
    if (w.graph32 && w.graph32_epoch != a->scratch_epoch) { cudaGraphExecDestroy(w.graph32); w.graph32 = nullptr; }
    if (use_graph && !w.graph32) {
        rc = launch_merge<float, true>(a, va.P, va.AP, k, cg, /*dry=*/true);
        if (rc) return rc;
        cudaGraph_t graph;
        CU(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kGraphIters && !rc; ++i) rc = iteration();
        cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
        g_launches -= 3LL * kGraphIters;
        if (rc) return rc;
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&w.graph32, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        w.graph32_epoch = a->scratch_epoch;
    }
    const int batch = use_graph ? kGraphIters : 4;
    rc = run_cg_batches(w, max_iters, batch, [&]() -> int {
        if (use_graph) {
            CU(cudaGraphLaunch(w.graph32, g_stream));
            g_launches += 3LL * batch;
            return SMLE_OK;
        }
        for (int i = 0; i < batch; ++i) {
            int r2 = iteration();
            if (r2) return r2;
        }
        return SMLE_OK;
    });
    if (rc) return rc;
    const size_t xb = sizeof(float) * w.nk;
    if (X_dev) CU(cudaMemcpyAsync(X_dev, w.Xd, xb, cudaMemcpyDeviceToDevice, g_stream));
    if (X_host) CU(cudaMemcpyAsync(X_host, w.Xd, xb, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host, w.ctrl, sizeof(int) * CTRL_WORDS, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host + CTRL_WORDS, cg.last_rel, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    const int iters = w.ctrl_host[CTRL_ITER];
    if (iters_out) *iters_out = iters;
    if (final_rel) memcpy(final_rel, w.ctrl_host + CTRL_WORDS, sizeof(double));
    if (want_hist) {
        int nh = iters < w.hist_cap ? iters : w.hist_cap;
        if (nh > hist_capacity) nh = hist_capacity;
        CU(cudaMemcpyAsync(hist_out, w.hist, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        if (hist_len) *hist_len = nh;
    } else if (hist_len) {
        *hist_len = 0;
    }
    return SMLE_OK;
}

int cg32_solve(smle_csr_t a, const float *B, float *X, int k, int max_iters, double tol, int is_device_ptr, int *iters_out,
               double *hist, int hist_capacity, int *hist_len, double *final_rel)
{
    if (!a || !B || !X || k < 1) return fail(SMLE_ERR_ARG, "smle_cg (fp32): bad argument");
    if (a->vbytes != 4) return fail(SMLE_ERR_ARG, "the fp32 solver needs an fp32 handle");
    if (a->m != a->n) return fail(SMLE_ERR_ARG, "CG needs a square matrix");
    int rc = ensure_init();
    if (rc) return rc;
    if (is_device_ptr) return cg32_solve_device(a, B, X, nullptr, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel);
    rc = ensure_workspace(a, k, 0);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    if (!w.Bd) CU(cudaMalloc(&w.Bd, sizeof(double) * (w.nk ? w.nk : 1)));
    CU(cudaMemcpyAsync(w.Bd, B, sizeof(float) * w.nk, cudaMemcpyHostToDevice, g_stream));
    return cg32_solve_device(a, (const float *)w.Bd, nullptr, X, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel);
}

// ---- SPAI-preconditioned multi-RHS CG ---------------------------------------------------------------
template <int G, int VEC>
int launch_pcg_vec_t(int which, const CgVecArgs &va, const CgScalars &cg, int grid)
{
    if (which == 0) pcg_update_xr_kernel<G, VEC><<<grid, kThreads, 0, g_stream>>>(va, cg);
    else pcg_update_p_kernel<G, VEC><<<grid, kThreads, 0, g_stream>>>(va, cg);
    ++g_launches;
    return check_launch("pcg vector kernel");
}

int launch_pcg_vec(int which, const CgVecArgs &va, const CgScalars &cg)
{
    int G, VEC;
    pick_shape<double>(va.k, &G, &VEC);
    const int W = kThreads / G;
    long long want = ((long long)va.n + W - 1) / W;
    int grid = (int)(want < (long long)g_sms * 8 ? want : (long long)g_sms * 8);
    if (grid < 1) grid = 1;
#define SMLE_CASE(g, v) if (G == g && VEC == v) return launch_pcg_vec_t<g, v>(which, va, cg, grid);
    SMLE_CASE(1, 1) SMLE_CASE(2, 1) SMLE_CASE(4, 1) SMLE_CASE(8, 1) SMLE_CASE(16, 1) SMLE_CASE(32, 1)
    SMLE_CASE(1, 2) SMLE_CASE(2, 2) SMLE_CASE(4, 2) SMLE_CASE(8, 2) SMLE_CASE(16, 2) SMLE_CASE(32, 2)
#undef SMLE_CASE
    return fail(SMLE_ERR_ARG, "no vector kernel for k=%d", va.k);
}

// z = M r with the fused r.z turned into rs_old / beta (mode), then p = z + beta p
int pcg_m_step(smle_csr_t m, const CgVecArgs &va, CgScalars cg, int mode, bool dry = false)
{
    cg.dot_mode = mode;
    int rc = launch_merge<double, true>(m, va.R, va.Z, va.k, cg, dry);
    if (!rc && !dry) rc = launch_pcg_vec(1, va, cg);
    return rc;
}

int pcg_iteration(smle_csr_t a, smle_csr_t m, const CgVecArgs &va, CgScalars cg)
{
    cg.dot_mode = DOT_PCG_ALPHA;
    int rc = launch_merge<double, true>(a, va.P, va.AP, va.k, cg);
    if (!rc) rc = launch_pcg_vec(0, va, cg);
    if (!rc) rc = pcg_m_step(m, va, cg, DOT_PCG_BETA);
    return rc;
}

int pcg_solve_device(smle_csr_t a, smle_csr_t m, const double *B, double *X_dev, double *X_host, int k, int max_iters,
                     double tol, int *iters_out, double *hist_out, int hist_capacity, int *hist_len, double *final_rel)
{
    const bool want_hist = hist_out != nullptr && hist_capacity > 0;
    int rc = ensure_workspace(a, k, want_hist ? (max_iters < hist_capacity ? max_iters : hist_capacity) : 0);
    if (!rc) rc = ensure_scratch(a, k);
    if (!rc) rc = ensure_scratch(m, k);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    if (!w.Z) CU(cudaMalloc(&w.Z, sizeof(double) * (w.nk ? w.nk : 1)));
    CgScalars cg = make_scalars(w, k);
    CgVecArgs va;
    va.B = B; va.X = w.Xd; va.R = w.R; va.P = w.P; va.AP = w.AP; va.Z = w.Z;
    va.n = a->m; va.k = k; va.part = w.part; va.ticket = a->ticket + 1;

    // x = 0, r = b, ||b||; z = M r, rs_old = r.z, p = z   (sparse_approximate_inverse.hpp:60-94)
    rc = launch_vec(0, va, cg, max_iters, tol);
    if (!rc) rc = pcg_m_step(m, va, cg, DOT_PCG_INIT);
    if (rc) return rc;

    const bool use_graph = getenv("SMLE_NO_GRAPH") == nullptr;
    if (w.pcg_graph && (w.pcg_m != (const void *)m || w.pcg_epoch_a != a->scratch_epoch || w.pcg_epoch_m != m->scratch_epoch)) {
        cudaGraphExecDestroy(w.pcg_graph);
        w.pcg_graph = nullptr;
    }
    if (use_graph && !w.pcg_graph) {
        CgScalars c1 = cg;
        c1.dot_mode = DOT_PCG_ALPHA;
        rc = launch_merge<double, true>(a, va.P, va.AP, k, c1, /*dry=*/true);
        if (!rc) rc = pcg_m_step(m, va, cg, DOT_PCG_BETA, /*dry=*/true);
        if (rc) return rc;
        cudaGraph_t graph;
        CU(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kGraphIters && !rc; ++i) rc = pcg_iteration(a, m, va, cg);
        cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
        g_launches -= 4LL * kGraphIters;
        if (rc) return rc;
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&w.pcg_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        w.pcg_m = m; w.pcg_epoch_a = a->scratch_epoch; w.pcg_epoch_m = m->scratch_epoch;
    }
    const int batch = use_graph ? kGraphIters : 4;
    rc = run_cg_batches(w, max_iters, batch, [&]() -> int {
        if (use_graph) {
            CU(cudaGraphLaunch(w.pcg_graph, g_stream));
            g_launches += 4LL * batch;
            return SMLE_OK;
        }
        for (int i = 0; i < batch; ++i) {
            int r2 = pcg_iteration(a, m, va, cg);
            if (r2) return r2;
        }
        return SMLE_OK;
    });
    if (rc) return rc;
    const size_t xb = sizeof(double) * w.nk;
    if (X_dev) CU(cudaMemcpyAsync(X_dev, w.Xd, xb, cudaMemcpyDeviceToDevice, g_stream));
    if (X_host) CU(cudaMemcpyAsync(X_host, w.Xd, xb, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host, w.ctrl, sizeof(int) * CTRL_WORDS, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host + CTRL_WORDS, cg.last_rel, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    const int iters = w.ctrl_host[CTRL_ITER];
    if (iters_out) *iters_out = iters;
    if (final_rel) memcpy(final_rel, w.ctrl_host + CTRL_WORDS, sizeof(double));
    if (want_hist) {
        int nh = iters < w.hist_cap ? iters : w.hist_cap;
        if (nh > hist_capacity) nh = hist_capacity;
        CU(cudaMemcpyAsync(hist_out, w.hist, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        if (hist_len) *hist_len = nh;
    } else if (hist_len) {
        *hist_len = 0;
    }
    return SMLE_OK;
}

} // namespace

// =========================================================================================
// exported C ABI
// =========================================================================================
extern "C" {

int smle_version(void) { return 100; }
const char *smle_last_error(void) { return g_err; }

int smle_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int smle_init(int device)
{
    int n = smle_device_count();
    if (n < 1) return fail(SMLE_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(SMLE_ERR_ARG, "device %d out of range (%d visible)", device, n);
    if (g_device == device) {   // already bound: keep the stream (cached CUDA graphs were captured on it)
        CU(cudaSetDevice(device));
        return SMLE_OK;
    }
    if (g_device >= 0)
        return fail(SMLE_ERR_ARG, "this process is bound to device %d (one process drives one GPU); "
                                  "destroy the handles and call smle_shutdown() before binding to device %d", g_device, device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    g_sms = prop.multiProcessorCount;
    const bool using_own = (g_stream == g_own_stream);
    CU(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    g_device = device;
    if (using_own) g_stream = g_own_stream;
    return SMLE_OK;
}

/* Handles must be destroyed first: they own CUDA graphs captured on the library's stream. */
void smle_shutdown(void)
{
    if (g_own_stream) { cudaStreamSynchronize(g_own_stream); cudaStreamDestroy(g_own_stream); }
    if (g_copy_stream) { cudaStreamSynchronize(g_copy_stream); cudaStreamDestroy(g_copy_stream); }
    g_own_stream = g_stream = g_copy_stream = nullptr;
    g_device = -1;
}

int smle_host_register(void *host_ptr, unsigned long long bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!host_ptr || !bytes) return SMLE_OK;
    cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 1; }   // already page-locked: nothing to undo
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SMLE_ERR_CUDA, "cudaHostRegister(%llu bytes) failed: %s", bytes, cudaGetErrorString(e)); }
    return SMLE_OK;
}

int smle_host_unregister(void *host_ptr)
{
    if (!host_ptr) return SMLE_OK;
    cudaError_t e = cudaHostUnregister(host_ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SMLE_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return SMLE_OK;
}

int smle_set_stream(void *s)
{
    int rc = ensure_init();
    if (rc) return rc;
    g_stream = s ? (cudaStream_t)s : g_own_stream;
    return SMLE_OK;
}

void *smle_get_stream(void) { return (void *)g_stream; }

int smle_sync(void)
{
    int rc = ensure_init();
    if (rc) return rc;
    CU(cudaStreamSynchronize(g_stream));
    return SMLE_OK;
}

long long smle_launch_count(void) { return g_launches; }
int smle_sm_count(void) { return ensure_init() ? 0 : g_sms; }

int smle_malloc(void **dev_ptr, unsigned long long bytes)
{
    if (!dev_ptr) return fail(SMLE_ERR_ARG, "null pointer");
    int rc = ensure_init();
    if (rc) return rc;
    cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(SMLE_ERR_ALLOC, "cudaMalloc(%llu) failed: %s", bytes, cudaGetErrorString(e));
    return SMLE_OK;
}

int smle_free(void *dev_ptr)
{
    if (g_stream) cudaStreamSynchronize(g_stream);
    CU(cudaFree(dev_ptr));
    return SMLE_OK;
}

int smle_copy_to_device(void *dev_dst, const void *host_src, unsigned long long bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    CU(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return SMLE_OK;
}

int smle_copy_to_host(void *host_dst, const void *dev_src, unsigned long long bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return SMLE_OK;
}

int smle_merge_path_partition(const int *row_end, int m, int nnz, int num_parts, int items_per_part,
                              int *out_xy)
{
    if (!row_end && m > 0) return fail(SMLE_ERR_ARG, "row_end_offsets is NULL");
    if (m < 0 || nnz < 0 || num_parts < 1 || !out_xy) return fail(SMLE_ERR_ARG, "bad argument");
    long long total = (long long)m + nnz;
    if (total > INT_MAX) return fail(SMLE_ERR_RANGE, "m + nnz overflows int");
    int rc = ensure_init();
    if (rc) return rc;
    int share = items_per_part > 0 ? items_per_part : (int)((total + num_parts - 1) / num_parts);
    int *d_row_end = nullptr;
    int2 *d_out = nullptr;
    CU(cudaMalloc(&d_row_end, sizeof(int) * (size_t)(m > 0 ? m : 1)));
    cudaError_t e = cudaMalloc(&d_out, sizeof(int2) * (size_t)(num_parts + 1));
    if (e != cudaSuccess) { cudaFree(d_row_end); return fail(SMLE_ERR_ALLOC, "device allocation failed"); }
    rc = SMLE_OK;
    if ((e = cudaMemcpyAsync(d_row_end, row_end, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, g_stream)) != cudaSuccess)
        rc = fail(SMLE_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e));
    if (!rc) {
        merge_partition_kernel<<<(num_parts + 1 + 127) / 128, 128, 0, g_stream>>>(d_row_end, m, nnz, share,
                                                                                  num_parts, d_out);
        ++g_launches;
        rc = check_launch("merge_partition_kernel");
    }
    if (!rc && (e = cudaMemcpyAsync(out_xy, d_out, sizeof(int2) * (size_t)(num_parts + 1),
                                    cudaMemcpyDeviceToHost, g_stream)) != cudaSuccess)
        rc = fail(SMLE_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(g_stream)) != cudaSuccess && !rc)
        rc = fail(SMLE_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e));
    cudaFree(d_row_end); cudaFree(d_out);
    return rc;
}

int smle_csr_create_f64(smle_csr_t *out, int m, int n, int nnz, const int *ro, const int *ci, const double *va)
{
    return csr_create<double>(out, m, n, nnz, ro, ci, va);
}
int smle_csr_create_f32(smle_csr_t *out, int m, int n, int nnz, const int *ro, const int *ci, const float *va)
{
    return csr_create<float>(out, m, n, nnz, ro, ci, va);
}

void smle_csr_destroy(smle_csr_t a)
{
    if (!a) return;
    if (g_stream) cudaStreamSynchronize(g_stream);
    free_workspace(a->ws);
    for (auto &kv : a->parts) {
        cudaFree(kv.second.xy); cudaFree(kv.second.maxlen); cudaFree(kv.second.halo);
        for (auto &sc : kv.second.scheds) { cudaFree(sc.second.sched); cudaFree(sc.second.off); }
    }
    cudaFree(a->ro); cudaFree(a->ci); cudaFree(a->va);
    cudaFree(a->carry_row); cudaFree(a->carry_val); cudaFree(a->dot_part); cudaFree(a->fix_part); cudaFree(a->dot_sum);
    cudaFree(a->ticket); cudaFree(a->tile_carry); cudaFree(a->cta_slot);
    delete a;
}

int smle_csr_dims(smle_csr_t a, int *m, int *n, int *nnz, int *value_bytes)
{
    if (!a) return fail(SMLE_ERR_ARG, "null handle");
    if (m) *m = a->m;
    if (n) *n = a->n;
    if (nnz) *nnz = a->nnz;
    if (value_bytes) *value_bytes = a->vbytes;
    return SMLE_OK;
}

int smle_csr_tile_coords(smle_csr_t a, int k, int *num_tiles, int *items_per_tile, int *out_xy, int capacity)
{
    if (!a || k < 1) return fail(SMLE_ERR_ARG, "bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    int items = 0;
    rc = spmv_tile_items(a, &items);
    if (rc) return rc;
    if (k > 1) {   // the tiling of the SpMM kernel that would run for this (matrix, k)
        int G, VEC;
        if (a->vbytes == 8) pick_shape<double>(k, &G, &VEC); else pick_shape<float>(k, &G, &VEC);
        bool rows_kernel = false;
        rc = spmm_use_rows(a, G, VEC, k, false, &rows_kernel);
        if (rc) return rc;
        items = rows_kernel ? spmm_tile_items(G) : kTileItems;
    }
    Partition *p;
    rc = get_partition(a, items, &p);
    if (rc) return rc;
    if (num_tiles) *num_tiles = p->num_tiles;
    if (items_per_tile) *items_per_tile = p->items_per_tile;
    if (out_xy) {
        if (capacity < 2 * (p->num_tiles + 1)) return fail(SMLE_ERR_ARG, "out_xy too small");
        CU(cudaMemcpyAsync(out_xy, p->xy, sizeof(int2) * (size_t)(p->num_tiles + 1), cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
    }
    return SMLE_OK;
}

int smle_spmv_f64(smle_csr_t a, const double *x, double *y, int dev) { return spmm<double>(a, x, y, 1, dev); }
int smle_spmv_f32(smle_csr_t a, const float *x, float *y, int dev) { return spmm<float>(a, x, y, 1, dev); }
int smle_spmm_f64(smle_csr_t a, const double *X, double *Y, int k, int dev) { return spmm<double>(a, X, Y, k, dev); }
int smle_spmm_f32(smle_csr_t a, const float *X, float *Y, int k, int dev) { return spmm<float>(a, X, Y, k, dev); }

int smle_cg_single_f64(smle_csr_t a, const double *b, double *x, int max_iters, double tol, int dev,
                       int *iters_out, double *final_rel_res)
{
    return cg_solve(a, b, x, 1, max_iters, tol, dev, iters_out, nullptr, 0, nullptr, final_rel_res);
}

int smle_cg_single_batch_f64(smle_csr_t a, const double *b_vectors, double *x_solutions, int num_vectors,
                             int max_iters, double tol, int *iters_each, long long *iters_total)
{
    if (!a || !b_vectors || !x_solutions || num_vectors < 0) return fail(SMLE_ERR_ARG, "smle_cg_single_batch: bad argument");
    if (a->vbytes != 8) return fail(SMLE_ERR_ARG, "CG needs an fp64 handle (reference CG is <double,int>)");
    if (a->m != a->n) return fail(SMLE_ERR_ARG, "CG needs a square matrix");
    int rc = ensure_init();
    if (!rc) rc = ensure_workspace(a, 1, 0);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    const size_t n = (size_t)a->m, vb = sizeof(double) * (n ? n : 1);
    for (int i = 0; i < 2; ++i)
        if (!w.Bd2[i]) CU(cudaMalloc(&w.Bd2[i], vb));
    if (!w.Xs) CU(cudaMalloc(&w.Xs, vb));
    cudaStream_t copy_stream = g_copy_stream;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // b_ready[0], b_ready[1], x_ready, x_free
    for (auto &e : ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            for (auto e2 : ev) if (e2) cudaEventDestroy(e2);
            return fail(SMLE_ERR_CUDA, "cudaEventCreate failed");
        }
    cudaEvent_t *b_ready = ev, x_ready = ev[2], x_free = ev[3];

    long long total = 0;
    // the loop body returns through `step`, so the events are destroyed on every path
    auto step = [&](int v) -> int {
        // b_v is on the device when the solve starts; b_{v+1} follows on the copy stream meanwhile (its
        // buffer was last read by solve v-1, which this thread has already waited for)
        CU(cudaStreamWaitEvent(g_stream, b_ready[v & 1], 0));
        if (v + 1 < num_vectors) {
            CU(cudaMemcpyAsync(w.Bd2[(v + 1) & 1], b_vectors + (size_t)(v + 1) * n, sizeof(double) * n,
                               cudaMemcpyHostToDevice, copy_stream));
            CU(cudaEventRecord(b_ready[(v + 1) & 1], copy_stream));
        }
        if (v > 0) CU(cudaStreamWaitEvent(g_stream, x_free, 0));   // x_{v-1} has left the staging buffer
        int iters = 0;
        int r2 = cg_solve_device(a, w.Bd2[v & 1], w.Xs, nullptr, 1, max_iters, tol, &iters, nullptr, 0, nullptr, nullptr);
        if (r2) return r2;
        // x_v travels to the host while system v+1 is being solved
        CU(cudaEventRecord(x_ready, g_stream));
        CU(cudaStreamWaitEvent(copy_stream, x_ready, 0));
        CU(cudaMemcpyAsync(x_solutions + (size_t)v * n, w.Xs, sizeof(double) * n, cudaMemcpyDeviceToHost, copy_stream));
        CU(cudaEventRecord(x_free, copy_stream));
        if (iters_each) iters_each[v] = iters;
        total += iters;
        return SMLE_OK;
    };
    rc = SMLE_OK;
    if (num_vectors > 0) {
        cudaError_t e0 = cudaMemcpyAsync(w.Bd2[0], b_vectors, sizeof(double) * n, cudaMemcpyHostToDevice, copy_stream);
        if (e0 == cudaSuccess) e0 = cudaEventRecord(b_ready[0], copy_stream);
        if (e0 != cudaSuccess) rc = fail(SMLE_ERR_CUDA, "upload of b failed: %s", cudaGetErrorString(e0));
    }
    for (int v = 0; v < num_vectors && !rc; ++v) rc = step(v);
    cudaError_t e = cudaStreamSynchronize(copy_stream);
    if (!rc && e != cudaSuccess) rc = fail(SMLE_ERR_CUDA, "copy stream failed: %s", cudaGetErrorString(e));
    for (auto e2 : ev) cudaEventDestroy(e2);
    if (iters_total) *iters_total = total;
    return rc;
}

int smle_cg_multi_f64(smle_csr_t a, const double *B, double *X, int k, int max_iters, double tol, int kernel,
                      int dev, int *iters_out, double *hist, int hist_capacity, int *hist_len,
                      double *final_rel_res)
{
    if (kernel < SMLE_SIMPLE || kernel > SMLE_NONZERO_SPLIT) return fail(SMLE_ERR_ARG, "unknown SpmmKernel %d", kernel);
    return cg_solve(a, B, X, k, max_iters, tol, dev, iters_out, hist, hist_capacity, hist_len, final_rel_res);
}

int smle_pcg_spai_multi_f64(smle_csr_t a, smle_csr_t m, const double *B, double *X, int k, int max_iters, double tol,
                            int kernel, int dev, int *iters_out, double *hist, int hist_capacity, int *hist_len,
                            double *final_rel_res)
{
    if (!a || !m || !B || !X || k < 1) return fail(SMLE_ERR_ARG, "smle_pcg_spai: bad argument");
    if (kernel < SMLE_SIMPLE || kernel > SMLE_NONZERO_SPLIT) return fail(SMLE_ERR_ARG, "unknown SpmmKernel %d", kernel);
    if (a->vbytes != 8 || m->vbytes != 8) return fail(SMLE_ERR_ARG, "PCG needs fp64 handles");
    if (a->m != a->n || m->m != a->m || m->n != a->n) return fail(SMLE_ERR_ARG, "A and M must be square and of the same size");
    int rc = ensure_init();
    if (rc) return rc;
    if (dev) return pcg_solve_device(a, m, B, X, nullptr, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel_res);
    rc = ensure_workspace(a, k, 0);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    if (!w.Bd) CU(cudaMalloc(&w.Bd, sizeof(double) * (w.nk ? w.nk : 1)));
    CU(cudaMemcpyAsync(w.Bd, B, sizeof(double) * w.nk, cudaMemcpyHostToDevice, g_stream));
    return pcg_solve_device(a, m, w.Bd, nullptr, X, k, max_iters, tol, iters_out, hist, hist_capacity, hist_len, final_rel_res);
}

int smle_cg_single_f32(smle_csr_t a, const float *b, float *x, int max_iters, float tol, int dev, int *iters_out,
                       double *final_rel_res)
{
    return cg32_solve(a, b, x, 1, max_iters, (double)tol, dev, iters_out, nullptr, 0, nullptr, final_rel_res);
}

int smle_cg_multi_f32(smle_csr_t a, const float *B, float *X, int k, int max_iters, float tol, int kernel, int dev,
                      int *iters_out, double *hist, int hist_capacity, int *hist_len, double *final_rel_res)
{
    if (kernel < SMLE_SIMPLE || kernel > SMLE_NONZERO_SPLIT) return fail(SMLE_ERR_ARG, "unknown SpmmKernel %d", kernel);
    return cg32_solve(a, B, X, k, max_iters, (double)tol, dev, iters_out, hist, hist_capacity, hist_len, final_rel_res);
}

int smle_cg_run_fixed_f64(smle_csr_t a, const double *B, double *X, int k, int iters)
{
    return cg_solve(a, B, X, k, iters, -1.0, 1, nullptr, nullptr, 0, nullptr, nullptr);
}

int smle_cg_profile_f64(smle_csr_t a, const double *B, double *X, int k, int iters, float *ms_per_kernel)
{
    if (!a || !B || !X || k < 1 || iters < 1 || !ms_per_kernel) return fail(SMLE_ERR_ARG, "bad argument");
    if (a->vbytes != 8 || a->m != a->n) return fail(SMLE_ERR_ARG, "CG needs a square fp64 handle");
    int rc = ensure_init();
    if (rc) return rc;
    rc = ensure_workspace(a, k, 0);
    if (!rc) rc = ensure_scratch(a, k);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    CgScalars cg = make_scalars(w, k);
    CgVecArgs va;
    va.B = B; va.X = w.Xd; va.R = w.R; va.P = w.P; va.AP = w.AP;
    va.n = a->m; va.k = k; va.part = w.part; va.ticket = a->ticket + 1;
    (void)X;
    rc = launch_merge<double, true>(a, va.P, va.AP, k, cg, /*dry=*/true);
    if (!rc) rc = launch_vec(0, va, cg, iters + 2, -1.0);
    if (rc) return rc;
    std::vector<cudaEvent_t> ev((size_t)iters * 4);
    for (auto &e : ev) CU(cudaEventCreate(&e));
    rc = launch_iteration(a, va, cg);   // warm-up iteration, untimed
    for (int i = 0; i < iters && !rc; ++i) {
        CU(cudaEventRecord(ev[(size_t)i * 4 + 0], g_stream));
        rc = launch_merge<double, true>(a, va.P, va.AP, k, cg);
        CU(cudaEventRecord(ev[(size_t)i * 4 + 1], g_stream));
        if (!rc) rc = launch_vec(1, va, cg);
        CU(cudaEventRecord(ev[(size_t)i * 4 + 2], g_stream));
        if (!rc) rc = launch_vec(2, va, cg);
        CU(cudaEventRecord(ev[(size_t)i * 4 + 3], g_stream));
    }
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (!rc && e != cudaSuccess) rc = fail(SMLE_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e));
    double acc[3] = {0, 0, 0};
    if (!rc) {
        for (int i = 0; i < iters; ++i)
            for (int j = 0; j < 3; ++j) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)i * 4 + j], ev[(size_t)i * 4 + j + 1]);
                acc[j] += ms;
            }
        for (int j = 0; j < 3; ++j) ms_per_kernel[j] = (float)(acc[j] / iters);
    }
    for (auto &ev1 : ev) cudaEventDestroy(ev1);
    return rc;
}

} // extern "C"

// =========================================================================================
// row-partitioned CG over NVLink peer memory (one process per GPU; smle_dist.cuh, smle_plan.cpp)
// =========================================================================================
struct smle_dist_s {
    smle_csr_t a = nullptr;            // local rows, columns remapped to [own | pad | halo] (owned)
    int rank = 0, world = 1, n_local = 0, n_halo = 0, halo_base = 0;
    unsigned char *comm = nullptr;     // [DistBlock (4 KB) | p vector (halo_base + n_halo doubles)]
    void *peer_base[kMaxRanks] = {};
    int *send_idx = nullptr;
    unsigned int *ticket = nullptr;
    DistCtl ctl;
    DistCtl *ctl_dev = nullptr;        // device copy (read by the SpMV kernel)
    int seq_base = 0;
    cudaGraphExec_t graph = nullptr;
    unsigned graph_epoch = 0;
    bool connected = false;
    double *b_stage = nullptr, *x_stage = nullptr;   // host-pointer calls
};

namespace {

double *dist_p(smle_dist_t d) { return (double *)(d->comm + kDistCtlBytes); }

int dist_vec_grid(int n)
{
    long long want = ((long long)(n >> 1) + kThreads * kVecUnroll - 1) / (kThreads * kVecUnroll);
    int grid = (int)(want < (long long)g_sms * 2 ? want : (long long)g_sms * 2);
    return grid < 1 ? 1 : grid;
}

int dist_push_grid(smle_dist_t d)
{
    int total = d->ctl.send_off[d->world];
    int pgrid = (total + kThreads - 1) / kThreads;
    if (pgrid < 1) pgrid = 1;
    return pgrid > g_sms ? g_sms : pgrid;
}

int dist_launch_iteration(smle_dist_t d, const CgVecArgs &va, const CgScalars &cg)
{
    // K1 local SpMV (boundary tiles wait for the halo) + p.Ap posted to the peers |
    // K2 all-reduce -> alpha, r update, r.r posted |
    // K3 all-reduce -> beta, x/p update + halo push, iteration state | [separate halo push kernel]
    g_spmv_dist = d->ctl_dev;
    int rc = launch_merge<double, true>(d->a, va.P, va.AP, 1, cg);
    g_spmv_dist = nullptr;
    if (rc) return rc;
    const int grid = dist_vec_grid(va.n);
    cg1d_update_r_kernel<<<grid, kThreads, 0, g_stream>>>(va, cg, d->ctl);
    cg1d_update_xp_kernel<<<grid, kThreads, 0, g_stream>>>(va, cg, d->ctl, 0);
    if (!d->ctl.fused) dist_halo_push_kernel<<<dist_push_grid(d), kThreads, 0, g_stream>>>(d->ctl, va.P, cg.ctrl);
    g_launches += d->ctl.fused ? 2 : 3;
    return check_launch("distributed CG iteration");
}

int dist_create(smle_dist_t *out, smle_csr_t local_a, int rank, int world, int n_local, int n_halo, int halo_base,
                const int *send_off, const int *send_idx, const int *send_dst, const int *needs_from)
{
    smle_dist_s *d = new (std::nothrow) smle_dist_s();
    if (!d) return fail(SMLE_ERR_ALLOC, "out of host memory");
    *out = d;   // the caller destroys it on failure
    d->a = local_a; d->rank = rank; d->world = world; d->n_local = n_local; d->n_halo = n_halo; d->halo_base = halo_base;
    local_a->halo_base = halo_base;   // partitions of this handle flag the tiles that gather halo columns
    size_t bytes = kDistCtlBytes + sizeof(double) * ((size_t)halo_base + n_halo + 2);
    CU(cudaMalloc(&d->comm, bytes));
    CU(cudaMemsetAsync(d->comm, 0, bytes, g_stream));
    int total = send_off[world];
    CU(cudaMalloc(&d->send_idx, sizeof(int) * (size_t)(total > 0 ? total : 1)));
    if (total > 0) CU(cudaMemcpyAsync(d->send_idx, send_idx, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, g_stream));
    CU(cudaMalloc(&d->ticket, sizeof(unsigned int) * 4));
    CU(cudaMemsetAsync(d->ticket, 0, sizeof(unsigned int) * 4, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    int rc = ensure_workspace(local_a, 1, 0);
    if (rc) return rc;
    DistCtl &c = d->ctl;
    memset(&c, 0, sizeof(c));
    c.rank = rank; c.world = world;
    c.self = (DistBlock *)d->comm;
    c.send_idx = d->send_idx;
    for (int q = 0; q <= world; ++q) c.send_off[q] = send_off[q];
    for (int q = 0; q < world; ++q) { c.send_dst[q] = send_dst[q]; c.needs_from[q] = needs_from[q]; }
    c.ticket = d->ticket;
    c.stop = local_a->ws.ctrl + CTRL_STOP;
    {   // fused push: possible when every send group is one ascending run of consecutive local rows
        static int allow = -1;
        if (allow < 0) { const char *e = getenv("SMLE_DIST_FUSED_PUSH"); allow = e ? atoi(e) : 1; }
        c.fused = allow;
        c.npush = 0;
        for (int q = 0; q < world && c.fused; ++q) {
            const int cnt = send_off[q + 1] - send_off[q];
            if (cnt <= 0) continue;
            for (int i = 1; i < cnt; ++i)
                if (send_idx[send_off[q] + i] != send_idx[send_off[q]] + i) { c.fused = 0; break; }
            c.push_q[c.npush] = q;
            c.push_lo[c.npush] = send_idx[send_off[q]];
            c.push_cnt[c.npush] = cnt;
            ++c.npush;
        }
        if (!c.fused) c.npush = 0;
    }
    c.peer[rank] = c.self;
    c.peer_p[rank] = dist_p(d);
    d->peer_base[rank] = d->comm;
    d->connected = false;   // smle_dist_connect finishes the setup (also for world == 1)
    return SMLE_OK;
}

} // namespace

extern "C" {

int smle_dist_bounds(const int *row_offsets, int m, int world, int *bounds)
{
    if (!row_offsets || m < 0 || world < 1 || !bounds) return fail(SMLE_ERR_ARG, "smle_dist_bounds: bad argument");
    std::vector<int> xy(2 * ((size_t)world + 1));
    int rc = smle_merge_path_partition(row_offsets + 1, m, row_offsets[m], world, 0, xy.data());
    if (rc) return rc;
    for (int g = 0; g <= world; ++g) bounds[g] = xy[2 * (size_t)g];   // a row cut mid-way belongs whole to the later part
    bounds[0] = 0;
    bounds[world] = m;
    return SMLE_OK;
}

int smle_dist_create_from_plan(smle_dist_t *out, smle_plan_t plan, const double *local_values)
{
    if (!out || !plan) return fail(SMLE_ERR_ARG, "smle_dist_create_from_plan: bad argument");
    const smle_plan_s &p = *plan;
    if (!p.finished) return fail(SMLE_ERR_ARG, "smle_dist_plan_finish has not been called");
    if (p.world > kMaxRanks) return fail(SMLE_ERR_ARG, "world %d exceeds the %d ranks of the peer-memory control block", p.world, kMaxRanks);
    if (p.nnz_local > 0 && !local_values) return fail(SMLE_ERR_ARG, "local_values is NULL");
    int rc = ensure_init();
    if (rc) return rc;
    smle_csr_t a = nullptr;
    rc = csr_create<double>(&a, p.n_local, p.halo_base + p.n_halo, p.nnz_local, p.lro.data(), p.lci.data(), local_values);
    if (rc) return rc;
    smle_dist_t d = nullptr;
    rc = dist_create(&d, a, p.rank, p.world, p.n_local, p.n_halo, p.halo_base, p.send_off.data(), p.send_idx.data(),
                     p.send_dst.data(), p.needs_from.data());
    if (rc) {
        if (d) smle_dist_destroy(d); else smle_csr_destroy(a);
        return rc;
    }
    *out = d;
    return SMLE_OK;
}

int smle_dist_ipc_handle(smle_dist_t d, unsigned char *out64)
{
    if (!d || !out64) return fail(SMLE_ERR_ARG, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d->comm));
    memcpy(out64, &h, 64);
    return SMLE_OK;
}

int smle_dist_connect(smle_dist_t d, const unsigned char *all_handles)
{
    if (!d || (!all_handles && d->world > 1)) return fail(SMLE_ERR_ARG, "bad argument");
    for (int q = 0; q < d->world; ++q) {
        if (q == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)q * 64, 64);
        void *base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(SMLE_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
        d->peer_base[q] = base;
        d->ctl.peer[q] = (DistBlock *)base;
        d->ctl.peer_p[q] = (double *)((unsigned char *)base + kDistCtlBytes);
    }
    if (!d->ctl_dev) CU(cudaMalloc(&d->ctl_dev, sizeof(DistCtl)));
    CU(cudaMemcpy(d->ctl_dev, &d->ctl, sizeof(DistCtl), cudaMemcpyHostToDevice));
    d->connected = true;
    return SMLE_OK;
}

void smle_dist_destroy(smle_dist_t d)
{
    if (!d) return;
    if (g_stream) cudaStreamSynchronize(g_stream);
    if (d->graph) cudaGraphExecDestroy(d->graph);
    for (int q = 0; q < d->world; ++q)
        if (q != d->rank && d->peer_base[q]) cudaIpcCloseMemHandle(d->peer_base[q]);
    cudaFree(d->comm); cudaFree(d->send_idx); cudaFree(d->ticket); cudaFree(d->ctl_dev);
    cudaFree(d->b_stage); cudaFree(d->x_stage);
    if (d->a) smle_csr_destroy(d->a);
    delete d;
}

int smle_dist_dims(smle_dist_t d, int *n_local, int *n_halo, int *rank, int *world)
{
    if (!d) return fail(SMLE_ERR_ARG, "null handle");
    if (n_local) *n_local = d->n_local;
    if (n_halo) *n_halo = d->n_halo;
    if (rank) *rank = d->rank;
    if (world) *world = d->world;
    return SMLE_OK;
}

// y_local = (A x)_local : pushes the halo of x, then the local merge-path SpMV whose boundary tiles
// wait for the neighbours' pushes.  Collective: every rank of the partition must call it.
int smle_dist_spmv_f64(smle_dist_t d, const double *x_local_dev, double *y_local_dev)
{
    if (!d || !x_local_dev || !y_local_dev) return fail(SMLE_ERR_ARG, "bad argument");
    if (!d->connected) return fail(SMLE_ERR_COMM, "smle_dist_connect has not been called");
    int rc = ensure_workspace(d->a, 1, 0);
    if (rc) return rc;
    CgWorkspace &w = d->a->ws;
    int ctrl[CTRL_WORDS] = {0, 0, 0, 1, 0, d->seq_base, 0, 0};
    CU(cudaMemcpyAsync(w.ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(dist_p(d), x_local_dev, sizeof(double) * (size_t)d->n_local, cudaMemcpyDeviceToDevice, g_stream));
    dist_halo_push_kernel<<<dist_push_grid(d), kThreads, 0, g_stream>>>(d->ctl, dist_p(d), w.ctrl);
    ++g_launches;
    rc = check_launch("dist_halo_push_kernel");
    if (rc) return rc;
    d->seq_base += 2;
    CgScalars cg = make_scalars(w, 1);   // the boundary tiles read the sequence base from ctrl
    g_spmv_dist = d->ctl_dev;
    rc = launch_merge<double, false>(d->a, dist_p(d), y_local_dev, 1, cg);
    g_spmv_dist = nullptr;
    if (rc) return rc;
    int err = 0;
    CU(cudaMemcpyAsync(&err, &d->ctl.self->error, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    if (err) return fail(SMLE_ERR_COMM, "a peer did not answer (halo wait timed out)");
    return SMLE_OK;
}

// Row-partitioned CGSolveSingle (single_strategy.hpp:105-170 semantics on the global system).
// b_local / x_local: this rank's rows (device pointers, or host pointers with is_device_ptr = 0).
// Collective over the partition.
int smle_dist_cg_f64(smle_dist_t d, const double *b_local, double *x_local, int max_iters, double tol,
                     int is_device_ptr, int *iters_out, double *final_rel_res)
{
    if (!d || !b_local || !x_local) return fail(SMLE_ERR_ARG, "bad argument");
    if (!d->connected) return fail(SMLE_ERR_COMM, "smle_dist_connect has not been called");
    smle_csr_t a = d->a;
    int rc = ensure_workspace(a, 1, 0);
    if (!rc) rc = ensure_scratch(a, 1);
    if (rc) return rc;
    const size_t vb = sizeof(double) * (size_t)(d->n_local > 0 ? d->n_local : 1);
    const double *b_dev = b_local;
    double *x_dev = x_local;
    if (!is_device_ptr) {
        if (!d->b_stage) CU(cudaMalloc(&d->b_stage, vb));
        if (!d->x_stage) CU(cudaMalloc(&d->x_stage, vb));
        CU(cudaMemcpyAsync(d->b_stage, b_local, sizeof(double) * (size_t)d->n_local, cudaMemcpyHostToDevice, g_stream));
        b_dev = d->b_stage;
        x_dev = d->x_stage;
    }
    CgWorkspace &w = a->ws;
    CgScalars cg = make_scalars(w, 1);
    CgVecArgs va;
    va.B = b_dev; va.X = w.Xd; va.R = w.R; va.P = dist_p(d); va.AP = w.AP;
    va.n = d->n_local; va.k = 1; va.part = w.part; va.ticket = a->ticket + 1;

    // init: x = 0, r = p = b, local b.b -> all-reduce -> rs_old, bnorm; first halo push
    rc = launch_vec(0, va, cg, max_iters, tol, d->seq_base);
    if (rc) return rc;
    dist_post_kernel<<<1, 32, 0, g_stream>>>(d->ctl, 2, cg.rs_old, cg.ctrl);
    cg1d_update_xp_kernel<<<1, kThreads, 0, g_stream>>>(va, cg, d->ctl, 1);
    dist_halo_push_kernel<<<dist_push_grid(d), kThreads, 0, g_stream>>>(d->ctl, va.P, cg.ctrl);
    g_launches += 3;
    rc = check_launch("distributed CG init");
    if (rc) return rc;

    const bool use_graph = getenv("SMLE_NO_GRAPH") == nullptr;
    if (d->graph && d->graph_epoch != a->scratch_epoch) {   // scratch was reallocated since the capture
        cudaGraphExecDestroy(d->graph);
        d->graph = nullptr;
    }
    if (use_graph && !d->graph) {
        g_spmv_dist = d->ctl_dev;   // the dry run must size the same partition (with halo flags) the solve uses
        rc = launch_merge<double, true>(a, va.P, va.AP, 1, cg, /*dry=*/true);
        g_spmv_dist = nullptr;
        if (rc) return rc;
        cudaGraph_t graph;
        CU(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kGraphIters && !rc; ++i) rc = dist_launch_iteration(d, va, cg);
        cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
        g_launches -= (d->ctl.fused ? 3LL : 4LL) * kGraphIters;
        if (rc) return rc;
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&d->graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        d->graph_epoch = a->scratch_epoch;
    }
    const int batch = use_graph ? kGraphIters : 4;
    const long long per_iter = d->ctl.fused ? 3LL : 4LL;
    rc = run_cg_batches(w, max_iters, batch, [&]() -> int {
        if (use_graph) {
            CU(cudaGraphLaunch(d->graph, g_stream));
            g_launches += per_iter * batch;
            return SMLE_OK;
        }
        for (int i = 0; i < batch; ++i) {
            int r2 = dist_launch_iteration(d, va, cg);
            if (r2) return r2;
        }
        return SMLE_OK;
    });
    if (rc) return rc;
    if (is_device_ptr) CU(cudaMemcpyAsync(x_dev, w.Xd, sizeof(double) * (size_t)d->n_local, cudaMemcpyDeviceToDevice, g_stream));
    else CU(cudaMemcpyAsync(x_local, w.Xd, sizeof(double) * (size_t)d->n_local, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host, w.ctrl, sizeof(int) * CTRL_WORDS, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host + CTRL_WORDS, cg.last_rel, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(w.ctrl_host + CTRL_WORDS + 2, &d->ctl.self->error, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    const int iters = w.ctrl_host[CTRL_ITER];
    d->seq_base += iters + 2;
    if (iters_out) *iters_out = iters;
    if (final_rel_res) memcpy(final_rel_res, w.ctrl_host + CTRL_WORDS, sizeof(double));
    if (w.ctrl_host[CTRL_WORDS + 2]) return fail(SMLE_ERR_COMM, "a peer did not answer (wait timed out)");
    return SMLE_OK;
}

// Per-kernel timing of the row-partitioned iteration (roofline report): `iters` iterations WITHOUT
// the CUDA graph, CUDA events around K1 (local SpMV + p.Ap post), K2, K3 (+ halo push).  The waits
// for the peers are inside the kernels, so their cost shows up in the kernel that waits.  Collective.
int smle_dist_cg_profile_f64(smle_dist_t d, const double *b_local_dev, int iters, float *ms_per_kernel)
{
    if (!d || !b_local_dev || iters < 1 || !ms_per_kernel) return fail(SMLE_ERR_ARG, "bad argument");
    if (!d->connected) return fail(SMLE_ERR_COMM, "smle_dist_connect has not been called");
    smle_csr_t a = d->a;
    int rc = ensure_workspace(a, 1, 0);
    if (!rc) rc = ensure_scratch(a, 1);
    if (rc) return rc;
    CgWorkspace &w = a->ws;
    CgScalars cg = make_scalars(w, 1);
    CgVecArgs va;
    va.B = b_local_dev; va.X = w.Xd; va.R = w.R; va.P = dist_p(d); va.AP = w.AP;
    va.n = d->n_local; va.k = 1; va.part = w.part; va.ticket = a->ticket + 1;
    rc = launch_merge<double, true>(a, va.P, va.AP, 1, cg, /*dry=*/true);
    if (!rc) rc = launch_vec(0, va, cg, iters + 2, -1.0, d->seq_base);
    if (rc) return rc;
    dist_post_kernel<<<1, 32, 0, g_stream>>>(d->ctl, 2, cg.rs_old, cg.ctrl);
    cg1d_update_xp_kernel<<<1, kThreads, 0, g_stream>>>(va, cg, d->ctl, 1);
    dist_halo_push_kernel<<<dist_push_grid(d), kThreads, 0, g_stream>>>(d->ctl, va.P, cg.ctrl);
    g_launches += 3;
    rc = check_launch("distributed CG init");
    if (rc) return rc;
    std::vector<cudaEvent_t> ev((size_t)iters * 4);
    for (auto &e : ev) CU(cudaEventCreate(&e));
    const int grid = dist_vec_grid(va.n);
    rc = dist_launch_iteration(d, va, cg);   // warm-up iteration, untimed
    for (int i = 0; i < iters && !rc; ++i) {
        CU(cudaEventRecord(ev[(size_t)i * 4 + 0], g_stream));
        g_spmv_dist = d->ctl_dev;
        rc = launch_merge<double, true>(a, va.P, va.AP, 1, cg);
        g_spmv_dist = nullptr;
        CU(cudaEventRecord(ev[(size_t)i * 4 + 1], g_stream));
        cg1d_update_r_kernel<<<grid, kThreads, 0, g_stream>>>(va, cg, d->ctl);
        CU(cudaEventRecord(ev[(size_t)i * 4 + 2], g_stream));
        cg1d_update_xp_kernel<<<grid, kThreads, 0, g_stream>>>(va, cg, d->ctl, 0);
        if (!d->ctl.fused) dist_halo_push_kernel<<<dist_push_grid(d), kThreads, 0, g_stream>>>(d->ctl, va.P, cg.ctrl);
        CU(cudaEventRecord(ev[(size_t)i * 4 + 3], g_stream));
        g_launches += d->ctl.fused ? 2 : 3;
        if (!rc) rc = check_launch("distributed CG iteration");
    }
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (!rc && e != cudaSuccess) rc = fail(SMLE_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e));
    double acc[3] = {0, 0, 0};
    if (!rc) {
        for (int i = 0; i < iters; ++i)
            for (int j = 0; j < 3; ++j) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)i * 4 + j], ev[(size_t)i * 4 + j + 1]);
                acc[j] += ms;
            }
        for (int j = 0; j < 3; ++j) ms_per_kernel[j] = (float)(acc[j] / iters);
    }
    for (auto &ev1 : ev) cudaEventDestroy(ev1);
    d->seq_base += iters + 4;   // iterations run (incl. the warm-up) + margin
    return rc;
}

// A/B aid: `iters` back-to-back all-reduces of one double through the peer-memory mailboxes, replayed
// from a CUDA graph of 64 (post + wait) pairs.  *us_each = microseconds per all-reduce on this rank.
// Collective.
int smle_dist_allreduce_bench_f64(smle_dist_t d, int iters, double *us_each)
{
    if (!d || iters < 64 || !us_each) return fail(SMLE_ERR_ARG, "bad argument (iters >= 64)");
    if (!d->connected) return fail(SMLE_ERR_COMM, "smle_dist_connect has not been called");
    int rc = ensure_workspace(d->a, 1, 0);
    if (rc) return rc;
    CgWorkspace &w = d->a->ws;
    CgScalars cg = make_scalars(w, 1);
    const int reps = iters / 64;
    int ctrl[CTRL_WORDS] = {0, 0, 0, reps * 64 + 64 + 1, 0, d->seq_base, 0, 0};
    CU(cudaMemcpyAsync(w.ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, g_stream));
    const double one = 1.0;
    CU(cudaMemcpyAsync(cg.rs_old, &one, sizeof(double), cudaMemcpyHostToDevice, g_stream));
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CU(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 64; ++i) {
        dist_post_kernel<<<1, 32, 0, g_stream>>>(d->ctl, 0, cg.rs_old, cg.ctrl);
        dist_allreduce_wait_kernel<<<1, 32, 0, g_stream>>>(d->ctl, cg.pAp, cg.ctrl);
    }
    cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
    if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(exec, g_stream);   // warm-up
    cudaEventRecord(e0, g_stream);
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(exec, g_stream);
    cudaEventRecord(e1, g_stream);
    e = cudaStreamSynchronize(g_stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaGraphExecDestroy(exec);
    g_launches += 2LL * 64 * (reps + 1);
    d->seq_base += reps * 64 + 64 + 2;
    if (e != cudaSuccess) return fail(SMLE_ERR_CUDA, "all-reduce bench failed: %s", cudaGetErrorString(e));
    double sum = 0.0;
    CU(cudaMemcpy(&sum, cg.pAp, sizeof(double), cudaMemcpyDeviceToHost));
    if (sum != (double)d->world) return fail(SMLE_ERR_COMM, "all-reduce bench: sum %g != world %d", sum, d->world);
    *us_each = 1e3 * ms / (reps * 64);
    return SMLE_OK;
}

} // extern "C"
