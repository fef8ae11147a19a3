// smle_multi.hpp -- one process per GPU from a single C++ driver (gpu_singlecg --gpus=N).
//
// The library binds one process to one GPU (like torch.distributed), so a multi-GPU driver forks N
// workers BEFORE the first CUDA call.  The matrix and the right-hand sides were loaded by the
// parent: the workers inherit those pages copy-on-write, i.e. the host holds ONE copy of the global
// system and every worker reads only its own rows of it.  What the ranks must exchange -- the
// planner's request blobs and the CUDA IPC handles (include/smle_b200.h, smle_dist_plan_*) -- goes
// through an anonymous shared mapping with a process-shared barrier; solutions are written by every
// rank straight into its rows of a shared x block.  Nothing here is on the data path: halo and dot
// products travel GPU to GPU inside the kernels.
#pragma once
#include <omp.h>
#include <pthread.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/smle_b200.h"

namespace smle_multi {

constexpr int kMaxWorld = 8;

inline void *shared_alloc(size_t bytes)
{
    void *p = mmap(nullptr, bytes ? bytes : 1, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("mmap"); exit(1); }
    return p;
}

struct Arena {
    pthread_barrier_t barrier;
    int world;
    long long blob_cap;                 // ints per rank in `blobs`
    long long blob_len[kMaxWorld];
    unsigned char ipc[kMaxWorld][64];
    int failed;                         // a rank reported an error
    // results (written by rank 0)
    double min_ms, iters_of_min_ms;
    int bounds[kMaxWorld + 1];
    int n_halo[kMaxWorld];
    int *blobs;                         // world * blob_cap ints (separate lazily committed mapping)
};

inline Arena *arena_create(int world, long long max_blob_ints)
{
    Arena *a = (Arena *)shared_alloc(sizeof(Arena));
    memset(a, 0, sizeof(Arena));
    a->world = world;
    a->blob_cap = max_blob_ints;
    a->blobs = (int *)shared_alloc(sizeof(int) * (size_t)world * (size_t)max_blob_ints);
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&a->barrier, &at, (unsigned)world);
    pthread_barrierattr_destroy(&at);
    return a;
}

inline void sync(Arena *a) { pthread_barrier_wait(&a->barrier); }

#define SMLE_MULTI_CHECK(call)                                                                   \
    do {                                                                                         \
        if ((call) != 0) {                                                                       \
            fprintf(stderr, "[rank %d] %s failed: %s\n", rank, #call, smle_last_error());        \
            ar->failed = 1;                                                                      \
            _exit(1);                                                                            \
        }                                                                                        \
    } while (0)

// This rank's share of TestCGSolveSingle (single_strategy.hpp:179-240) on the row-partitioned system:
// L vectors (vector v = b[v*n ...], column-major) solved one after another, min wall time over
// timing_iterations, iterations summed over the vectors.  a_* is the GLOBAL CSR (inherited pages);
// x must be a shared mapping.  Runs in the forked worker; never returns to the caller's main.
template <typename CsrT>
[[noreturn]] inline void worker(Arena *ar, int rank, const CsrT &a, const double *b, double *x, int L, int max_iters,
                                double tol, int timing_iterations)
{
    const int world = ar->world, m = a.num_rows;
    // The parent generated the matrix with OpenMP: its worker threads do not exist in this forked child,
    // and a parallel region that tried to wake them would wait forever (libgomp).  Teams of one thread
    // never touch the inherited pool; the host work left here (planner remap) is small.
    omp_set_num_threads(1);
    SMLE_MULTI_CHECK(smle_init(rank));
    std::vector<int> bounds((size_t)world + 1);
    SMLE_MULTI_CHECK(smle_dist_bounds(a.row_offsets, m, world, bounds.data()));   // merge-path search on the GPU
    const int r0 = bounds[rank], r1 = bounds[rank + 1], lo = a.row_offsets[r0];
    std::vector<int> lro((size_t)(r1 - r0) + 1);
    for (int i = 0; i <= r1 - r0; ++i) lro[(size_t)i] = a.row_offsets[r0 + i] - lo;
    smle_plan_t plan = nullptr;
    SMLE_MULTI_CHECK(smle_dist_plan_create(&plan, rank, world, bounds.data(), a.num_cols, lro.data(), a.column_indices + lo));
    const long long len = smle_dist_plan_request_size(plan);
    if (len > ar->blob_cap) { fprintf(stderr, "[rank %d] request blob too large\n", rank); ar->failed = 1; _exit(1); }
    SMLE_MULTI_CHECK(smle_dist_plan_request(plan, ar->blobs + (size_t)rank * (size_t)ar->blob_cap));
    ar->blob_len[rank] = len;
    sync(ar);
    {   // all blobs, concatenated in rank order
        std::vector<long long> off((size_t)world + 1, 0);
        for (int q = 0; q < world; ++q) off[(size_t)q + 1] = off[(size_t)q] + ar->blob_len[q];
        std::vector<int> all((size_t)off[(size_t)world]);
        for (int q = 0; q < world; ++q)
            memcpy(all.data() + off[(size_t)q], ar->blobs + (size_t)q * (size_t)ar->blob_cap, sizeof(int) * (size_t)ar->blob_len[q]);
        SMLE_MULTI_CHECK(smle_dist_plan_finish(plan, all.data(), off.data()));
    }
    smle_dist_t d = nullptr;
    SMLE_MULTI_CHECK(smle_dist_create_from_plan(&d, plan, a.values + lo));
    int n_halo = 0;
    smle_dist_plan_dims(plan, nullptr, &n_halo, nullptr, nullptr);
    smle_dist_plan_destroy(plan);
    SMLE_MULTI_CHECK(smle_dist_ipc_handle(d, ar->ipc[rank]));
    ar->n_halo[rank] = n_halo;
    if (rank == 0) memcpy(ar->bounds, bounds.data(), sizeof(int) * ((size_t)world + 1));
    sync(ar);
    SMLE_MULTI_CHECK(smle_dist_connect(d, &ar->ipc[0][0]));
    // this rank's rows of b and x stay page-locked for the copies inside the solves
    const size_t n = (size_t)m;
    for (int v = 0; v < L; ++v) {
        smle_host_register((void *)(b + (size_t)v * n + r0), sizeof(double) * (size_t)(r1 - r0));
        smle_host_register((void *)(x + (size_t)v * n + r0), sizeof(double) * (size_t)(r1 - r0));
    }
    sync(ar);
    double min_ms = 1e300, iters_min = 0;
    for (int t = 0; t < timing_iterations; ++t) {
        sync(ar);
        auto t0 = std::chrono::steady_clock::now();
        long long total = 0;
        for (int v = 0; v < L; ++v) {
            int it = 0;
            SMLE_MULTI_CHECK(smle_dist_cg_f64(d, b + (size_t)v * n + r0, x + (size_t)v * n + r0, max_iters, tol, 0, &it, nullptr));
            total += it;
        }
        sync(ar);   // the slowest rank closes the timing pass
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < min_ms) { min_ms = ms; iters_min = (double)total; }
    }
    if (rank == 0) { ar->min_ms = min_ms; ar->iters_of_min_ms = iters_min; }
    sync(ar);
    smle_dist_destroy(d);
    fflush(stdout);
    _exit(0);
}

// ---- column-sharded multi-RHS CG (first row) --------------------------------------
// The k recurrences of CGSolveMultiple never interact: rank g solves columns [lo, hi) of the row-major
// n x k blocks with A replicated on its GPU.  What the reference returns for the whole block follows
// from the shards: it iterates until EVERY column has latched (the slowest shard's count), and a shard
// that has stopped is frozen, so it keeps contributing its final residual to the per-iteration maximum.
struct ColumnShards {
    int iters[kMaxWorld];
    int hist_len[kMaxWorld];
    double min_ms;
    double *hist;        // world x hist_cap (shared mapping)
    int hist_cap;
};

inline void shard_columns(int k, int rank, int world, int *lo, int *hi)
{
    const int base = k / world, extra = k % world;
    *lo = rank * base + (rank < extra ? rank : extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
}

inline ColumnShards *column_shards_create(int world, int max_iters)
{
    ColumnShards *c = (ColumnShards *)shared_alloc(sizeof(ColumnShards));
    memset(c, 0, sizeof(ColumnShards));
    c->hist_cap = max_iters > 0 ? max_iters : 1;
    c->hist = (double *)shared_alloc(sizeof(double) * (size_t)world * (size_t)c->hist_cap);
    return c;
}

// -> iterations of the whole block; hist (optional) receives the per-iteration maximum over all columns
inline int merge_column_shards(const ColumnShards *c, int world, std::vector<double> *hist)
{
    int iters = 0;
    for (int r = 0; r < world; ++r) iters = c->iters[r] > iters ? c->iters[r] : iters;
    if (hist) {
        hist->assign((size_t)iters, 0.0);
        for (int r = 0; r < world; ++r) {
            const double *h = c->hist + (size_t)r * (size_t)c->hist_cap;
            const int len = c->hist_len[r];
            if (len <= 0) continue;
            for (int i = 0; i < iters; ++i) {
                const double v = h[i < len ? i : len - 1];
                if (v > (*hist)[(size_t)i]) (*hist)[(size_t)i] = v;
            }
        }
    }
    return iters;
}

// This rank's share of TestCGMultipleRHS (no_pretreatment.hpp:205-256): timing_iterations warm-up solves,
// then timed ones (min wall time, closed by the slowest rank); B and X are the WHOLE row-major n x k blocks
// (B inherited from the parent, X a shared mapping), of which the rank packs / fills its columns.
template <typename CsrT>
[[noreturn]] inline void worker_columns(Arena *ar, ColumnShards *cs, int rank, const CsrT &a, const double *B, double *X, int k,
                                        int max_iters, double tol, int timing_iterations, int kernel_type)
{
    const int world = ar->world;
    omp_set_num_threads(1);   // see worker(): no OpenMP teams in a forked child
    SMLE_MULTI_CHECK(smle_init(rank));
    int lo, hi;
    shard_columns(k, rank, world, &lo, &hi);
    const int kl = hi - lo;
    const size_t n = (size_t)a.num_rows;
    cs->iters[rank] = 0; cs->hist_len[rank] = 0;
    if (kl > 0) {
        smle_csr_t h = nullptr;
        SMLE_MULTI_CHECK(smle_csr_create_f64(&h, a.num_rows, a.num_cols, a.num_nonzeros, a.row_offsets, a.column_indices, a.values));
        std::vector<double> Bl(n * (size_t)kl), Xl(n * (size_t)kl);
        for (size_t r = 0; r < n; ++r)
            for (int c = 0; c < kl; ++c) Bl[r * (size_t)kl + c] = B[r * (size_t)k + lo + c];
        double *hist = cs->hist + (size_t)rank * (size_t)cs->hist_cap;
        double min_ms = 1e300;
        for (int t = 0; t < 2 * timing_iterations; ++t) {   // warm-ups, then timed solves
            sync(ar);
            auto t0 = std::chrono::steady_clock::now();
            int it = 0, hl = 0;
            SMLE_MULTI_CHECK(smle_cg_multi_f64(h, Bl.data(), Xl.data(), kl, max_iters, tol, kernel_type, 0, &it, hist, cs->hist_cap, &hl, nullptr));
            sync(ar);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (t >= timing_iterations && ms < min_ms) min_ms = ms;
            cs->iters[rank] = it; cs->hist_len[rank] = hl;
        }
        if (rank == 0) cs->min_ms = min_ms;
        for (size_t r = 0; r < n; ++r)
            for (int c = 0; c < kl; ++c) X[r * (size_t)k + lo + c] = Xl[r * (size_t)kl + c];
        smle_csr_destroy(h);
    } else {
        for (int t = 0; t < 2 * timing_iterations; ++t) { sync(ar); sync(ar); }
    }
    sync(ar);
    fflush(stdout);
    _exit(0);
}

// Parent side: fork `world` workers running fn(rank), wait for them; if one fails the others are
// killed (they may be waiting for it at the barrier).  Returns 0 when every worker exited 0.
template <typename Fn>
inline int run_workers(Arena *ar, Fn fn)
{
    const int world = ar->world;
    std::vector<pid_t> pids;
    fflush(stdout);
    fflush(stderr);
    for (int r = 0; r < world; ++r) {
        pid_t p = fork();
        if (p < 0) { perror("fork"); for (pid_t q : pids) kill(q, SIGKILL); return 1; }
        if (p == 0) { fn(r); _exit(0); }
        pids.push_back(p);
    }
    int bad = 0, left = world;
    while (left > 0) {
        int st = 0;
        pid_t p = waitpid(-1, &st, 0);
        if (p < 0) break;
        --left;
        if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
            if (!bad) for (pid_t q : pids) if (q != p) kill(q, SIGTERM);
            bad = 1;
        }
    }
    return bad || ar->failed;
}

} // namespace smle_multi
