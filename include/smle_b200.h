/*
 * include/smle_b200.h -- C ABI of the B200-native merge-path SpMV / SpMM / CG library
 * (libsmle_b200.so, built from sparse-matrix-linear-equations_b200/csrc for sm_100a).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  The reference
 * (YuyaW-0118/Sparse-Matrix-Linear-Equations) has no FFI layer: the path sits behind C++
 * function templates called directly by its drivers.  Each entry point below names the
 * reference function (file:line under the reference tree) it replaces; the header-only C++
 * adapters in sparse-matrix-linear-equations_b200/host/smle_adapters.hpp keep those
 * functions' exact signatures on top of this ABI, and INTEGRATION.md shows the call-site
 * change a maintainer makes.
 *
 * Conventions
 *  - plain C types only; matrices are CSR triples exactly as CsrMatrix<ValueT,int> holds
 *    them (sparse_matrix.h:648-653): int row_offsets[m+1], int column_indices[nnz],
 *    ValueT values[nnz]; dense blocks are ROW-MAJOR n x k (merge_based.hpp:97-113).
 *  - every function returns 0 on success and a negative smle_status on failure; the text of
 *    the last failure on the calling thread is smle_last_error().  The library never calls
 *    exit() (the reference does: sparse_matrix.h:186-190) and never falls back to the CPU:
 *    without a usable CUDA device every compute entry point fails with SMLE_ERR_CUDA.
 *  - host pointers are borrowed for the duration of the call; a handle owns device copies.
 *  - `is_device_ptr` != 0 means the vector/block arguments already live in device memory of
 *    the current device (no copies are made); 0 means host memory (pinned or pageable).
 *  - one process drives one GPU (torch.distributed style); the library is not re-entrant
 *    across host threads, like the reference (hyper_parameters.hpp globals).
 */
#ifndef SMLE_B200_H
#define SMLE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smle_csr_s *smle_csr_t;

enum smle_status {
    SMLE_OK = 0,
    SMLE_ERR_ARG = -1,    /* bad argument (null pointer, negative size, k < 1, ...) */
    SMLE_ERR_CUDA = -2,   /* CUDA runtime error or no device */
    SMLE_ERR_ALLOC = -3,  /* host or device allocation failed */
    SMLE_ERR_RANGE = -4,  /* m + nnz does not fit the reference's 32-bit merge path */
    SMLE_ERR_COMM = -5    /* multi-GPU peer setup failed */
};

/* SpmmKernel of work_2025/types.hpp:11-16.  All three select the one merge-path kernel on
 * the GPU (the split strategy is a CPU threading choice); the value is accepted so that
 * CGSolveMultiple call sites keep compiling (no_pretreatment.hpp:94-105). */
enum smle_spmm_kernel { SMLE_SIMPLE = 0, SMLE_MERGE = 1, SMLE_NONZERO_SPLIT = 2 };

/* ---- library / device --------------------------------------------------------------- */
int smle_version(void);
const char *smle_last_error(void);
int smle_device_count(void);                 /* 0 when no CUDA device is usable */
int smle_init(int device);                   /* bind this process to `device`, create the stream */
void smle_shutdown(void);
int smle_set_stream(void *cuda_stream);      /* run on the caller's cudaStream_t (NULL: own stream) */
void *smle_get_stream(void);                 /* the cudaStream_t kernels are launched on */
int smle_sync(void);                         /* cudaStreamSynchronize on that stream */
long long smle_launch_count(void);           /* kernels launched by this library so far */
int smle_sm_count(void);

/* ---- device buffers ------------------------------------------------------------------------
 * For callers that keep vectors resident in HBM across calls (is_device_ptr = 1) without
 * linking the CUDA runtime themselves; the reference has no counterpart (single address space). */
int smle_malloc(void **dev_ptr, unsigned long long bytes);
int smle_free(void *dev_ptr);
int smle_copy_to_device(void *dev_dst, const void *host_src, unsigned long long bytes);
int smle_copy_to_host(void *host_dst, const void *dev_src, unsigned long long bytes);
/* Page-lock caller-owned host memory for the duration of a series of calls (cudaHostRegister): with
 * pageable buffers the runtime stages every copy and the batch solve below cannot overlap copies
 * with solves.  The adapters register the driver's b / x blocks once, outside the timing loop
 * (the reference allocates them with mkl_malloc, cpu_singlecg.cpp:80-86).  Returns 0 when the range
 * was registered, 1 when it already was page-locked (nothing to unregister), < 0 on failure. */
int smle_host_register(void *host_ptr, unsigned long long bytes);
int smle_host_unregister(void *host_ptr);

/* ---- merge-path partition -----------------------------------------------------------
 * Replaces MergePathSearch (work_2025/spmm/merge_based.hpp:22-44; copies at
 * cpu_spmv.cpp:213-235, cub/thread/thread_search.cuh:48-83) evaluated on the share
 * boundaries the reference threads use (merge_based.hpp:72-82): diagonal_t =
 * min(share*t, m+nnz), share = items_per_part > 0 ? items_per_part
 * : ceil((m+nnz)/num_parts).  row_end_offsets = row_offsets+1 (host pointer, m entries).
 * out_xy receives 2*(num_parts+1) ints: (rows consumed, nonzeros consumed) per boundary.
 * The search runs ON THE GPU (one thread per boundary) and must be bit-exact. */
int smle_merge_path_partition(const int *row_end_offsets, int m, int nnz, int num_parts,
                              int items_per_part, int *out_xy);

/* ---- CSR handle ------------------------------------------------------------------------
 * Device-resident copy of a CsrMatrix<ValueT,int> (sparse_matrix.h:633-653).  The handle
 * also caches the tile coordinates of the merge path (they depend on A only). */
int smle_csr_create_f64(smle_csr_t *out, int m, int n, int nnz, const int *row_offsets,
                        const int *column_indices, const double *values);
int smle_csr_create_f32(smle_csr_t *out, int m, int n, int nnz, const int *row_offsets,
                        const int *column_indices, const float *values);
void smle_csr_destroy(smle_csr_t a);
int smle_csr_dims(smle_csr_t a, int *m, int *n, int *nnz, int *value_bytes);
/* tile coordinates the SpMV/SpMM kernel for `k` right-hand sides uses: out_xy gets
 * 2*(num_tiles+1) ints (capacity in ints); either output may be NULL to query sizes. */
int smle_csr_tile_coords(smle_csr_t a, int k, int *num_tiles, int *items_per_tile,
                         int *out_xy, int capacity);

/* ---- SpMV / SpMM -----------------------------------------------------------------------
 * smle_spmv_*: y = A x.     Replaces OmpMergeCsrmv (cpu_spmv.cpp:360-421).
 * smle_spmm_*: Y = A X, X n x k and Y m x k row-major.
 *              Replaces OmpMergeCsrmm (work_2025/spmm/merge_based.hpp:49-153) and, behind
 *              the SpmmKernel switch, OmpNonzeroSplitCsrmm / OmpCsrSpmmT
 *              (nonzero_splitting.hpp:52-150, row_splitting.hpp:18-54). */
int smle_spmv_f64(smle_csr_t a, const double *x, double *y, int is_device_ptr);
int smle_spmv_f32(smle_csr_t a, const float *x, float *y, int is_device_ptr);
int smle_spmm_f64(smle_csr_t a, const double *X, double *Y, int k, int is_device_ptr);
int smle_spmm_f32(smle_csr_t a, const float *X, float *Y, int k, int is_device_ptr);

/* ---- conjugate gradient ------------------------------------------------------------------
 * smle_cg_single_f64 replaces CGSolveSingle (work_2025/main/single_strategy.hpp:105-170):
 *   x0 = 0, r = p = b; stop when sqrt(r.r)/||b|| < tol, tested before the p update;
 *   *iters_out counts SpMV applications (== max_iters when not converged); ||b|| == 0 -> 1.
 * smle_cg_multi_f64 replaces CGSolveMultiple (work_2025/main/no_pretreatment.hpp:35-197):
 *   k lock-step recurrences over row-major n x k blocks, per-column convergence latch
 *   (alpha = beta = 0 afterwards), stops when every column has latched.
 *   max_err_hist (nullable, hist_capacity doubles) receives the per-iteration maximum
 *   relative residual the reference pushes into `max_errors` (:133-155); *hist_len (nullable)
 *   the number of iterations recorded.  `kernel` is an smle_spmm_kernel (see above).
 * final_rel_res (nullable): sqrt(r.r)/||b|| at exit (max over columns for multi). */
int smle_cg_single_f64(smle_csr_t a, const double *b, double *x, int max_iters, double tol,
                       int is_device_ptr, int *iters_out, double *final_rel_res);
/* smle_cg_single_batch_f64 replaces the solve loop of TestCGSolveSingle
 * (work_2025/main/single_strategy.hpp:199-226): num_vectors systems, vector v = b_vectors[v*n ...]
 * and x_solutions[v*n ...] (host memory), solved one after another exactly like
 * smle_cg_single_f64.  The upload of b_{v+1} and the download of x_{v-1} run on a copy stream
 * while system v is being solved.  iters_each (nullable, num_vectors ints) receives the
 * per-vector counts, *iters_total (nullable) their sum -- what the reference reports. */
int smle_cg_single_batch_f64(smle_csr_t a, const double *b_vectors, double *x_solutions,
                             int num_vectors, int max_iters, double tol, int *iters_each,
                             long long *iters_total);
int smle_cg_multi_f64(smle_csr_t a, const double *B, double *X, int k, int max_iters,
                      double tol, int kernel, int is_device_ptr, int *iters_out,
                      double *max_err_hist, int hist_capacity, int *hist_len,
                      double *final_rel_res);

/* fp32 variants: CGSolveSingle / CGSolveMultiple as the reference's templates
 * instantiate them for <float,int> (the reference drivers only use <double,int>).  The handle must be
 * an fp32 one (smle_csr_create_f32); blocks are float, the scalars of the recurrence and all dot
 * products are accumulated in double; histories and residuals are reported as double. */
int smle_cg_single_f32(smle_csr_t a, const float *b, float *x, int max_iters, float tol,
                       int is_device_ptr, int *iters_out, double *final_rel_res);
int smle_cg_multi_f32(smle_csr_t a, const float *B, float *X, int k, int max_iters, float tol,
                      int kernel, int is_device_ptr, int *iters_out, double *max_err_hist,
                      int hist_capacity, int *hist_len, double *final_rel_res);

/* ---- SPAI-preconditioned multi-RHS CG -------------------------------
 * smle_spai_build_f64 replaces SparseApproximateInversion
 *   (work_2025/cg/sparse_approximate_inversion.hpp:41-321): HOST code, like the reference -- the
 *   values of M on A's pattern (static pattern S_M = S_A, per-column least squares, symmetrised);
 *   m_values[nnz] pairs with A's row_offsets / column_indices.
 * smle_pcg_spai_multi_f64 replaces SPAISolveMultiple
 *   (work_2025/main/sparse_approximate_inverse.hpp:31-230): `m` is a handle of M (same size as A);
 *   per iteration two sparse products (A p with the fused p.Ap, z = M r with the fused r.z) and two
 *   fused vector kernels; alpha is also 0 when p.Ap == 0 and beta when the old r.z == 0, the
 *   convergence test sqrt(r.r)/||b|| < tol comes before the M step, as in the reference.
 *   The reference's NONZERO_SPLIT kernel adds into the last row of its output instead of storing
 *   it (nonzero_splitting.hpp:137-149), so SPAISolveMultiple, which does not clear Z / AP between
 *   iterations, never converges with it; here all three `kernel` values compute Y = A X. */
int smle_spai_build_f64(int m, int nnz, const int *row_offsets, const int *column_indices, const double *values,
                        double *m_values);
int smle_pcg_spai_multi_f64(smle_csr_t a, smle_csr_t m, const double *B, double *X, int k, int max_iters, double tol,
                            int kernel, int is_device_ptr, int *iters_out, double *max_err_hist, int hist_capacity,
                            int *hist_len, double *final_rel_res);

/* Fixed-count variant for measurement: exactly `iters` CG iterations (no convergence exit),
 * device pointers only.  Same kernels and graph as smle_cg_multi_f64. */
int smle_cg_run_fixed_f64(smle_csr_t a, const double *B, double *X, int k, int iters);
/* Per-kernel timing for the roofline report: runs `iters` CG iterations WITHOUT the CUDA graph,
 * bracketing each of the three kernels of an iteration with CUDA events on the launch stream.
 * ms_per_kernel[3] = mean milliseconds of {SpMM+dot, r-update+dot, x/p-update}. Device pointers. */
int smle_cg_profile_f64(smle_csr_t a, const double *B, double *X, int k, int iters,
                        float *ms_per_kernel);

/* ---- row-partitioned CG over NVLink peer memory (one process per GPU) -----------------------
 * Net-new relative to the reference (which has no distributed code); semantics of the solve are
 * CGSolveSingle's (single_strategy.hpp:105-170) on the global system.  The partition is cut at
 * the reference's merge-path coordinates: part g owns rows [x_g, x_{g+1}) with
 * x_g = MergePathSearch(min(g*ceil((m+nnz)/G), m+nnz)).x (merge_based.hpp:22-44 on the share
 * diagonals :72-82); a row cut mid-way belongs whole to the later part.
 *
 * smle_dist_bounds   bounds[world+1] from the GLOBAL row offsets (m+1 host ints; the search runs on
 *                    the GPU like smle_merge_path_partition and is bit-exact with the reference).
 *                    Only the row offsets are global -- no rank needs the whole matrix.
 *
 * Planner (host only, csrc/smle_plan.cpp).  A rank hands in ITS rows only:
 *   smle_dist_plan_create      rows [bounds[rank], bounds[rank+1]) as local row offsets (starting
 *                              at 0) + GLOBAL column indices.  Builds the halo index map (sorted
 *                              unique out-of-range columns) and the local column indices
 *                              [0,n_local) own | pad | [halo_base, halo_base+n_halo) halo, halo_base
 *                              = n_local rounded up to a 128-byte line.
 *   smle_dist_plan_request     the int blob this rank publishes (size: _request_size ints).
 *   smle_dist_plan_finish      all ranks' blobs, concatenated in rank order (blob_off[world+1] int
 *                              offsets), turn into the push plan: which local rows go to which peer
 *                              and where they land in that peer's extended vector.  The caller moves
 *                              the blobs (torch.distributed all-gather, shared memory, MPI ...).
 *   smle_dist_plan_send        the push plan, for inspection: send_off[world+1], send_idx
 *                              [send_off[world]], send_dst[world], needs_from[world] (each nullable).
 *   smle_dist_create_from_plan uploads the local system (values = the rank's rows, in CSR order)
 *                              and allocates the communication buffer.
 * smle_dist_ipc_handle / smle_dist_connect exchange the 64-byte CUDA IPC handles of the per-rank
 * communication buffers (again moved by the caller; world == 1 passes NULL).
 * smle_dist_spmv_f64 and smle_dist_cg_f64 are collective over the partition; vectors hold this
 * rank's rows (smle_dist_spmv_f64: device pointers; smle_dist_cg_f64: device or host memory per
 * is_device_ptr).  Halo pushes and the dot-product all-reduce are done by the kernels themselves
 * through peer stores and mailbox flags -- no NCCL on the data path.  A peer that stays silent for
 * 4 s makes every rank return SMLE_ERR_COMM instead of hanging. */
typedef struct smle_dist_s *smle_dist_t;
typedef struct smle_plan_s *smle_plan_t;
int smle_dist_bounds(const int *row_offsets, int m, int world, int *bounds);
int smle_dist_plan_create(smle_plan_t *out, int rank, int world, const int *bounds, int num_cols_global,
                          const int *local_row_offsets, const int *global_column_indices);
int smle_dist_plan_dims(smle_plan_t p, int *n_local, int *n_halo, int *halo_base, int *nnz_local);
int smle_dist_plan_local_columns(smle_plan_t p, int *out);
int smle_dist_plan_halo_columns(smle_plan_t p, int *out);
long long smle_dist_plan_request_size(smle_plan_t p);
int smle_dist_plan_request(smle_plan_t p, int *blob);
int smle_dist_plan_finish(smle_plan_t p, const int *all_blobs, const long long *blob_off);
int smle_dist_plan_send(smle_plan_t p, int *send_off, int *send_idx, int *send_dst, int *needs_from);
void smle_dist_plan_destroy(smle_plan_t p);
int smle_dist_create_from_plan(smle_dist_t *out, smle_plan_t p, const double *local_values);
int smle_dist_ipc_handle(smle_dist_t d, unsigned char *out64);
int smle_dist_connect(smle_dist_t d, const unsigned char *all_handles);
int smle_dist_dims(smle_dist_t d, int *n_local, int *n_halo, int *rank, int *world);
int smle_dist_spmv_f64(smle_dist_t d, const double *x_local_dev, double *y_local_dev);
int smle_dist_cg_f64(smle_dist_t d, const double *b_local, double *x_local, int max_iters, double tol,
                     int is_device_ptr, int *iters_out, double *final_rel_res);
/* per-kernel timing of the row-partitioned iteration, like smle_cg_profile_f64 (collective) */
int smle_dist_cg_profile_f64(smle_dist_t d, const double *b_local_dev, int iters, float *ms_per_kernel);
/* measurement aid: microseconds per all-reduce of one double through the peer-memory mailboxes
 * (CUDA graph of post + wait kernel pairs; iters >= 64); collective */
int smle_dist_allreduce_bench_f64(smle_dist_t d, int iters, double *us_each);
void smle_dist_destroy(smle_dist_t d);

/* ---- matrix / RHS generators (host side) -------------------------------------------------
 * CSR output identical to the reference generator followed by CsrMatrix::Init
 * (sparse_matrix.h:668-733): InitGrid2d :458-527, InitGrid3d :533-623, InitWheel :417-450,
 * InitDense :385-412.  `diag`/`offd` are the tuple values for row==col / row!=col
 * (1.0/1.0 = reference default; 4/-1 and 6/-1 = the Poisson fill of SURVEY.md App. B).
 * Caller allocates row_offsets[m+1], column_indices[nnz], values[nnz] from *_shape. */
int smle_gen_grid2d_shape(int width, int self_loop, int *m, int *n, int *nnz);
int smle_gen_grid3d_shape(int width, int self_loop, int *m, int *n, int *nnz);
int smle_gen_wheel_shape(int spokes, int *m, int *n, int *nnz);
int smle_gen_dense_shape(int rows, int cols, int *m, int *n, int *nnz);
int smle_gen_rmat_shape(int scale, int edge_factor, int *m, int *n, int *nnz);
int smle_gen_grid2d_f64(int width, int self_loop, double diag, double offd, int *row_offsets,
                        int *column_indices, double *values);
int smle_gen_grid2d_f32(int width, int self_loop, float diag, float offd, int *row_offsets,
                        int *column_indices, float *values);
int smle_gen_grid3d_f64(int width, int self_loop, double diag, double offd, int *row_offsets,
                        int *column_indices, double *values);
int smle_gen_grid3d_f32(int width, int self_loop, float diag, float offd, int *row_offsets,
                        int *column_indices, float *values);
/* Slab-local generation for the row-partitioned solve: the global row offsets alone
 * (m+1 ints), and rows [r0, r1) of the same matrix -- local_row_offsets[r1-r0+1] starting at 0,
 * GLOBAL column indices and values, row_offsets[r1]-row_offsets[r0] of each. */
int smle_gen_grid3d_row_offsets(int width, int self_loop, int *row_offsets);
int smle_gen_grid3d_rows_f64(int width, int self_loop, double diag, double offd, int r0, int r1,
                             int *local_row_offsets, int *column_indices, double *values);
int smle_gen_wheel_f64(int spokes, double value, int *row_offsets, int *column_indices,
                       double *values);
int smle_gen_wheel_f32(int spokes, float value, int *row_offsets, int *column_indices,
                       float *values);
int smle_gen_dense_f64(int rows, int cols, double value, int *row_offsets, int *column_indices,
                       double *values);
int smle_gen_dense_f32(int rows, int cols, float value, int *row_offsets, int *column_indices,
                       float *values);
/* R-MAT stand-in for the SuiteSparse set (get_uf_datasets.sh needs a network): 2^scale rows,
 * edge_factor*2^scale edges, quadrant probabilities a,b,c,(1-a-b-c), duplicates kept, sorted
 * by (row, col) like CsrMatrix::Init; values uniform in (0,1] or all 1.0 (unit_values). */
int smle_gen_rmat_f64(int scale, int edge_factor, double a, double b, double c,
                      unsigned long long seed, int unit_values, int *row_offsets,
                      int *column_indices, double *values);
int smle_gen_rmat_f32(int scale, int edge_factor, double a, double b, double c,
                      unsigned long long seed, int unit_values, int *row_offsets,
                      int *column_indices, float *values);
/* RHS as the CG drivers build it: srand(seed); b[i] = rand()/RAND_MAX, i < count
 * (cpu_singlecg.cpp:88-90, cpu_multicg.cpp:164-166); glibc rand(). */
int smle_gen_rhs_rand_f64(unsigned seed, long long count, double *out);
/* entries [first, first+count) of the same stream (the generator is sequential: `first` values are
 * drawn and dropped), for a rank that holds only its rows of b */
int smle_gen_rhs_rand_range_f64(unsigned seed, long long first, long long count, double *out);
/* the drivers' tolerance quirk: ||b[0:n]||_2 * tol (cpu_singlecg.cpp:23-34, cpu_multicg.cpp:50-62) */
double smle_driver_threshold_f64(const double *b, int n, double tol);

#ifdef __cplusplus
}
#endif
#endif /* SMLE_B200_H */
