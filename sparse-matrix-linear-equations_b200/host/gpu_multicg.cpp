// gpu_multicg -- B200 counterpart of the reference's cpu_multicg driver (cpu_multicg.cpp:109-333)
// for its un-preconditioned part: multi-RHS CG on row-major n x k blocks.
// Flags: --mtx --threads(ignored) --num_vectors(16) --max_iters(50000) --tolerance(1e-5) --quiet,
// plus the generators (Poisson fill) and --seed (the reference seeds with time(NULL), :164; a
// fixed default keeps runs comparable).  Prints the reference's "Min time ... Iters ... GFLOPS/s"
// line (:203-204) and writes data/error_data/<name>_cg_errors.csv (:67-86).
// --sweep reproduces the missing cpu_multicg2 (Makefile:191, eval_gflops.sh:61-66): kernels
// {SIMPLE,MERGE,NONZERO_SPLIT} x num_vectors {2,...,128}, CSV
// matrix_name,kernel,num_vectors,min_ms,gflops,iterations (verification/gflops/gflop_analyze.py).
// --gpus=N shards the num_vectors columns over N GPUs (one forked worker per GPU, A replicated; the block's
// iteration count and error history follow from the shards' -- smle_multi.hpp).
#include "smle_adapters.hpp"
#include "smle_host.hpp"
#include "smle_multi.hpp"

using namespace smle_host;

int main(int argc, char **argv)
{
    Args args(argc, argv);
    int max_iters = 50000, k = 16, device = 0, timing_iters = -1, gpus = 1;
    unsigned seed = 42;
    double tolerance = 1.0e-5;
    std::string output_csv;
    args.get("max_iters", max_iters); args.get("tolerance", tolerance); args.get("num_vectors", k);
    args.get("device", device); args.get("seed", seed); args.get("timing_iters", timing_iters); args.get("output", output_csv);
    args.get("gpus", gpus);
    const bool quiet = args.flag("quiet");
    if (gpus < 1 || gpus > smle_multi::kMaxWorld) { fprintf(stderr, "--gpus must be 1..%d\n", smle_multi::kMaxWorld); return 1; }
    // with --gpus > 1 the workers bind the devices after the fork: no CUDA call in this process
    if (gpus == 1 && smle_init(device)) smle_adapters::die("smle_init");

    Csr<double> a;
    std::string label = matrix_from_args(args, a, true);
    if (label.empty()) { fprintf(stderr, "Please specify a matrix file with --mtx=<filename> (or --grid3d=<w>)\n"); return 1; }
    std::string name = base_name(label);
    printf("%s, ", label.c_str());
    print_stats_csv(a);
    printf("\n");
    const long long n = a.num_rows;
    const double flops_single = 2.0 * a.num_nonzeros + 10.0 * a.num_rows;   // cpu_multicg.cpp:176

    auto solve = [&](int kk, int kernel, int titers, double &min_ms, double &iters, std::vector<double> *errs) {
        std::vector<double> B((size_t)n * kk), X((size_t)n * kk);
        smle_gen_rhs_rand_f64(seed, n * kk, B.data());
        double threshold = args.flag("raw_tolerance") ? tolerance : smle_driver_threshold_f64(B.data(), (int)n, tolerance);   // :168
        if (!quiet) printf("Convergence threshold: %.6e\n", threshold);
        if (gpus == 1) {
            TestGpuCGMultipleRHS(a, B.data(), X.data(), max_iters, threshold, kk, titers, kernel, min_ms, iters, errs);
            return;
        }
        double *Xs = (double *)smle_multi::shared_alloc(sizeof(double) * (size_t)n * kk);
        smle_multi::Arena *ar = smle_multi::arena_create(gpus, 16);
        smle_multi::ColumnShards *cs = smle_multi::column_shards_create(gpus, max_iters);
        if (smle_multi::run_workers(ar, [&](int rank) {
                smle_multi::worker_columns(ar, cs, rank, a, B.data(), Xs, kk, max_iters, threshold, titers, kernel);
            })) { fprintf(stderr, "a worker failed\n"); exit(1); }
        iters = smle_multi::merge_column_shards(cs, gpus, errs);
        min_ms = cs->min_ms;
        if (args.flag("check")) {   // true residual of column 0 on the host (reporting aid)
            double rr = 0, bb = 0;
            for (int r = 0; r < a.num_rows; ++r) {
                double s = 0;
                for (int z = a.row_offsets[r]; z < a.row_offsets[r + 1]; ++z) s += a.values[z] * Xs[(size_t)a.column_indices[z] * kk];
                rr += (B[(size_t)r * kk] - s) * (B[(size_t)r * kk] - s); bb += B[(size_t)r * kk] * B[(size_t)r * kk];
            }
            printf("  true residual of column 0: %.3e\n", sqrt(rr / bb));
        }
    };

    if (args.flag("sweep")) {
        const char *kname[3] = {"SIMPLE", "MERGE", "NONZERO_SPLIT"};
        if (output_csv.empty()) output_csv = "data/gflops/" + name + "_gflops.csv";
        FILE *f = fopen(output_csv.c_str(), "w");
        if (f) fprintf(f, "matrix_name,kernel,num_vectors,min_ms,gflops,iterations\n");
        for (int kernel = 0; kernel < 3; ++kernel)
            for (int kk : {2, 4, 8, 16, 32, 64, 128}) {
                double min_ms, iters;
                solve(kk, kernel, timing_iters > 0 ? timing_iters : 1, min_ms, iters, nullptr);
                double gflops = flops_single * kk * iters / (min_ms / 1000.0) / 1e9;
                printf("    %s, L=%d, method=%s: %.3f ms, %d iters, %.2f GFLOPS\n", name.c_str(), kk, kname[kernel], min_ms, (int)iters, gflops);
                if (f) fprintf(f, "%s,%s,%d,%.3f,%.2f,%d\n", name.c_str(), kname[kernel], kk, min_ms, gflops, (int)iters);
            }
        if (f) { fclose(f); printf("Results saved to: %s\n", output_csv.c_str()); }
        return 0;
    }

    int titers = timing_iters > 0 ? timing_iters
                                  : (int)std::min(100ull, std::max(3ull, (16ull << 30) / ((unsigned long long)a.num_nonzeros * k)));   // :150
    std::vector<double> errs;
    double min_ms, iters;
    if (!quiet) printf("\n--- 2. CG (Multiple-RHS w/ merge-path SpMM on B200) ---\n");
    solve(k, GPU_NONZERO_SPLIT, titers, min_ms, iters, &errs);   // the reference passes NONZERO_SPLIT (:202)
    double gflops = flops_single * k * iters / (min_ms / 1000.0) / 1e9;
    printf("Min time: %8.3f ms, Iters: %6.1f, Overall GFLOPS/s: %6.2f\n", min_ms, iters, gflops);
    std::string path = "data/error_data/" + name + "_cg_errors.csv";
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { fprintf(stderr, "Error: Cannot open file %s for writing\n", path.c_str()); }
    else {
        fprintf(f, "iteration,max_error\n");
        for (size_t i = 0; i < errs.size(); ++i) fprintf(f, "%zu,%e\n", i, errs[i]);
        fclose(f);
        printf("Saved CG error history to %s (%zu iterations)\n", path.c_str(), errs.size());
    }
    if (!quiet) printf("All tests completed.\n");
    return 0;
}
