// gpu_singlecg -- B200 counterpart of the reference's cpu_singlecg driver (cpu_singlecg.cpp).
// Flags: --mtx --output --threads(ignored) --max_iters(10000) --tolerance(1e-5) --timing_iters(1)
// --quiet, plus the generators (--grid2d/--grid3d with the Poisson fill, since the reference's
// value-1.0 grids are indefinite) and --num_vectors (default 16 = num_vectors_list,
// cpu_singlecg.cpp:160).  RHS: srand(42) stream (:88-90); the tolerance handed to the solver is
// the driver's threshold ||b[0:n]|| * tol (:92,101) unless --raw_tolerance is given.
// CSV: matrix_name,kernel,num_vectors,min_ms,gflops,iterations (:199).
// --gpus=N (SURVEY.md section 5's flag): the same pass over the L vectors with the system
// ROW-PARTITIONED over N GPUs (one worker process per GPU, forked here; smle_multi.hpp), rows cut at
// the reference's merge-path coordinates, halo and dot products GPU to GPU.
#include "smle_adapters.hpp"
#include "smle_host.hpp"
#include "smle_multi.hpp"

using namespace smle_host;

int main(int argc, char **argv)
{
    Args args(argc, argv);
    std::string output_csv;
    int max_iters = 10000, timing_iters = 1, device = 0, L = 16, gpus = 1;
    double tolerance = 1.0e-5;
    args.get("output", output_csv); args.get("max_iters", max_iters); args.get("tolerance", tolerance);
    args.get("timing_iters", timing_iters); args.get("device", device); args.get("num_vectors", L);
    args.get("gpus", gpus);
    const bool quiet = args.flag("quiet");
    if (gpus < 1 || gpus > smle_multi::kMaxWorld) { fprintf(stderr, "--gpus must be 1..%d\n", smle_multi::kMaxWorld); return 1; }
    // with --gpus the workers bind the devices after the fork: no CUDA call in this process before it
    const bool partitioned = gpus > 1 || args.flag("partitioned");
    if (!partitioned && smle_init(device)) smle_adapters::die("smle_init");

    Csr<double> a;
    std::string label = matrix_from_args(args, a, true);
    if (label.empty()) {
        fprintf(stderr, "Usage: %s --mtx=<filename> | --grid3d=<w> | --grid2d=<w> [options]\n", argv[0]);
        return 1;
    }
    std::string name = base_name(label);
    printf("Matrix: %s\n  Rows: %d, Cols: %d, NNZ: %d\n", name.c_str(), a.num_rows, a.num_cols, a.num_nonzeros);

    const long long n = a.num_rows;
    std::vector<double> b((size_t)n * L), x_own;
    smle_gen_rhs_rand_f64(42, n * L, b.data());
    double threshold = args.flag("raw_tolerance") ? tolerance : smle_driver_threshold_f64(b.data(), (int)n, tolerance);

    double min_ms = 0, total_iters = 0;
    if (!partitioned) {
        x_own.resize((size_t)n * L);
        TestGpuCGSolveSingle(a, b.data(), x_own.data(), max_iters, threshold, L, timing_iters, min_ms, total_iters);
    } else {
        if (a.num_rows != a.num_cols) { fprintf(stderr, "the row-partitioned solve needs a square matrix\n"); return 1; }
        // (no CUDA call here: a process that touched the driver cannot hand it to forked children)
        double *x = (double *)smle_multi::shared_alloc(sizeof(double) * (size_t)n * L);   // every rank writes its rows
        smle_multi::Arena *ar = smle_multi::arena_create(gpus, (long long)n + gpus + 16);
        int bad = smle_multi::run_workers(ar, [&](int rank) {
            smle_multi::worker(ar, rank, a, b.data(), x, L, max_iters, threshold, timing_iters);
        });
        if (bad) { fprintf(stderr, "a worker failed\n"); return 1; }
        min_ms = ar->min_ms; total_iters = ar->iters_of_min_ms;
        if (!quiet) {
            printf("  row partition over %d GPU(s):", gpus);
            for (int g = 0; g < gpus; ++g) printf(" [%d,%d) halo %d", ar->bounds[g], ar->bounds[g + 1], ar->n_halo[g]);
            printf("\n");
        }
        if (args.flag("check")) {   // true residual of vector 0 on the host (reporting aid)
            double rr = 0, bb = 0;
            for (int r = 0; r < a.num_rows; ++r) {
                double s = 0;
                for (int z = a.row_offsets[r]; z < a.row_offsets[r + 1]; ++z) s += a.values[z] * x[a.column_indices[z]];
                rr += (b[r] - s) * (b[r] - s); bb += b[r] * b[r];
            }
            printf("  true residual of vector 0: %.3e\n", sqrt(rr / bb));
        }
    }
    double gflops = (2.0 * a.num_nonzeros + 10.0 * a.num_rows) * total_iters / (min_ms / 1000.0) / 1e9;   // :94,108
    printf("    %s, L=%d, method=SINGLE_LOOP: %.3f ms, %lld iters, %.2f GFLOPS\n", name.c_str(), L, min_ms,
           (long long)total_iters, gflops);

    if (output_csv.empty()) output_csv = "data/simple_gflops/" + name + "_gflops.csv";
    FILE *f = fopen(output_csv.c_str(), "w");
    if (!f) { fprintf(stderr, "Error: Cannot open file %s for writing\n", output_csv.c_str()); return 0; }
    fprintf(f, "matrix_name,kernel,num_vectors,min_ms,gflops,iterations\n%s,SINGLE_LOOP,%d,%.3f,%.2f,%lld\n",
            name.c_str(), L, min_ms, gflops, (long long)total_iters);
    fclose(f);
    printf("Results saved to: %s\n", output_csv.c_str());
    if (!quiet) printf("All simple benchmarks completed.\n");
    return 0;
}
