// gpu_spmv -- B200 counterpart of the reference's cpu_spmv driver (cpu_spmv.cpp:750-991).
// Same flags (--quiet --v --fp32 --i= --mtx= --grid2d= --grid3d= --dense= --threads=) plus
// --wheel= (documented but never parsed by the reference, cpu_spmv.cpp:949,969-976), --rmat=,
// --self_loop, --poisson, --device=.  In --quiet mode it prints the CSV fragment eval_csrmv.sh
// expects: file, stats..., method_name, setup_ms, avg_spmv_ms, gflops, effective_GBs.
#include <chrono>
#include "smle_adapters.hpp"
#include "smle_host.hpp"

using namespace smle_host;

template <typename V>
int run(const Args &args, bool quiet)
{
    Csr<V> a;
    std::string label = matrix_from_args(args, a, false);
    if (label.empty()) return 1;
    if (a.num_rows == 1 || a.num_cols == 1 || a.num_nonzeros == 1) { if (!quiet) printf("Trivial dataset\n"); return 0; }
    printf("%s, ", label.c_str());
    if (quiet) print_stats_csv(a);
    else printf("\n\t num_rows: %d\n\t num_cols: %d\n\t num_nonzeros: %d\n", a.num_rows, a.num_cols, a.num_nonzeros);

    long long iters = -1;
    args.get("i", iters);
    if (iters < 0) iters = std::min(200000ll, std::max(100ll, (16ll << 30) / std::max(1, a.num_nonzeros)));

    std::vector<V> x((size_t)a.num_cols, V(0.0019)), y((size_t)a.num_rows), gold((size_t)a.num_rows);   // cpu_spmv.cpp:855
    std::vector<double> scale((size_t)a.num_rows);
    for (int r = 0; r < a.num_rows; ++r) {   // SpmvGold (cpu_spmv.cpp:245-265), alpha = 1, beta = 0
        V s = 0;
        double mag = 0;
        for (int z = a.row_offsets[r]; z < a.row_offsets[r + 1]; ++z) {
            s += a.values[z] * x[a.column_indices[z]];
            mag += fabs((double)a.values[z] * (double)x[a.column_indices[z]]);
        }
        gold[r] = s;
        scale[r] = mag;   // (|A||x|)_r: what a different summation order is measured against -- on a Poisson
    }                     // matrix with constant x the interior rows cancel to exactly 0 in the gold order only

    auto t0 = std::chrono::steady_clock::now();
    smle_csr_t h = smle_adapters::handle_of(a);
    double setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    // correctness through the reference-signature adapter (host buffers)
    GpuMergeCsrmv<V, int>(0, a, a.row_offsets + 1, a.column_indices, a.values, x.data(), y.data());
    double worst = 0;
    for (int r = 0; r < a.num_rows; ++r) {
        double d = fabs((double)y[r] - (double)gold[r]) / std::max(1e-300, scale[r]);
        worst = std::max(worst, d);
    }
    if (!quiet) printf("\tMerge CsrMV (B200): max row-scaled difference vs SpmvGold %.3e  %s\n", worst,
                       worst < (sizeof(V) == 8 ? 1e-12 : 1e-5) ? "PASS" : "FAIL");

    // timing with device-resident vectors: `iters` warm runs, `iters` timed runs (cpu_spmv.cpp:458-474)
    void *dx = nullptr, *dy = nullptr;
    if (smle_malloc(&dx, sizeof(V) * x.size()) || smle_malloc(&dy, sizeof(V) * y.size()) ||
        smle_copy_to_device(dx, x.data(), sizeof(V) * x.size())) smle_adapters::die("device buffers");
    auto spmv = [&]() {
        int rc;
        if constexpr (sizeof(V) == 8) rc = smle_spmv_f64(h, (const double *)dx, (double *)dy, 1);
        else rc = smle_spmv_f32(h, (const float *)dx, (float *)dy, 1);
        if (rc) smle_adapters::die("smle_spmv");
    };
    for (long long i = 0; i < iters; ++i) spmv();
    smle_sync();
    t0 = std::chrono::steady_clock::now();
    for (long long i = 0; i < iters; ++i) spmv();
    smle_sync();
    double avg_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    smle_free(dx); smle_free(dy);

    // DisplayPerf (cpu_spmv.cpp:716-742)
    double total_bytes = (double)a.num_nonzeros * (sizeof(V) * 2 + sizeof(int)) + (double)a.num_rows * (sizeof(int) + sizeof(V));
    double gflops = 2.0 * a.num_nonzeros / avg_ms / 1.0e6, gbs = total_bytes / avg_ms / 1.0e6;
    if (!quiet) printf("Merge CsrMV (B200), fp%d: %.4f setup ms, %.4f avg ms, %.5f gflops, %.3lf effective GB/s\n",
                       int(sizeof(V) * 8), setup_ms, avg_ms, gflops, gbs);
    else printf("Merge CsrMV (B200), %.5f, %.5f, %.6f, %.3lf, ", setup_ms, avg_ms, gflops, gbs);
    printf("\n");
    return 0;
}

int main(int argc, char **argv)
{
    Args args(argc, argv);
    if (args.flag("help")) {
        printf("%s [--quiet] [--v] [--i=<timing iterations>] [--fp64 (default) | --fp32] [--device=<gpu>]\n"
               "\t--mtx=<matrix market file> | --dense=<cols> | --grid2d=<width> | --grid3d=<width> | --wheel=<spokes> | --rmat=<scale>\n"
               "\t[--self_loop] [--poisson]\n", argv[0]);
        return 0;
    }
    int device = 0;
    args.get("device", device);
    if (smle_init(device)) smle_adapters::die("smle_init");
    const bool quiet = args.flag("quiet");
    return args.flag("fp32") ? run<float>(args, quiet) : run<double>(args, quiet);
}
