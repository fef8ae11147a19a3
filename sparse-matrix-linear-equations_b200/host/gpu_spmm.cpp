// gpu_spmm -- B200 counterpart of the reference's cpu_spmm driver (cpu_spmm_v2.cpp:901-1183).
// Same flags plus the generators; --num_vectors defaults to 32 (cpu_spmm_v2.cpp:1154,1168).
// X is filled with 10.0 as in the reference (:1034-1042); the result is checked against a
// serial row-major SpMM.  CSV fragment: method, setup_ms, avg_ms, gflops, effective GB/s.
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
    int k = 32;
    args.get("num_vectors", k);
    if (k < 1) { fprintf(stderr, "--num_vectors must be >= 1\n"); return 1; }
    printf("%s, ", label.c_str());
    if (quiet) print_stats_csv(a);
    else printf("\n\t num_rows: %d\n\t num_cols: %d\n\t num_nonzeros: %d\n\t num_vectors: %d\n", a.num_rows, a.num_cols, a.num_nonzeros, k);

    long long iters = -1;
    args.get("i", iters);
    if (iters < 0) iters = std::min(1000ll, std::max(10ll, (16ll << 30) / std::max(1ll, (long long)a.num_nonzeros * k)));

    std::vector<V> X((size_t)a.num_cols * k, V(10.0)), Y((size_t)a.num_rows * k);
    if (args.flag("random_x")) { srand(42); for (auto &v : X) v = (V)rand() / (V)RAND_MAX; }

    auto t0 = std::chrono::steady_clock::now();
    smle_csr_t h = smle_adapters::handle_of(a);
    double setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    GpuMergeCsrmm<V, int>(0, a, a.row_offsets + 1, a.column_indices, a.values, X.data(), Y.data(), k);
    double worst = 0;
    std::vector<double> acc((size_t)k);
    for (int r = 0; r < a.num_rows; ++r) {
        std::fill(acc.begin(), acc.end(), 0.0);
        double scale = 0;
        for (int z = a.row_offsets[r]; z < a.row_offsets[r + 1]; ++z)
            for (int c = 0; c < k; ++c) {
                double t = (double)a.values[z] * (double)X[(size_t)a.column_indices[z] * k + c];
                acc[c] += t; if (c == 0) scale += fabs(t);
            }
        for (int c = 0; c < k; ++c) worst = std::max(worst, fabs(acc[c] - (double)Y[(size_t)r * k + c]) / std::max(1e-300, scale));
    }
    if (!quiet) printf("\tMerge CsrMM (B200): max row-scaled difference vs serial SpMM %.3e  %s\n", worst,
                       worst < (sizeof(V) == 8 ? 1e-12 : 1e-5) ? "PASS" : "FAIL");

    void *dX = nullptr, *dY = nullptr;
    if (smle_malloc(&dX, sizeof(V) * X.size()) || smle_malloc(&dY, sizeof(V) * Y.size()) ||
        smle_copy_to_device(dX, X.data(), sizeof(V) * X.size())) smle_adapters::die("device buffers");
    auto spmm = [&]() {
        int rc;
        if constexpr (sizeof(V) == 8) rc = smle_spmm_f64(h, (const double *)dX, (double *)dY, k, 1);
        else rc = smle_spmm_f32(h, (const float *)dX, (float *)dY, k, 1);
        if (rc) smle_adapters::die("smle_spmm");
    };
    for (long long i = 0; i < std::min(iters, 10ll); ++i) spmm();
    smle_sync();
    t0 = std::chrono::steady_clock::now();
    for (long long i = 0; i < iters; ++i) spmm();
    smle_sync();
    double avg_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    smle_free(dX); smle_free(dY);

    // cpu_spmm_v2.cpp:873-884
    double total_bytes = (double)a.num_nonzeros * (sizeof(V) * 2 + sizeof(int)) + (double)a.num_rows * k * (sizeof(int) + sizeof(V));
    double gflops = 2.0 * a.num_nonzeros * k / avg_ms / 1.0e6, gbs = total_bytes / avg_ms / 1.0e6;
    if (!quiet) printf("Merge CsrMM (B200), fp%d, k=%d: %.4f setup ms, %.4f avg ms, %.5f gflops, %.3lf effective GB/s\n",
                       int(sizeof(V) * 8), k, setup_ms, avg_ms, gflops, gbs);
    else printf("Merge CsrMM (B200), %.5f, %.5f, %.6f, %.3lf, ", setup_ms, avg_ms, gflops, gbs);
    printf("\n");
    return 0;
}

int main(int argc, char **argv)
{
    Args args(argc, argv);
    int device = 0;
    args.get("device", device);
    if (smle_init(device)) smle_adapters::die("smle_init");
    const bool quiet = args.flag("quiet");
    return args.flag("fp32") ? run<float>(args, quiet) : run<double>(args, quiet);
}
