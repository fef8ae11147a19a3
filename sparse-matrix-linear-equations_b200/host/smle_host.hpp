// smle_host.hpp -- host-side support for the gpu_* drivers: a CsrMatrix-compatible container,
// Matrix Market reader/writer, the generator front-ends of the C ABI, the `--key[=value]`
// command line convention and the row-length statistics the reference drivers print.
//
// The drivers keep the CLI surface of the reference (cpu_spmv.cpp:925-991, cpu_spmm_v2.cpp,
// cpu_singlecg.cpp:219-279, cpu_multicg.cpp:293-333) so that eval_*.sh work with ./gpu_* in
// place of ./cpu_*.  Nothing here is on the hot path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/smle_b200.h"

namespace smle_host {

// Same field names and meaning as CsrMatrix<ValueT,int> (sparse_matrix.h:648-653).
template <typename V>
struct Csr {
    int num_rows = 0, num_cols = 0, num_nonzeros = 0;
    int *row_offsets = nullptr, *column_indices = nullptr;
    V *values = nullptr;
    std::vector<int> ro_, ci_;
    std::vector<V> va_;
    void adopt()
    {
        row_offsets = ro_.data(); column_indices = ci_.data(); values = va_.data();
    }
    void alloc(int m, int n, int nnz)
    {
        num_rows = m; num_cols = n; num_nonzeros = nnz;
        ro_.assign((size_t)m + 1, 0); ci_.assign((size_t)nnz, 0); va_.assign((size_t)nnz, V(0));
        adopt();
    }
};

template <typename V>
struct CooTuple { int row, col; V val; };

// COO -> CSR exactly as CsrMatrix::Init (sparse_matrix.h:668-733): stable sort by (row, col),
// duplicates kept, trailing empty rows filled.
template <typename V>
void coo_to_csr(std::vector<CooTuple<V>> &coo, int m, int n, Csr<V> &out)
{
    std::stable_sort(coo.begin(), coo.end(), [](const CooTuple<V> &a, const CooTuple<V> &b) {
        return a.row < b.row || (a.row == b.row && a.col < b.col);
    });
    out.alloc(m, n, (int)coo.size());
    int prev = -1;
    for (int z = 0; z < (int)coo.size(); ++z) {
        for (int r = prev + 1; r <= coo[z].row; ++r) out.ro_[r] = z;
        prev = coo[z].row;
        out.ci_[z] = coo[z].col;
        out.va_[z] = coo[z].val;
    }
    for (int r = prev + 1; r <= m; ++r) out.ro_[r] = (int)coo.size();
}

// Matrix Market reader with the reference's semantics (CooMatrix::InitMarket,
// sparse_matrix.h:211-380): "symmetric"/"skew" banners mirror off-diagonal entries, "array"
// files are dense column-major, missing values take default_value, indices are 1-based.
template <typename V>
bool read_matrix_market(const std::string &path, Csr<V> &out, V default_value = V(1))
{
    std::ifstream ifs(path.c_str());
    if (!ifs.good()) { fprintf(stderr, "Error opening file\n"); return false; }
    bool array = false, symmetric = false, skew = false, have_size = false;
    int m = 0, n = 0, declared = 0;
    long long dense_idx = 0;
    std::vector<CooTuple<V>> coo;
    std::string line;
    while (std::getline(ifs, line)) {
        if (line.empty()) continue;
        if (line[0] == '%') {
            if (line.size() > 1 && line[1] == '%') {
                symmetric = line.find("symmetric") != std::string::npos;
                skew = line.find("skew") != std::string::npos;
                array = line.find("array") != std::string::npos;
            }
            continue;
        }
        if (!have_size) {
            int got = sscanf(line.c_str(), "%d %d %d", &m, &n, &declared);
            if (!array && got == 3) coo.reserve((size_t)declared * (symmetric ? 2 : 1));
            else if (array && got >= 2) coo.reserve((size_t)m * n);
            else { fprintf(stderr, "Error parsing MARKET matrix: invalid problem description: %s\n", line.c_str()); return false; }
            have_size = true;
            continue;
        }
        if (array) {
            double v;
            if (sscanf(line.c_str(), "%lf", &v) != 1) { fprintf(stderr, "Error parsing MARKET matrix: badly formed value\n"); return false; }
            int col = (int)(dense_idx / m), row = (int)(dense_idx - (long long)m * col);
            coo.push_back({row, col, (V)v});
            ++dense_idx;
            continue;
        }
        const char *l = line.c_str();
        char *t = nullptr;
        long row = strtol(l, &t, 0);
        if (t == l) { fprintf(stderr, "Error parsing MARKET matrix: badly formed row\n"); return false; }
        l = t;
        long col = strtol(l, &t, 0);
        if (t == l) { fprintf(stderr, "Error parsing MARKET matrix: badly formed col\n"); return false; }
        l = t;
        double v = strtod(l, &t);
        if (t == l) v = (double)default_value;
        coo.push_back({(int)row - 1, (int)col - 1, (V)v});
        if (symmetric && row != col) coo.push_back({(int)col - 1, (int)row - 1, (V)(skew ? -v : v)});
    }
    if (!have_size) { fprintf(stderr, "Error parsing MARKET matrix: no size line\n"); return false; }
    coo_to_csr(coo, m, n, out);
    return true;
}

// Matrix Market writer (coordinate real general): lets the UNMODIFIED reference CG drivers,
// which only accept --mtx (cpu_singlecg.cpp:230), run on the generated Poisson / R-MAT inputs.
template <typename V>
bool write_matrix_market(const std::string &path, const Csr<V> &a)
{
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return false;
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n", a.num_rows, a.num_cols, a.num_nonzeros);
    for (int r = 0; r < a.num_rows; ++r)
        for (int z = a.row_offsets[r]; z < a.row_offsets[r + 1]; ++z)
            fprintf(f, "%d %d %.17g\n", r + 1, a.column_indices[z] + 1, (double)a.values[z]);
    fclose(f);
    return true;
}

// generator front-ends (CSR identical to the reference generator + CsrMatrix::Init)
#define SMLE_HOST_GEN(NAME, SHAPE_CALL, F64_CALL, F32_CALL)                                              \
    {                                                                                                    \
        int m, n, nnz;                                                                                   \
        if (SHAPE_CALL) return false;                                                                    \
        out.alloc(m, n, nnz);                                                                            \
        int rc;                                                                                          \
        if constexpr (sizeof(V) == 8) rc = F64_CALL; else rc = F32_CALL;                                 \
        return rc == 0;                                                                                  \
    }

template <typename V>
bool gen_grid2d(int w, bool self_loop, V diag, V offd, Csr<V> &out)
SMLE_HOST_GEN(grid2d, smle_gen_grid2d_shape(w, self_loop, &m, &n, &nnz),
              smle_gen_grid2d_f64(w, self_loop, diag, offd, out.row_offsets, out.column_indices, (double *)out.values),
              smle_gen_grid2d_f32(w, self_loop, diag, offd, out.row_offsets, out.column_indices, (float *)out.values))

template <typename V>
bool gen_grid3d(int w, bool self_loop, V diag, V offd, Csr<V> &out)
SMLE_HOST_GEN(grid3d, smle_gen_grid3d_shape(w, self_loop, &m, &n, &nnz),
              smle_gen_grid3d_f64(w, self_loop, diag, offd, out.row_offsets, out.column_indices, (double *)out.values),
              smle_gen_grid3d_f32(w, self_loop, diag, offd, out.row_offsets, out.column_indices, (float *)out.values))

template <typename V>
bool gen_wheel(int spokes, Csr<V> &out)
SMLE_HOST_GEN(wheel, smle_gen_wheel_shape(spokes, &m, &n, &nnz),
              smle_gen_wheel_f64(spokes, 1.0, out.row_offsets, out.column_indices, (double *)out.values),
              smle_gen_wheel_f32(spokes, 1.0f, out.row_offsets, out.column_indices, (float *)out.values))

template <typename V>
bool gen_dense(int rows, int cols, Csr<V> &out)
SMLE_HOST_GEN(dense, smle_gen_dense_shape(rows, cols, &m, &n, &nnz),
              smle_gen_dense_f64(rows, cols, 1.0, out.row_offsets, out.column_indices, (double *)out.values),
              smle_gen_dense_f32(rows, cols, 1.0f, out.row_offsets, out.column_indices, (float *)out.values))

template <typename V>
bool gen_rmat(int scale, int edge_factor, unsigned long long seed, Csr<V> &out)
SMLE_HOST_GEN(rmat, smle_gen_rmat_shape(scale, edge_factor, &m, &n, &nnz),
              smle_gen_rmat_f64(scale, edge_factor, 0.57, 0.19, 0.19, seed, 0, out.row_offsets, out.column_indices, (double *)out.values),
              smle_gen_rmat_f32(scale, edge_factor, 0.57, 0.19, 0.19, seed, 0, out.row_offsets, out.column_indices, (float *)out.values))
#undef SMLE_HOST_GEN

// `--key[=value]` arguments, the convention of the reference's CommandLineArgs (utils.h:278-520)
struct Args {
    std::vector<std::string> keys, vals;
    Args(int argc, char **argv)
    {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (a.size() < 3 || a[0] != '-' || a[1] != '-') continue;
            size_t eq = a.find('=');
            keys.push_back(a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2));
            vals.push_back(eq == std::string::npos ? "" : a.substr(eq + 1));
        }
    }
    bool flag(const char *k) const
    {
        for (auto &x : keys) if (x == k) return true;
        return false;
    }
    template <typename T> void get(const char *k, T &out) const
    {
        for (size_t i = 0; i < keys.size(); ++i)
            if (keys[i] == k && !vals[i].empty()) {
                if constexpr (std::is_same<T, std::string>::value) out = vals[i];
                else if constexpr (std::is_floating_point<T>::value) out = (T)atof(vals[i].c_str());
                else out = (T)atoll(vals[i].c_str());
            }
    }
};

// the row-length statistics GraphStats::Display prints in --quiet mode (sparse_matrix.h:72-106,
// computed as in CsrMatrix::Stats :897-921)
template <typename V>
void print_stats_csv(const Csr<V> &a)
{
    double mean = double(a.num_nonzeros) / a.num_rows, var = 0, skew = 0;
    for (int r = 0; r < a.num_rows; ++r) {
        double d = double(a.row_offsets[r + 1] - a.row_offsets[r]) - mean;
        var += d * d; skew += d * d * d;
    }
    var /= a.num_rows;
    double sd = sqrt(var);
    skew = (skew / a.num_rows) / pow(sd, 3.0);
    printf("%d, %d, %d, %.5f, %.5f, %.5f, %.5f, ", a.num_rows, a.num_cols, a.num_nonzeros, mean, sd, sd / mean, skew);
}

inline std::string base_name(const std::string &path)
{
    size_t s = path.find_last_of("/\\");
    std::string b = s == std::string::npos ? path : path.substr(s + 1);
    size_t d = b.find_last_of('.');
    return d == std::string::npos ? b : b.substr(0, d);
}

// one matrix from the shared generator / --mtx flags; returns its label ("" on failure)
template <typename V>
std::string matrix_from_args(const Args &args, Csr<V> &a, bool poisson_default)
{
    std::string mtx;
    int grid2d = -1, grid3d = -1, wheel = -1, dense = -1, rmat = -1, edge_factor = 16;
    args.get("mtx", mtx); args.get("grid2d", grid2d); args.get("grid3d", grid3d); args.get("wheel", wheel);
    args.get("dense", dense); args.get("rmat", rmat); args.get("edge_factor", edge_factor);
    const bool poisson = args.flag("poisson") || poisson_default;   // diag 4|6, off-diag -1, with self loop
    char label[256];
    if (!mtx.empty()) {
        if (!read_matrix_market(mtx, a)) return "";
        return mtx;
    } else if (grid2d > 0) {
        // the reference SpMV drivers call InitGrid2d(w, false) (cpu_spmv.cpp:783)
        if (!(poisson ? gen_grid2d<V>(grid2d, true, 4, -1, a) : gen_grid2d<V>(grid2d, args.flag("self_loop"), 1, 1, a))) return "";
        snprintf(label, sizeof label, "grid2d_%d", grid2d);
    } else if (grid3d > 0) {
        if (!(poisson ? gen_grid3d<V>(grid3d, true, 6, -1, a) : gen_grid3d<V>(grid3d, args.flag("self_loop"), 1, 1, a))) return "";
        snprintf(label, sizeof label, "grid3d_%d", grid3d);
    } else if (wheel > 0) {
        if (!gen_wheel<V>(wheel, a)) return "";
        snprintf(label, sizeof label, "wheel_%d", wheel);
    } else if (dense > 0) {
        int rows = (1 << 24) / dense;   // 16M nonzeros, as cpu_spmv.cpp:803
        if (!gen_dense<V>(rows, dense, a)) return "";
        snprintf(label, sizeof label, "dense_%d_x_%d", rows, dense);
    } else if (rmat > 0) {
        if (!gen_rmat<V>(rmat, edge_factor, 42, a)) return "";
        snprintf(label, sizeof label, "rmat_%d_%d", rmat, edge_factor);
    } else {
        fprintf(stderr, "No graph type specified.\n");
        return "";
    }
    return label;
}

} // namespace smle_host
