// gather_ceiling.cu -- how many independent random gathers per second does a B200 SM sustain?
//
// Evidence for DESIGN.md section 4.2 (general tiles): two different decompositions of the R-MAT product
// (a thread per row with 8 gathers in flight; a gather per thread-item with 6 in flight on every thread)
// run at the same ~0.35 gathers per clock per SM, and L2 priorities do not move it.  This program takes
// the sparse kernel away: y[i] = sum_u x[idx[i, u]] with uniformly random idx, coalesced index reads,
// U independent loads per thread before the first use, full occupancy.  What it reaches is the ceiling of
// the SM's load path for scattered sectors (one 32 B sector per lane and request); the table size moves
// the data between L2 (32 MB), the L2 boundary (128 MB) and DRAM (1 GB).
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/gather_ceiling gather_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <typename V, int U>
__global__ void __launch_bounds__(512, (U <= 4 ? 4 : 2)) gather_kernel(const V *__restrict__ x, const unsigned *__restrict__ idx, V *__restrict__ y,
                                                        size_t n_out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        unsigned c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(idx + (size_t)u * n_out + i);   // coalesced, streamed
        V v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(x + c[u]);
        V s = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) s += v[u];
        __stcs(y + i, s);
    }
}

__global__ void fill_idx(unsigned *idx, size_t n, unsigned table, unsigned long long seed)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;   // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        idx[i] = (unsigned)(z % table);
    }
}

template <typename V, int U>
void run(const char *vname, size_t table_bytes, int sms, double mhz)
{
    const size_t table = table_bytes / sizeof(V);
    const size_t gathers = (size_t)1 << 28;   // 268 M gathers per launch (R-MAT scale 24 has as many nonzeros)
    const size_t n_out = gathers / U;
    V *x, *y;
    unsigned *idx;
    cudaMalloc(&x, table * sizeof(V));
    cudaMalloc(&y, n_out * sizeof(V));
    cudaMalloc(&idx, gathers * sizeof(unsigned));
    cudaMemset(x, 0, table * sizeof(V));
    fill_idx<<<sms * 8, 256>>>(idx, gathers, (unsigned)table, 12345);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_kernel<V, U>, 512, 0);
    const int grid = sms * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) gather_kernel<V, U><<<grid, 512>>>(x, idx, y, n_out);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) gather_kernel<V, U><<<grid, 512>>>(x, idx, y, n_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const cudaError_t err = cudaGetLastError();
    const double gps = (double)gathers / ms / 1e6;   // G gathers / s
    printf("{\"value_type\": \"%s\", \"loads_in_flight_per_thread\": %d, \"table_MB\": %zu, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"G_gathers_per_s\": %.1f, \"gathers_per_clk_per_sm\": %.3f, \"err\": \"%s\"}\n",
           vname, U, table_bytes >> 20, occ, ms, gps, gps * 1e3 / mhz / sms, cudaGetErrorString(err));
    fflush(stdout);
    cudaFree(x);
    cudaFree(y);
    cudaFree(idx);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_mhz_max\": %.0f}\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    for (size_t mb : {32, 128, 1024}) {
        run<double, 4>("f64", mb << 20, sms, mhz);
        run<double, 8>("f64", mb << 20, sms, mhz);
        run<double, 16>("f64", mb << 20, sms, mhz);
        run<float, 8>("f32", mb << 20, sms, mhz);
    }
    return 0;
}
