// smle_common.cuh -- shared device helpers for the sm_100a merge-path SpMV/SpMM/CG kernels.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <string.h>

namespace smle {

constexpr int kThreads = 256;          // threads per CTA for every kernel in the library
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------------------------------
// CG scalar state, device resident (all arrays have k entries).  The iteration loop never
// reads these on the host except through one async copy of `ctrl` per graph batch.
// ---------------------------------------------------------------------------------------
enum CtrlSlot { CTRL_ITER = 0, CTRL_STOP = 1, CTRL_HALT = 2, CTRL_MAX_ITERS = 3, CTRL_NCONV = 4,
                CTRL_SEQ_BASE = 5,   // multi-GPU: sequence numbers of this solve start here
                CTRL_WORDS = 8 };

struct CgScalars {
    double *rs_old;    // r.r of the previous iteration          (no_pretreatment.hpp:81,179-181)
    double *rs_new;    // r.r of this iteration                  (:130)
    double *pAp;       // p.Ap                                   (:107)
    double *alpha;     // rs_old/pAp, 0 once latched             (:109-120)
    double *beta;      // rs_new/rs_old, 0 once latched          (:165-176)
    double *bnorm;     // ||b||, 0 replaced by 1                 (:71-79)
    int *conv;         // per-column convergence latch           (:133-155)
    int *ctrl;         // CtrlSlot words
    double *hist;      // max relative residual per iteration    (:150-155), nullable
    double *last_rel;  // max relative residual of the last iteration (1 double)
    int hist_cap;
    double *tol;       // relative tolerance, device resident so CUDA graphs stay valid across calls
    int dot_mode;      // what the fused dot product of an SpMV / SpMM<DOT> launch becomes (cg_dot_scalars)
};

// The last CTA of an SpMV / SpMM<DOT> launch turns the fused dot product X[:,c].Y[:,c] into the scalars
// of the solver that launched it:
//   0  CG        p.Ap -> alpha = rs_old / pAp, 0 once latched     (no_pretreatment.hpp:107-120)
//   1  SPAI-PCG  p.Ap -> alpha, also 0 when pAp == 0              (sparse_approximate_inverse.hpp:119-127)
//   2  SPAI-PCG  r.z with z = M r -> beta = rs_new / rs_old (0 once latched or rs_old == 0),
//                rs_old <- rs_new                                  (:183-192)
//   3  SPAI-PCG  initial r.z -> rs_old, beta = 0 (so that p = z + beta p starts as z)  (:88-94)
enum DotMode { DOT_CG_ALPHA = 0, DOT_PCG_ALPHA = 1, DOT_PCG_BETA = 2, DOT_PCG_INIT = 3 };

__device__ __forceinline__ void cg_dot_scalars(const CgScalars &cg, int c, double dot)
{
    const bool latched = cg.conv[c] != 0;
    if (cg.dot_mode == DOT_CG_ALPHA) {
        cg.pAp[c] = dot;
        cg.alpha[c] = latched ? 0.0 : cg.rs_old[c] / dot;
    } else if (cg.dot_mode == DOT_PCG_ALPHA) {
        cg.pAp[c] = dot;
        cg.alpha[c] = (!latched && dot != 0.0) ? cg.rs_old[c] / dot : 0.0;
    } else if (cg.dot_mode == DOT_PCG_BETA) {
        const double ro = cg.rs_old[c];
        cg.rs_new[c] = dot;
        cg.beta[c] = (!latched && ro != 0.0) ? dot / ro : 0.0;
        cg.rs_old[c] = dot;
    } else {
        cg.rs_new[c] = dot;
        cg.rs_old[c] = dot;
        cg.beta[c] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// vector load / store of VEC consecutive values (VEC*sizeof(V) in {4, 8, 16} bytes)
// ---------------------------------------------------------------------------------------
template <typename V, int VEC>
__device__ __forceinline__ void ldg_vec(V (&out)[VEC], const V *p)
{
    if constexpr (sizeof(V) * VEC == 16) {
        int4 t = __ldg(reinterpret_cast<const int4 *>(p));
        memcpy(out, &t, 16);
    } else if constexpr (sizeof(V) * VEC == 8) {
        int2 t = __ldg(reinterpret_cast<const int2 *>(p));
        memcpy(out, &t, 8);
    } else {
        int t = __ldg(reinterpret_cast<const int *>(p));
        memcpy(out, &t, 4);
    }
}

// coherent (L2) load: for data another CTA of the same launch may have written
template <typename V, int VEC>
__device__ __forceinline__ void ldcg_vec(V (&out)[VEC], const V *p)
{
    if constexpr (sizeof(V) * VEC == 16) {
        int4 t = __ldcg(reinterpret_cast<const int4 *>(p));
        memcpy(out, &t, 16);
    } else if constexpr (sizeof(V) * VEC == 8) {
        int2 t = __ldcg(reinterpret_cast<const int2 *>(p));
        memcpy(out, &t, 8);
    } else {
        int t = __ldcg(reinterpret_cast<const int *>(p));
        memcpy(out, &t, 4);
    }
}

template <typename V, int VEC>
__device__ __forceinline__ void ld_vec(V (&out)[VEC], const V *p)
{
    if constexpr (sizeof(V) * VEC == 16) {
        int4 t = *reinterpret_cast<const int4 *>(p);
        memcpy(out, &t, 16);
    } else if constexpr (sizeof(V) * VEC == 8) {
        int2 t = *reinterpret_cast<const int2 *>(p);
        memcpy(out, &t, 8);
    } else {
        int t = *reinterpret_cast<const int *>(p);
        memcpy(out, &t, 4);
    }
}

template <typename V, int VEC>
__device__ __forceinline__ void st_vec(V *p, const V (&in)[VEC])
{
    if constexpr (sizeof(V) * VEC == 16) {
        int4 t;
        memcpy(&t, in, 16);
        *reinterpret_cast<int4 *>(p) = t;
    } else if constexpr (sizeof(V) * VEC == 8) {
        int2 t;
        memcpy(&t, in, 8);
        *reinterpret_cast<int2 *>(p) = t;
    } else {
        int t;
        memcpy(&t, in, 4);
        *reinterpret_cast<int *>(p) = t;
    }
}

// ---------------------------------------------------------------------------------------
// L2 residency control.  B200 has a 126 MB L2: the CG vectors of the single-RHS configs fit,
// the matrix does not.  Every global access of the CG kernels therefore carries an explicit
// eviction priority: streams that are dead after this access (the CSR arrays, AP's last read,
// x) are evict_first, vectors that the NEXT kernel re-reads (p for the SpMV gathers, r) are
// evict_last.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t make_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ double2 ld_f64x2_hint(const double *p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ double ld_f64_hint(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// read-only (nc) scalar gather with an L2 eviction priority
__device__ __forceinline__ double ldg_hint(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ float ldg_hint(const float *p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ void st_f64x2_hint(double *p, double2 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

__device__ __forceinline__ void st_f64_hint(double *p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The three kernels of a CG iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, inside the CUDA graph too: kernel n+1 is
// scheduled onto the SMs while the tail of kernel n (its last CTA's serial epilogue) still runs,
// and blocks in griddep_wait() until kernel n has completed and its writes are visible.  Every
// kernel triggers its dependents only AFTER its own wait, so when a kernel starts, everything
// two launches back is complete.  Both are no-ops for launches without the attribute.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// "last CTA done" election (threadFenceReduction pattern): every CTA publishes its global
// writes, takes a ticket, and the CTA that draws the last ticket runs the serial epilogue
// (carry fix-up, deterministic reduction of the per-CTA partials, CG scalars).  The ticket
// counter is reset by the winner so the next launch starts from zero.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_cta_election(unsigned int *ticket, unsigned int num_ctas)
{
    __shared__ bool s_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        s_is_last = (t == num_ctas - 1u);
    }
    __syncthreads();
    bool last = s_is_last;
    if (last) {
        __threadfence();
        if (threadIdx.x == 0) *ticket = 0u;
    }
    return last;
}

// ---------------------------------------------------------------------------------------
// Deterministic column-wise reduction of per-CTA partials by ONE CTA:
//   out[c] = sum_e (src1[e*k+c] + (src2 ? src2[e*k+c] : 0)),  e in [0, entries)
// The order of additions is fixed by (blockDim.x, entries, k) only, so results are
// reproducible run to run.  red: shared scratch of blockDim.x values.
// ---------------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ void cta_reduce_columns(const V *src1, const V *src2, int entries, int k,
                                                   V *out, V *red)
{
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int cbase = 0; cbase < k; cbase += nthreads) {
        const int kk = min(k - cbase, nthreads);       // columns in this pass
        const int parts = nthreads / kk;               // threads cooperating per column
        const int c = tid % kk, part = tid / kk;
        V s = 0;
        if (part < parts) {
            for (int e = part; e < entries; e += parts) {
                size_t o = (size_t)e * k + cbase + c;
                s += __ldcg(src1 + o);
                if (src2) s += __ldcg(src2 + o);
            }
        }
        __syncthreads();
        red[tid] = s;
        __syncthreads();
        if (tid < kk) {
            V t = 0;
            for (int p = 0; p < parts; ++p) t += red[p * kk + tid];
            out[cbase + tid] = t;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Carry slots: a bit pattern no arithmetic result can have marks an empty slot.
// ---------------------------------------------------------------------------------------
template <typename V> struct CarrySentinel;
template <> struct CarrySentinel<double> { static constexpr unsigned long long bits = 0xFFF75EA1C0DED00Dull; };
template <> struct CarrySentinel<float> { static constexpr unsigned int bits = 0xFFA5C0DEu; };

template <typename V>
__device__ __forceinline__ bool is_sentinel(V v)
{
    if constexpr (sizeof(V) == 8) return (unsigned long long)__double_as_longlong(v) == CarrySentinel<double>::bits;
    else return __float_as_uint(v) == CarrySentinel<float>::bits;
}

template <typename V>
__device__ __forceinline__ V sentinel_value()
{
    if constexpr (sizeof(V) == 8) return __longlong_as_double((long long)CarrySentinel<double>::bits);
    else return __uint_as_float(CarrySentinel<float>::bits);
}

template <typename V>
__global__ void fill_sentinel_kernel(V *p, size_t count)
{
    const V s = sentinel_value<V>();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        p[i] = s;
}

// ---- wait-free carry exchange ------------------------------------------------------------------
// Slot t holds the sentinel until one of the two parties of the row cut by the boundary between
// tiles t and t+1 arrives: the tile that has the row's leading part (publisher) or the tile the
// row continues in (owner).  Each swaps its value in; whoever finds the other's value there
// finishes the row (owner part + carry, the reference's order) and re-arms the slot.  Nobody
// ever waits, so the schedule of tiles over CTAs is free.
template <typename V>
__device__ __forceinline__ V slot_exchange(V *slot, V mine)
{
    if constexpr (sizeof(V) == 8) {
        const unsigned long long o = atomicExch(reinterpret_cast<unsigned long long *>(slot),
                                                (unsigned long long)__double_as_longlong(mine));
        return __longlong_as_double((long long)o);
    } else {
        const unsigned int o = atomicExch(reinterpret_cast<unsigned int *>(slot), __float_as_uint(mine));
        return __uint_as_float(o);
    }
}

template <typename V>
__device__ __forceinline__ void slot_reset(V *slot)
{
    *reinterpret_cast<volatile V *>(slot) = sentinel_value<V>();
}

// ---------------------------------------------------------------------------------------
// Single-column variant of the reduction above (k = 1 kernels): a fixed shuffle tree instead of a
// serial pass of one thread over blockDim.x shared-memory values (~2 us on the critical path of
// every CG kernel's last CTA).  Order fixed by (blockDim.x, entries): deterministic.  All threads
// of the CTA must call it; every thread returns the total.  red: >= 33 values of shared scratch.
// ---------------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ V cta_reduce_one(const V *src1, const V *src2, int entries, V *red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
    V s = 0;
    for (int e = tid; e < entries; e += blockDim.x) {
        s += __ldcg(src1 + e);
        if (src2) s += __ldcg(src2 + e);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        V t = lane < nw ? red[lane] : V(0);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

} // namespace smle
