// oracle/dropin_singlecg.cpp -- the reference's cpu_singlecg driver, UNMODIFIED, on the GPU library.
//
// TEST INFRASTRUCTURE ONLY (built by oracle/Makefile into oracle/_ref/, like the other reference
// builds).  It proves the drop-in boundary against the reference's REAL types: the translation unit
// is the reference's own cpu_singlecg.cpp (compiled where it lies under /root/reference, never
// copied) with its own headers -- sparse_matrix.h's CsrMatrix<double,int>, utils.h's
// CommandLineArgs, work_2025/types.hpp, hyper_parameters.hpp, single_strategy.hpp -- and the ONE
// change INTEGRATION.md describes: the callee at cpu_singlecg.cpp:101 is renamed from
// TestCGSolveSingle to the adapter TestGpuCGSolveSingle (host/smle_adapters.hpp).
// tests/test_gpu_dropin.py runs it next to the plain CPU build of the same file and compares the CSVs.

// 1. the reference headers first, under their own names (all are include-guarded, so the
//    #includes inside cpu_singlecg.cpp below become no-ops)
#include <omp.h>
#include <mkl.h>
#include "sparse_matrix.h"
#include "utils.h"
#include "work_2025/hyper_parameters.hpp"
#include "work_2025/main/single_strategy.hpp"

// 2. the adapters: the reference's signatures on top of the C ABI
#include "smle_adapters.hpp"

// 3. the reference driver with only the callee renamed
#define TestCGSolveSingle TestGpuCGSolveSingle
#include "cpu_singlecg.cpp"
