// oracle/dropin_parallel_efficiency.cpp -- the reference's verification/efficiency/
// parallel_efficiency.cpp, UNMODIFIED, on the GPU library.  TEST INFRASTRUCTURE ONLY; see
// dropin_singlecg.cpp.  The renamed callee is TestCGMultipleRHS at parallel_efficiency.cpp:102
// (CGSolveMultiple with the raw 1e-5 tolerance and SpmmKernel NONZERO_SPLIT).
#include <omp.h>
#include <mkl.h>
#include "sparse_matrix.h"
#include "utils.h"
#include "work_2025/hyper_parameters.hpp"
#include "work_2025/main/no_pretreatment.hpp"

#include "smle_adapters.hpp"

#define TestCGMultipleRHS TestGpuCGMultipleRHS
#include "verification/efficiency/parallel_efficiency.cpp"
