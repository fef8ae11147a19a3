#!/bin/bash
# round-2 batch r (1 GPU): one-CTA-per-SM configurations of the single-vector kernel on R-MAT (L1 capacity for the
# scattered gathers), the automatic pick for skewed matrices, parity subset
mkdir -p gpurun_out
for sc in 22 23; do for cfg in 640x6x2x1 640x4x2x1 960x3x2x1 960x2x2x1 960x2x3x1 480x6x2x1; do echo "== rmat1 $sc SMLE_SPMV_CFG=$cfg"; SMLE_SPMV_CFG=$cfg PROF_TIME=1 timeout 200 python tools/prof_kernels.py rmat1 $sc 2>&1 | grep "^spmv"; done; done > gpurun_out/r02r_one_cta_sweep.txt 2>&1; cat gpurun_out/r02r_one_cta_sweep.txt
for what in "rmat1 24" "wheel1 24"; do echo "== $what automatic pick"; SMLE_DEBUG_DISPATCH=1 PROF_TIME=1 timeout 300 python tools/prof_kernels.py $what 2>&1 | grep "^spmv\|smle"; done > gpurun_out/r02r_auto_pick.txt 2>&1; cat gpurun_out/r02r_auto_pick.txt
T="tests/test_gpu_spmv_spmm.py tests/test_gpu_partition.py tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py::test_wheel_2_20_hub_row"
(timeout 500 python -m pytest $T -q -x 2>&1 | tail -8) > gpurun_out/r02r_pytest.log; cat gpurun_out/r02r_pytest.log
