#!/bin/bash
# round-2 batch zd (1 GPU): stages of 1440 items (480x3x2x2: 61 KB per CTA, L1 ~124 KB) against the adopted 480x4x2x2 on R-MAT 23 / 24;
# parity subset with every matrix forced onto that configuration
mkdir -p gpurun_out
for sc in 23 24; do echo "== rmat1 $sc SMLE_SPMV_CFG=480x3x2x2"; SMLE_SPMV_CFG=480x3x2x2 PROF_TIME=1 timeout 200 python tools/prof_kernels.py rmat1 $sc 2>&1 | grep "^spmv\|rror"; done > gpurun_out/r02zd_stage1440.txt 2>&1; cat gpurun_out/r02zd_stage1440.txt
(SMLE_SPMV_CFG=480x3x2x2 timeout 120 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_cg.py -q -x -k "spmv or skewed or golden" 2>&1 | tail -3) > gpurun_out/r02zd_pytest.log; cat gpurun_out/r02zd_pytest.log
