#!/bin/bash
# round-2 batch s (1 GPU): one CTA per SM, driver-default carve-out, tile sizes 2880..5760 on R-MAT; parity under the largest tile
mkdir -p gpurun_out
for sc in 22 23; do for cfg in 640x6x2x1 640x9x2x1 480x8x2x1 960x6x2x1 320x12x2x1 480x6x2x1; do echo "== rmat1 $sc SMLE_SPMV_CFG=$cfg"; SMLE_SPMV_CFG=$cfg PROF_TIME=1 timeout 200 python tools/prof_kernels.py rmat1 $sc 2>&1 | grep "^spmv\|rror"; done; done > gpurun_out/r02s_one_cta_sweep.txt 2>&1; cat gpurun_out/r02s_one_cta_sweep.txt
(SMLE_SPMV_CFG=640x9x2x1 timeout 300 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py::test_wheel_2_20_hub_row -q -x -k "spmv or skewed or golden or wheel" 2>&1 | tail -5) > gpurun_out/r02s_pytest_tile5760.log; cat gpurun_out/r02s_pytest_tile5760.log
