#!/bin/bash
# round-2 batch p (1 GPU): product-staged general tiles after the reduction changes (4-way partial sums, huge segments
# finished by one warp each behind a single extra barrier) -- parity subset, A/B, configuration sweep on R-MAT 22
mkdir -p gpurun_out
T="tests/test_gpu_spmv_spmm.py tests/test_gpu_partition.py tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py::test_wheel_2_20_hub_row"
(timeout 500 python -m pytest $T -q -x 2>&1 | tail -8) > gpurun_out/r02p_pytest_default.log; cat gpurun_out/r02p_pytest_default.log
(timeout 900 python tools/ab_general_tiles.py rmat:22 rmat:23 rmat:24 wheel:24 wheel:24:f32 > gpurun_out/r02p_general_tiles_ab.jsonl) 2>&1 | tail -5; cut -c1-330 gpurun_out/r02p_general_tiles_ab.jsonl
for cfg in 480x6x2 224x8x2 320x6x3 640x6x2 480x4x3 256x12x2 960x4x2; do echo "== rmat1 22 SMLE_SPMV_CFG=$cfg"; SMLE_SPMV_CFG=$cfg PROF_TIME=1 timeout 200 python tools/prof_kernels.py rmat1 22 2>&1 | grep "^spmv"; done > gpurun_out/r02p_cfg_sweep_rmat22.txt 2>&1; cat gpurun_out/r02p_cfg_sweep_rmat22.txt
