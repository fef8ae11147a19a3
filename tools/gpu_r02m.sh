#!/bin/bash
# round-2 batch m (1 GPU): product-staged general tiles of the single-vector kernel -- parity subset under the three
# variants, A/B on R-MAT 22/23/24 and wheel 2^24; keep-in-L2 stream priority for small systems (grid2d 1000^2)
mkdir -p gpurun_out
T="tests/test_gpu_spmv_spmm.py tests/test_gpu_partition.py tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py::test_wheel_2_20_hub_row"
(timeout 500 python -m pytest $T -q -x 2>&1 | tail -8) > gpurun_out/r02m_pytest_default.log; cat gpurun_out/r02m_pytest_default.log
(SMLE_SPMV_DEBUG=8 timeout 300 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_cg.py -q -x -k "spmv or skewed or golden" 2>&1 | tail -5) > gpurun_out/r02m_pytest_l2pol.log; cat gpurun_out/r02m_pytest_l2pol.log
(timeout 900 python tools/ab_general_tiles.py rmat:22 rmat:23 rmat:24 rmat:24:f32 wheel:24 wheel:24:f32 > gpurun_out/r02m_general_tiles_ab.jsonl) 2>&1 | tail -5; cat gpurun_out/r02m_general_tiles_ab.jsonl
for keep in 0 96; do echo "== SMLE_SPMV_KEEP_MB=$keep"; SMLE_SPMV_KEEP_MB=$keep timeout 200 python bench.py --workload spmv 2>&1 | grep "^{"; done > gpurun_out/r02m_spmv_keep_ab.txt; cat gpurun_out/r02m_spmv_keep_ab.txt
