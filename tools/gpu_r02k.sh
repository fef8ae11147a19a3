#!/bin/bash
# round-2 batch k (2 GPUs): column-sharded multi-RHS CG tests (Python and the forked C++ driver), the row-partitioned
# tests again after the last kernel changes; on GPU 0 also the A/B of the queue threshold of the general-tile SpMV path
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_dist.py "tests/test_gpu_drivers.py::test_gpu_multicg_columns_sharded_over_two_gpus" "tests/test_gpu_drivers.py::test_gpu_singlecg_row_partitioned_over_two_gpus" -q 2>&1 | tail -15) > gpurun_out/r02k_pytest_n2.log; cat gpurun_out/r02k_pytest_n2.log
for what in "rmat1 22" "rmat1 23"; do for med in 32 16 8; do echo "== $what SMLE_SPMV_MEDLO=$med"; SMLE_SPMV_MEDLO=$med PROF_TIME=1 timeout 200 python tools/prof_kernels.py $what 2>&1 | grep "^spmv"; done; done > gpurun_out/r02k_medlo_ab.txt 2>&1; cat gpurun_out/r02k_medlo_ab.txt
