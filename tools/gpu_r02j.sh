#!/bin/bash
# round-2 batch j (1 GPU): A/B of the balanced dealing of medium rows in the general-tile SpMV path (SMLE_SPMV_DEBUG=2)
mkdir -p gpurun_out
(SMLE_SPMV_DEBUG=2 timeout 400 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_partition.py "tests/test_gpu_baseline_sizes.py::test_wheel_2_20_hub_row" -q 2>&1 | tail -8) > gpurun_out/r02j_pytest_balance.log; cat gpurun_out/r02j_pytest_balance.log
for what in "rmat1 22" "rmat1 23" "wheel1 24"; do for dbg in 0 2; do echo "== $what SMLE_SPMV_DEBUG=$dbg"; SMLE_SPMV_DEBUG=$dbg PROF_TIME=1 timeout 200 python tools/prof_kernels.py $what 2>&1 | grep "^spmv"; done; done > gpurun_out/r02j_balance_ab.txt 2>&1; cat gpurun_out/r02j_balance_ab.txt
