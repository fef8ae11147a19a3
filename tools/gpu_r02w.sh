#!/bin/bash
# round-2 batch w (2 GPUs): the multi-GPU tests and the N = 2 bench line after today's changes to the single-vector kernel
# (producer metadata prefetch feeds the halo flag, carry chain); wheel timing after the shorter carry-chain link
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_dist.py "tests/test_gpu_drivers.py::test_gpu_multicg_columns_sharded_over_two_gpus" "tests/test_gpu_drivers.py::test_gpu_singlecg_row_partitioned_over_two_gpus" -q 2>&1 | tail -15) > gpurun_out/r02w_pytest_n2.log; cat gpurun_out/r02w_pytest_n2.log
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r02w_bench_n2.json) 2> gpurun_out/r02w_bench_n2.err; tail -2 gpurun_out/r02w_bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02w_bench_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_iteration'],d['e2e']['value'],d['roofline']['kernel_ms'],d['parity_check'])"
for what in "wheel1 24" "wheel1 20"; do echo "== $what"; PROF_TIME=1 timeout 300 python tools/prof_kernels.py $what 2>&1 | grep "^spmv"; done > gpurun_out/r02w_wheel.txt 2>&1; cat gpurun_out/r02w_wheel.txt
(timeout 200 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_baseline_sizes.py -q -x 2>&1 | tail -3) > gpurun_out/r02w_pytest_subset.log; cat gpurun_out/r02w_pytest_subset.log
