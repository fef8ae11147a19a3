#!/bin/bash
# round-2 batch h (2 GPUs): row-partitioned tests at world 2 (planner, both halo paths, host buffers, the forked C++ driver)
# and the bench at N = 2 after the K3 change (boundary rows first, fence in pushing threads only)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_dist.py "tests/test_gpu_drivers.py::test_gpu_singlecg_row_partitioned_over_two_gpus" -q 2>&1 | tail -15) > gpurun_out/r02h_pytest_n2.log; cat gpurun_out/r02h_pytest_n2.log
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r02h_bench_n2.json) 2> gpurun_out/r02h_bench_n2.err; tail -2 gpurun_out/r02h_bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_iteration'],d['e2e']['value'],d['roofline']['kernel_ms'],d['parity_check'])"
