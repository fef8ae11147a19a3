#!/bin/bash
# round-2 experiment batch (1 GPU): full gpu tests, bench, SpMM schedule A/B, small-problem SpMV sweep, scale-24 stress
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_pytest.log
(timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_n1.json) 2> gpurun_out/r02c_bench_n1.err; tail -3 gpurun_out/r02c_bench_n1.err
for k in 8 16 32; do SWEEP_CG=1 timeout 200 python tools/sweep_spmm.py 200 $k sched0 sched1; done > gpurun_out/r02c_spmm_sched.txt 2>&1; cat gpurun_out/r02c_spmm_sched.txt
for cfg in 480x6x2 480x4x3 480x3x4 320x6x3 640x6x2 224x8x2; do echo "cfg $cfg"; SMLE_SPMV_CFG=$cfg timeout 120 python bench.py --workload spmv; done > gpurun_out/r02c_spmv_c1.txt 2>&1; cat gpurun_out/r02c_spmv_c1.txt
(SMLE_RMAT_SCALE=24 timeout 500 python bench.py --workload stress > gpurun_out/r02c_stress24.jsonl) 2>&1 | tail -3; cat gpurun_out/r02c_stress24.jsonl
