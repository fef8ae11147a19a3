#!/bin/bash
# round-2 batch t (1 GPU): full gpu suite, smoke, the bench line, grid2d 1000^2 and the scale-24 stress numbers after the
# skewed-matrix work and the producer / epilogue prefetches of the single-vector kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r02t_pytest.log; cat gpurun_out/r02t_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02t_smoke.log 2>&1; tail -2 gpurun_out/r02t_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02t_bench_n1.json) 2> gpurun_out/r02t_bench_n1.err; tail -2 gpurun_out/r02t_bench_n1.err; head -c 300 gpurun_out/r02t_bench_n1.json; echo
(timeout 200 python bench.py --workload spmv > gpurun_out/r02t_spmv_grid2d.jsonl) 2>&1 | tail -3; cat gpurun_out/r02t_spmv_grid2d.jsonl
(SMLE_RMAT_SCALE=24 timeout 400 python bench.py --workload stress > gpurun_out/r02t_stress24.jsonl) 2>&1 | tail -3; cut -c1-200 gpurun_out/r02t_stress24.jsonl
