#!/bin/bash
# round-2 batch l (1 GPU): full gpu test suite after the fp32 solvers and the default dealing of medium rows; smoke; bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r02l_pytest.log; cat gpurun_out/r02l_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02l_smoke.log 2>&1; tail -2 gpurun_out/r02l_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02l_bench_n1.json) 2> gpurun_out/r02l_bench_n1.err; tail -2 gpurun_out/r02l_bench_n1.err; head -c 300 gpurun_out/r02l_bench_n1.json; echo
(SMLE_RMAT_SCALE=24 timeout 400 python bench.py --workload stress > gpurun_out/r02l_stress24.jsonl) 2>&1 | tail -3
