#!/bin/bash
# round-2 batch zc (1 GPU): final state -- full gpu suite, smoke, the bench line as the driver runs it, scale-24 stress numbers
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r02zc_pytest.log; cat gpurun_out/r02zc_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02zc_smoke.log 2>&1; tail -2 gpurun_out/r02zc_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02zc_bench_n1.json) 2> gpurun_out/r02zc_bench_n1.err; tail -2 gpurun_out/r02zc_bench_n1.err; head -c 200 gpurun_out/r02zc_bench_n1.json; echo
(SMLE_RMAT_SCALE=24 timeout 400 python bench.py --workload stress > gpurun_out/r02zc_stress24.jsonl) 2>&1 | tail -3; cut -c1-175 gpurun_out/r02zc_stress24.jsonl
