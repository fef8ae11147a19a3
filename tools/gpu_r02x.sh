#!/bin/bash
# round-2 batch x (1 GPU): final state -- full gpu suite, smoke, the bench line as the driver runs it, one ncu capture of the
# single-vector kernel on R-MAT scale 23 in the configuration the handle now picks (640x6x2, one CTA per SM)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r02x_pytest.log; cat gpurun_out/r02x_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02x_smoke.log 2>&1; tail -2 gpurun_out/r02x_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02x_bench_n1.json) 2> gpurun_out/r02x_bench_n1.err; tail -2 gpurun_out/r02x_bench_n1.err; head -c 300 gpurun_out/r02x_bench_n1.json; echo
ncu --set full --import-source on --clock-control none -k regex:spmv_kernel -s 3 -c 1 -o gpurun_out/r02x_rmat23_final python tools/prof_kernels.py rmat1 23 > gpurun_out/r02x_ncu_rmat23.log 2>&1; tail -1 gpurun_out/r02x_ncu_rmat23.log | cut -c1-160
