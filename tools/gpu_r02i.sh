#!/bin/bash
# round-2 batch i (1 GPU): full gpu test suite, the bench line as the driver runs it (with extras and cpu_baseline),
# smoke(), one ncu capture of the general-tile SpMV on R-MAT scale 23
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02i_smoke.log 2>&1; tail -2 gpurun_out/r02i_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02i_bench_n1.json) 2> gpurun_out/r02i_bench_n1.err; tail -2 gpurun_out/r02i_bench_n1.err; head -c 400 gpurun_out/r02i_bench_n1.json; echo
python tools/prof_kernels.py rmat1 23 > gpurun_out/r02i_plain_rmat23.log 2>&1 && ncu --set full --clock-control none -k regex:spmv_kernel -s 3 -c 1 -o /tmp/r02i_rmat23 python tools/prof_kernels.py rmat1 23 > gpurun_out/r02i_ncu_rmat23.log 2>&1 && ncu -i /tmp/r02i_rmat23.ncu-rep --page raw --csv > gpurun_out/r02i_raw_spmv_rmat23.csv 2>/dev/null; tail -1 gpurun_out/r02i_ncu_rmat23.log | cut -c1-160
PROF_TIME=1 python tools/prof_kernels.py rmat1 23 2>&1 | tail -2
