#!/bin/bash
# round-2 batch (1 GPU): full gpu tests; SpMM band-window sweep; ncu capture of the SpMV+dot kernel inside CG
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r02e_pytest.log; cat gpurun_out/r02e_pytest.log
(SWEEP_CG=1 timeout 400 python tools/sweep_spmm.py 200 32 sched0 band4 band8 band16 band32) > gpurun_out/r02e_spmm_band.txt 2>&1; cat gpurun_out/r02e_spmm_band.txt
python tools/prof_kernels.py cg 150 > gpurun_out/r02e_plain_cg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 4 -c 2 -o gpurun_out/r02e_prof_spmv_dot_cg150 python tools/prof_kernels.py cg 150 > gpurun_out/r02e_ncu_cg.log 2>&1; tail -2 gpurun_out/r02e_ncu_cg.log
