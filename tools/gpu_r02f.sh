#!/bin/bash
# round-2 batch f (1 GPU): driver tests with full output; band-window engagement check; SpMV configuration sweep inside
# CG at 150^3; ncu: launch list of the bench command, --set full of the dominant kernel at the bench size (300^3) and
# of the SpMM k=32 kernels at 200^3
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_spmm_band.py -q 2>&1 | tail -60) > gpurun_out/r02f_pytest_drivers.log; cat gpurun_out/r02f_pytest_drivers.log
(SMLE_DEBUG_DISPATCH=1 timeout 200 python tools/sweep_spmm.py 200 32 sched0 band16 band2) > gpurun_out/r02f_spmm_band.txt 2>&1; cat gpurun_out/r02f_spmm_band.txt
(timeout 300 python tools/sweep_spmv.py 150 480x6x2 480x4x3 480x3x4 320x6x3 480x5x2) > gpurun_out/r02f_spmv_cfg_cg150.txt 2>&1; cat gpurun_out/r02f_spmv_cfg_cg150.txt
python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02f_bench_plain.json 2> gpurun_out/r02f_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02f_ncu_launches.log 2>&1; tail -2 gpurun_out/r02f_ncu_launches.log
python tools/prof_kernels.py cg 300 > gpurun_out/r02f_plain_cg300.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 3 -c 1 -o gpurun_out/r02f_prof_spmv_dot_cg300 python tools/prof_kernels.py cg 300 > gpurun_out/r02f_ncu_cg300.log 2>&1; tail -2 gpurun_out/r02f_ncu_cg300.log
python tools/prof_kernels.py spmm32 200 > gpurun_out/r02f_plain_spmm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 2 -c 1 -o gpurun_out/r02f_prof_spmm32_200 python tools/prof_kernels.py spmm32 200 > gpurun_out/r02f_ncu_spmm.log 2>&1; tail -2 gpurun_out/r02f_ncu_spmm.log
SMLE_SPMM_BAND=1 python tools/prof_kernels.py spmm32 200 > gpurun_out/r02f_plain_spmm_band.log 2>&1 && SMLE_SPMM_BAND=1 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 2 -c 1 -o gpurun_out/r02f_prof_spmm32_200_band python tools/prof_kernels.py spmm32 200 > gpurun_out/r02f_ncu_spmm_band.log 2>&1; tail -2 gpurun_out/r02f_ncu_spmm_band.log
