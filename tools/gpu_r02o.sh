#!/bin/bash
# round-2 batch o (1 GPU): R-MAT general tiles -- more CTAs than resident (SMLE_SPMV_WAVES), one ncu capture (full set,
# source counters) of the product-staged path at scale 22
mkdir -p gpurun_out
for w in 1 2 4; do for sc in 22 23; do echo "== rmat1 $sc SMLE_SPMV_WAVES=$w"; SMLE_SPMV_WAVES=$w PROF_TIME=1 timeout 200 python tools/prof_kernels.py rmat1 $sc 2>&1 | grep "^spmv"; done; done > gpurun_out/r02o_waves_ab.txt 2>&1; cat gpurun_out/r02o_waves_ab.txt
ncu --set full --import-source on --clock-control none -k regex:spmv_kernel -s 3 -c 1 -o gpurun_out/r02o_rmat22_staged python tools/prof_kernels.py rmat1 22 > gpurun_out/r02o_ncu_rmat22.log 2>&1; tail -1 gpurun_out/r02o_ncu_rmat22.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep
