#!/bin/bash
# round-2 batch v (1 GPU): the wheel's hub tiles and spoke tiles sit in different CTAs (contiguous runs): more, shorter runs
mkdir -p gpurun_out
for what in "wheel1 24" "rmat1 23"; do for w in 1 2 4 8; do echo "== $what SMLE_SPMV_WAVES=$w"; SMLE_SPMV_WAVES=$w PROF_TIME=1 timeout 300 python tools/prof_kernels.py $what 2>&1 | grep "^spmv"; done; done > gpurun_out/r02v_waves_skewed.txt 2>&1; cat gpurun_out/r02v_waves_skewed.txt
