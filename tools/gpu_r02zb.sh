#!/bin/bash
# round-2 batch zb (1 GPU): two CTAs per SM with small stages (480x4x2x2: 76 KB per CTA, L1 ~100 KB, 30 warps) against the
# picked 640x6x2x1 (one CTA, L1 ~92 KB, 20 warps) on R-MAT fp64; fp32 back on its default
mkdir -p gpurun_out
for sc in 22 23 24; do for cfg in 480x4x2x2 640x6x2x1; do echo "== rmat1 $sc SMLE_SPMV_CFG=$cfg"; SMLE_SPMV_CFG=$cfg PROF_TIME=1 timeout 300 python tools/prof_kernels.py rmat1 $sc 2>&1 | grep "^spmv\|rror"; done; done > gpurun_out/r02zb_two_small_ctas.txt 2>&1; cat gpurun_out/r02zb_two_small_ctas.txt
