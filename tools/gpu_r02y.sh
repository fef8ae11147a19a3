#!/bin/bash
# round-2 batch y (1 GPU): CG on systems that fit in L2 -- matrix stream with the normal L2 priority (default) against evict-first
mkdir -p gpurun_out
for keep in 0 96; do SMLE_SPMV_KEEP_MB=$keep timeout 200 python tools/small_cg_keep_ab.py 48 64 80 100 2>&1 | grep "^{"; done > gpurun_out/r02y_small_cg_keep_ab.jsonl; cat gpurun_out/r02y_small_cg_keep_ab.jsonl
