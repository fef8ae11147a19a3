#!/bin/bash
# round-2 batch n (1 GPU): ceilings that explain two numbers -- the scattered-gather rate of an SM (R-MAT) and what a
# plain copy of 80 MB reaches (grid2d 1000^2); the stream-only mode of the single-vector kernel on grid2d 1000^2
mkdir -p gpurun_out
timeout 300 tools/bin/gather_ceiling > gpurun_out/r02n_gather_ceiling.jsonl 2>&1; cat gpurun_out/r02n_gather_ceiling.jsonl
timeout 200 python tools/small_copy_ref.py > gpurun_out/r02n_small_copy_ref.jsonl 2>&1; cat gpurun_out/r02n_small_copy_ref.jsonl
for dbg in 0 1; do echo "== grid2d 1000 SMLE_SPMV_DEBUG=$dbg"; SMLE_SPMV_DEBUG=$dbg PROF_TIME=1 timeout 200 python tools/prof_kernels.py grid2d 1000 2>&1 | grep "^spmv"; done > gpurun_out/r02n_grid2d_stream_only.txt 2>&1; cat gpurun_out/r02n_grid2d_stream_only.txt
