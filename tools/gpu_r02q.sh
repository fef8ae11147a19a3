#!/bin/bash
# round-2 batch q (1 GPU): one CTA per SM with tiles of 3840 items (L1 keeps ~92 KB instead of ~28 KB) on the other skewed inputs
mkdir -p gpurun_out
for what in "rmat1 23" "rmat1 24" "wheel1 24"; do for cfg in 480x6x2 640x6x2 960x4x2; do echo "== $what SMLE_SPMV_CFG=$cfg"; SMLE_SPMV_CFG=$cfg PROF_TIME=1 timeout 300 python tools/prof_kernels.py $what 2>&1 | grep "^spmv"; done; done > gpurun_out/r02q_cfg_skewed.txt 2>&1; cat gpurun_out/r02q_cfg_skewed.txt
