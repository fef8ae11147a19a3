#!/bin/bash
# round-2 batch u (1 GPU): SpMM k=32 200^3 -- 8-lane workers with two vectors per lane (half the shared-memory broadcasts
# per dense-row load) at 2 / 3 / 4 loads in flight, 480-thread variants of the same, driver-chosen carve-out
mkdir -p gpurun_out
timeout 600 python tools/sweep_spmm.py 200 32 960x1920x2x1x4x1@2 carve0 carve44 960x1920x2x1x2x2@2 960x1920x2x1x3x2@2 960x1920x2x1x2x2@1 480x1920x2x1x4x2@2 480x1920x2x2x2x2@2 > gpurun_out/r02u_spmm_nv2_sweep.txt 2>&1; cat gpurun_out/r02u_spmm_nv2_sweep.txt
