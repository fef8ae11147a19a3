#!/bin/bash
# round-2 batch g (1 GPU): driver + band tests; band-window sweep with dispatch messages; ncu captures exported as
# raw CSV (the .ncu-rep files stay on the box: three of them exceed what gpurun copies back)
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_spmm_band.py -q 2>&1 | tail -40) > gpurun_out/r02g_pytest_drivers.log; cat gpurun_out/r02g_pytest_drivers.log
(SWEEP_CG=1 SMLE_DEBUG_DISPATCH=1 timeout 300 python tools/sweep_spmm.py 200 32 sched0 band4 band16 band32) > gpurun_out/r02g_spmm_band.txt 2>&1; cat gpurun_out/r02g_spmm_band.txt
python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02g_bench_plain.json 2> gpurun_out/r02g_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02g_ncu_launches.log 2>&1; tail -1 gpurun_out/r02g_ncu_launches.log | cut -c1-200
prof() {  # name, kernel regex, skip, env..., -- command
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/r02g_plain_$name.log 2>&1 && ncu --set full --clock-control none -k regex:$regex -s $skip -c 1 -o /tmp/r02g_$name "$@" > gpurun_out/r02g_ncu_$name.log 2>&1 && ncu -i /tmp/r02g_$name.ncu-rep --page raw --csv > gpurun_out/r02g_raw_$name.csv 2>/dev/null
  tail -1 gpurun_out/r02g_ncu_$name.log | cut -c1-160
}
prof spmv_dot_cg300 spmv_kernel 3 python tools/prof_kernels.py cg 300
prof spmm32_200 spmm_rows_kernel 2 python tools/prof_kernels.py spmm32 200
SMLE_SPMM_BAND=1 prof spmm32_200_band spmm_rows_kernel 2 python tools/prof_kernels.py spmm32 200
ls -la gpurun_out | tail -15
