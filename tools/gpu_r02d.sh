#!/bin/bash
# round-2 multi-GPU batch (8 GPUs): dist tests at world 4/8, per-rank kernel timeline, bench at N=8 and N=4,
# mailbox vs NCCL all-reduce A/B
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "8-slab or 8-global or 4-random" 2>&1 | tail -8) > gpurun_out/r02d_pytest_n8.log; cat gpurun_out/r02d_pytest_n8.log
(timeout 200 $TR --nproc-per-node 8 --master-port 29821 bench.py --workload rowcg_fixed > gpurun_out/r02d_rowcg_n8.jsonl) 2>&1 | tail -3; cat gpurun_out/r02d_rowcg_n8.jsonl
(timeout 300 $TR --nproc-per-node 8 --master-port 29822 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02d_bench_n8.json) 2> gpurun_out/r02d_bench_n8.err; tail -3 gpurun_out/r02d_bench_n8.err; head -c 600 gpurun_out/r02d_bench_n8.json; echo
(timeout 200 $TR --nproc-per-node 4 --master-port 29823 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/r02d_bench_n4.json) 2> gpurun_out/r02d_bench_n4.err; tail -3 gpurun_out/r02d_bench_n4.err; head -c 600 gpurun_out/r02d_bench_n4.json; echo
(timeout 120 $TR --nproc-per-node 8 --master-port 29824 tools/ab_allreduce.py > gpurun_out/r02d_ab_allreduce_n8.json) 2>&1 | tail -3; cat gpurun_out/r02d_ab_allreduce_n8.json
(timeout 200 $TR --nproc-per-node 4 --master-port 29825 bench.py --workload rowcg_fixed > gpurun_out/r02d_rowcg_n4.jsonl) 2>&1 | tail -3; cat gpurun_out/r02d_rowcg_n4.jsonl
