#!/bin/bash
# one multi-GPU gpurun call: >=2-GPU tests, bench at N GPUs (weak headline + strong + parity + learner all-reduce)
G=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29602 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/r02g_bench_n$G.json 2> gpurun_out/r02g_bench_n$G.err; echo "bench rc=$?"; cat gpurun_out/r02g_bench_n$G.json; grep -v "OMP_NUM\|^\*\*\*\|^W1" gpurun_out/r02g_bench_n$G.err | tail -8
timeout 600 $TR --master-port 29603 tools/learner_dp_check.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*" | tail -3 | tee gpurun_out/r02g_learner_dp_n$G.json
