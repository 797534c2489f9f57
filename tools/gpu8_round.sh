#!/bin/bash
# one 8-GPU gpurun call: sharded parity (p2p), scaling bench at N=8 (weak, 256 actors/GPU), config-5 shape, learner all-reduce
set -x
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29601 tools/multi_gpu_check.py --exchange p2p --actors 37 --capacity 3000 2>&1 | grep -v OMP | tail -$((G+1))
timeout 300 $TR --master-port 29602 bench.py --gpus $G --steps 200 --warmup 20 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err; echo "rc=$?"; cat gpurun_out/bench_n$G.json; tail -3 gpurun_out/bench_n$G.err
timeout 300 $TR --master-port 29603 bench.py --gpus $G --steps 200 --warmup 20 --actors-per-gpu 128 > gpurun_out/bench_n${G}_cfg5.json 2> gpurun_out/bench_n${G}_cfg5.err; echo "rc=$?"; cat gpurun_out/bench_n${G}_cfg5.json
timeout 200 $TR --master-port 29604 tools/learner_allreduce_bench.py 2>&1 | grep -v OMP | tail -2 | tee gpurun_out/learner_allreduce_n$G.json
