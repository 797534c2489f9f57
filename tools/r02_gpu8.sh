#!/bin/bash
# one multi-GPU gpurun call (G = 8 by default): p2p parity incl. pipelined bursts, weak bench (carries strong + learner
# objects), config-5 shape (1024 actors in total), learner all-reduce behind the reference's R2D2 network
G=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/multi_gpu_check.py --exchange p2p --actors 37 --capacity 3000 2>&1 | grep "parity ok\|Error\|assert" | tail -$((G+2))
timeout 600 $TR --master-port 29602 bench.py --gpus $G --steps 200 --warmup 20 > gpurun_out/r02_bench_n$G.json 2> gpurun_out/r02_bench_n$G.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r02_bench_n$G.json
timeout 600 $TR --master-port 29603 bench.py --gpus $G --config 5 --steps 200 --warmup 20 > gpurun_out/r02_bench_n${G}_cfg5.json 2> gpurun_out/r02_bench_n${G}_cfg5.err; echo "cfg5 rc=$?"; cut -c1-400 gpurun_out/r02_bench_n${G}_cfg5.json
NGU_BENCH_PIPELINE=0 timeout 600 $TR --master-port 29605 bench.py --gpus $G --steps 200 --warmup 20 --no-parity --no-secondary > gpurun_out/r02_bench_n${G}_unpipelined.json 2> gpurun_out/r02_bench_n${G}_unpipelined.err; echo "unpiped rc=$?"; cut -c1-400 gpurun_out/r02_bench_n${G}_unpipelined.json
timeout 600 $TR --master-port 29604 tools/learner_dp_check.py > gpurun_out/r02_learner_dp_n$G.log 2>&1; echo "learner rc=$?"; grep learner_dp_check gpurun_out/r02_learner_dp_n$G.log | tail -1
