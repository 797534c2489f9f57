#!/bin/bash
# learner all-reduce hook behind the reference's R2D2 network on G GPUs: full log kept
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29603 tools/learner_dp_check.py > gpurun_out/r02_learner_dp_n$G.log 2>&1; echo "rc=$?"
grep -v "OMP_NUM\|^\*\*\*\|^W1\|No GPU\|Using GPU" gpurun_out/r02_learner_dp_n$G.log | tail -40
