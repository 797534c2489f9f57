#!/bin/bash
# end-of-round multi-GPU call: >=2-GPU tests on the final code, per-rank exchange wait / arrival skew
G=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/exchange_skew.py --actors 256 --capacity 30000 2> gpurun_out/r02f_skew_n${G}_256.err | grep '^{' | tee gpurun_out/r02f_skew_n${G}_256x30k.json
timeout 300 $TR --master-port 29612 tools/exchange_skew.py --actors 32 --capacity 30000 2> gpurun_out/r02f_skew_n${G}_32.err | grep '^{' | tee gpurun_out/r02f_skew_n${G}_32x30k.json
tail -3 gpurun_out/r02f_skew_n${G}_256.err | cut -c1-300
