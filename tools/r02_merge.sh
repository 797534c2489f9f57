#!/bin/bash
# two-level static merge: whole GPU suite, small-shape timelines, ordinary/pipelined A/B, host-call latency
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/small_timeline.py 2>&1 | tee gpurun_out/r02_small_timeline_c.jsonl
timeout 600 python tools/pipe_ab.py 2>&1 | tee gpurun_out/r02_pipe_ab_c.jsonl
timeout 300 python tools/host_profile.py 256x30000 32x30000 64x5000 1x30000 2>&1 | grep -v "OMP" | tee gpurun_out/r02_host_profile_c.txt
