#!/bin/bash
# round-2 first look: where do the small / strong-scaling shapes spend their step?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 300 python tools/small_timeline.py 1x30000 64x5000 32x30000 64x30000 128x30000 2>&1 | grep -v OMP | tee gpurun_out/r02_small_timeline.jsonl
timeout 300 python tools/epilogue_phases.py 32x30000 64x5000 1x30000 2>&1 | grep -v OMP | tee gpurun_out/r02_epilogue_phases.jsonl
timeout 300 python tools/host_profile.py 256x30000 64x5000 1x30000 32x30000 2>&1 | grep -v OMP | tee gpurun_out/r02_host_profile.txt
