#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/small_timeline.py 1x30000 64x5000 32x30000 8x30000 16x30000 2>&1 | grep -v OMP | tee gpurun_out/r02_small_timeline_b.jsonl
