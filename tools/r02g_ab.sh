#!/bin/bash
# producer diet (rare side loads behind a branch): parity first, then A/B against the previous build on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipelined.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/lib_ab.py build/libprev.so ngu_b200/lib/libngu_epi.so | tee gpurun_out/r02g_lib_ab.jsonl
