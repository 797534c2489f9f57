#!/bin/bash
# pipelined step: parity tests, then the A/B timing; real-CNN drop-in error probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipelined.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/pipe_ab.py "$@" 2>&1 | tee gpurun_out/r02_pipe_ab.jsonl
timeout 300 python tools/cnn_dropin_probe.py 2>&1 | grep -v "^No GPU\|^Using" | tee gpurun_out/r02_cnn_probe.jsonl
