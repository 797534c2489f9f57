#!/bin/bash
# full capture (with source counters) of the lean scan kernel for the bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --no-stream-only"
$CMD > gpurun_out/r02g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1000 -c 3 -f -o gpurun_out/r02g_prof_scan $CMD > gpurun_out/r02g_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02g_ncu_full.log | cut -c1-200
