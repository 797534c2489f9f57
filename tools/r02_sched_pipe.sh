#!/bin/bash
# does a dynamic tail help the PIPELINED step on mid-size shards (late-starting CTAs take less work)?
mkdir -p gpurun_out
for S in default 0.5,4,8 0.6,4,8 0.4,4,8 0.3,4,8 0.5,6,8 0.5,2,8 0.7,3,8; do
  echo "== NGU_EPI_SCHED=$S"
  if [ "$S" = default ]; then python tools/pipe_ab.py "$@"; else NGU_EPI_SCHED=$S python tools/pipe_ab.py "$@"; fi
done 2>&1 | grep -v "^$" | tee gpurun_out/r02_sched_pipe.txt
