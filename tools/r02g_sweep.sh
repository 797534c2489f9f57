#!/bin/bash
# schedule / prefetch / taper knobs re-swept on the lean scan kernel (ordinary, pipelined, scan alone, host call)
mkdir -p gpurun_out
python tools/knob_sweep.py 256 30000 2 - "NGU_EPI_SCHED=0.6,8" "NGU_EPI_SCHED=0.6,12" "NGU_EPI_SCHED=0.7,16" "NGU_EPI_SCHED=0.5,16" "NGU_EPI_SCHED=0.7,8" "NGU_EPI_SCHED=0.8,8" "NGU_EPI_SCHED=0.6,24" \
   "NGU_EPI_PREFETCH_TILES=2" "NGU_EPI_PREFETCH_TILES=8" "NGU_EPI_TAPER=0.2,0.1" "NGU_EPI_TAPER=0.3,0.2" | tee gpurun_out/r02g_knob_sweep.jsonl
