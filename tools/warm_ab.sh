#!/bin/bash
# A/B of the epilogue's instruction-cache warm-up pass (NGU_EPI_WARM): phase stamps and bench lines
mkdir -p gpurun_out
for w in 1 0; do
  echo "== NGU_EPI_WARM=$w"
  NGU_EPI_WARM=$w timeout 200 python tools/epilogue_phases.py 2>&1 | tail -3
  NGU_EPI_WARM=$w timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline 2>gpurun_out/warm$w.err | tee gpurun_out/warm$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], 'scan_ms', d['roofline']['launch_ms'])"
done
