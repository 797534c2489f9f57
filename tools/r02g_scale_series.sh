#!/bin/bash
# same-box scaling series N = 1, 2, 4, 8 (what the driver does at round end), short lines: parity on, secondary off
mkdir -p gpurun_out
for G in 1 2 4 8; do
  if [ $G -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-secondary --no-cpu-baseline > gpurun_out/r02g_scale_n$G.json 2> gpurun_out/r02g_scale_n$G.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29700+G)) bench.py --gpus $G --steps 100 --warmup 10 > gpurun_out/r02g_scale_n$G.json 2> gpurun_out/r02g_scale_n$G.err
  fi
  echo "N=$G rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r02g_scale_n$G.json").read().strip().splitlines()[-1])
s=d.get("strong",{})
print(json.dumps(dict(n=$G, value=round(d["value"]), us=round(d["ms_per_step"]*1e3,2), unpipelined_us=round(d["detail"]["ms_per_step_unpipelined"]*1e3,2), e2e=round(d["e2e"]["value"]), e2e_us=round(d["e2e"]["ms_per_call"]*1e3,2), scan_us=round(d["roofline"]["launch_ms"]*1e3,2), mhz=d["clocks"]["sm_mhz"], reasons=d["clocks"]["reasons"], parity=d["parity"].get("p2p_vs_oracle", d["parity"].get("single_gpu_vs_oracle")), strong_us=round(s.get("ms_per_step",0)*1e3,2), strong_value=round(s.get("value",0)), strong_unpipelined_us=round(s.get("ms_per_step_unpipelined",0)*1e3,2))))
PY
done | tee gpurun_out/r02g_scale_summary.jsonl
