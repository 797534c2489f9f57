#!/bin/bash
# eight independent single-GPU benches side by side: per-GPU speed spread without any exchange
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/indep_$i.json 2> gpurun_out/indep_$i.err &
done
wait
python - <<'PY'
import json, glob
rows = []
for p in sorted(glob.glob("gpurun_out/indep_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        rows.append((p[-6], round(d["ms_per_step"] * 1e3, 2), round(d["roofline"]["launch_ms"] * 1e3, 2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    except Exception as e:
        rows.append((p, "failed", str(e)))
print(json.dumps({"independent_per_gpu": rows}))
PY
