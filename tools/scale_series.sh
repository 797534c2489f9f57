#!/bin/bash
# N = 1, 2, 4, 8 back to back on one 8-GPU box, launched the way the driver launches them
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
port=29700
for n in 2 4 8; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale_n{n}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(n, "failed", e); continue
    if n == 1: base = d["value"]
    print(json.dumps(dict(n_gpus=n, value=round(d["value"]), ms_per_step=round(d["ms_per_step"], 5), e2e=round(d["e2e"]["value"]),
                          scan_frac=round(d["roofline"]["frac"], 3), eff=round(d["value"] / (n * base), 3) if base else None,
                          sm_mhz=d["clocks"]["sm_mhz"], reasons=d["clocks"]["reasons"])))
PY
