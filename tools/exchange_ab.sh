#!/bin/bash
# 2-GPU step time with the p2p exchange, with the exchange switched off (ranks independent) and with the NCCL collective
G=${1:-2}; port=29720
for x in auto off collective auto off; do
  port=$((port+1))
  NGU_BENCH_EXCHANGE=$x timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $port bench.py --gpus $G --steps 200 --warmup 20 2>/dev/null | grep "^{" > gpurun_out/xab.json
  python - "$x" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/xab.json").read())
    print(sys.argv[1], round(d["ms_per_step"] * 1e3, 2), "us/step, value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "|", d["config"]["exchange"][:45])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
