#!/bin/bash
# config 5 (1024 actors over 8 GPUs + the learner all-reduce buckets) on the final kernel
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29633 bench.py --gpus 8 --config 5 --steps 200 --warmup 20 > gpurun_out/r02g_bench_n8_cfg5.json 2> gpurun_out/r02g_bench_n8_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench_n8_cfg5.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["detail"].get("ms_per_step_unpipelined"), d["e2e"]["value"], d["e2e"].get("ms_per_call"), d["roofline"]["frac"], d["roofline"].get("step_gbs"), d["clocks"], d["parity"].get("p2p_vs_oracle"))
print(json.dumps(d.get("learner_allreduce"))[:1500])
PY
