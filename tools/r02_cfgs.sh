#!/bin/bash
# bench.py on the other SURVEY 8d configurations (1 GPU) + the default invocation, timed
mkdir -p gpurun_out
for c in 1 2; do
  t0=$(date +%s.%N)
  timeout 600 python bench.py --config $c --steps 200 --warmup 20 > gpurun_out/r02_bench_cfg$c.json 2> gpurun_out/r02_bench_cfg$c.err; echo "cfg$c rc=$? wall $(echo "$(date +%s.%N) - $t0" | bc) s"; cut -c1-500 gpurun_out/r02_bench_cfg$c.json
done
t0=$(date +%s.%N)
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$? wall $(echo "$(date +%s.%N) - $t0" | bc) s"; cut -c1-300 gpurun_out/r02_bench_default.json
t0=$(date +%s.%N)
timeout 600 python bench.py --scaling strong --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_strong1.json 2> gpurun_out/r02_bench_strong1.err; echo "strong rc=$? wall $(echo "$(date +%s.%N) - $t0" | bc) s"
