#!/bin/bash
# one gpurun call: tests, smoke, bench (both arms), ncu launch list + full capture of the scan kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1000 -c 3 -o gpurun_out/prof_scan python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log | cut -c1-300
