#!/bin/bash
# one 1-GPU gpurun call: tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_ref.json; tail -3 gpurun_out/r02_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
