#!/bin/bash
# end-of-round 1-GPU call on the lean scan kernel: full GPU suite, a 600-seed fuzz soak, smoke, both bench arms with
# default flags, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
NGU_FUZZ_SEEDS=600 timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/r02g_bench_n1.json; tail -2 gpurun_out/r02g_bench_n1.err | cut -c1-300
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --no-stream-only"
$CMD > gpurun_out/r02g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/r02g_launches.csv $CMD > gpurun_out/r02g_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
