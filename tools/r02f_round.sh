#!/bin/bash
# end-of-round 1-GPU call: what the driver runs (tests, smoke, both bench arms with default flags, wall-clocked),
# then the ncu launch list and the full capture of the scan kernel for the same bench command
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
t1=$(date +%s); echo "tests wall $((t1-t0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
t2=$(date +%s); echo "smoke wall $((t2-t1)) s"
timeout 900 python bench.py --impl reference > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"
t3=$(date +%s); echo "reference arm wall $((t3-t2)) s"; cut -c1-600 gpurun_out/r02f_bench_ref.json
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
t4=$(date +%s); echo "bench arm wall $((t4-t3)) s"; cut -c1-1500 gpurun_out/r02f_bench_n1.json; tail -3 gpurun_out/r02f_bench_n1.err
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --no-stream-only"
$CMD > gpurun_out/r02f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
$CMD > gpurun_out/r02f_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1000 -c 3 -f -o gpurun_out/r02f_prof_scan $CMD > gpurun_out/r02f_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02f_ncu_full.log | cut -c1-200
t5=$(date +%s); echo "ncu wall $((t5-t4)) s"
