#!/bin/bash
# round-2 ncu evidence: launch list of the bench command, full capture of the scan kernel at 256 x 30k (pipelined launches,
# as bench.py times them) and on the small / strong-scaling shapes; each capture only after the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --no-stream-only"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1000 -c 3 -f -o gpurun_out/r02_prof_scan $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_full.log | cut -c1-200
for cfg in "32 30000" "64 5000" "1 30000"; do
  set -- $cfg
  python tools/ncu_shape.py $1 $2 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 30 -c 2 -f -o gpurun_out/r02_prof_scan_${1}x${2} python tools/ncu_shape.py $1 $2 > gpurun_out/r02_ncu_${1}x${2}.log 2>&1
  echo "${1}x${2} rc=$?"
done
ls -la gpurun_out/r02_prof_scan*.ncu-rep
