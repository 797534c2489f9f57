#!/bin/bash
# ncu --set full captures of the scan kernel on the small / strong-scaling shapes (each after a plain run exited 0)
mkdir -p gpurun_out
for cfg in "32 30000 0" "32 30000 2,0.5" "32 30000 1" "64 5000 0" "1 30000 0" "16 30000 1"; do
  set -- $cfg
  tag="${1}x${2}_l2_${3//,/_}"
  NGU_EPI_L2=$3 python tools/ncu_shape.py $1 $2 > /dev/null 2>&1 && \
  NGU_EPI_L2=$3 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 30 -c 2 -f -o gpurun_out/r02_scan_$tag python tools/ncu_shape.py $1 $2 > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag rc=$?"
done
ls -la gpurun_out/*.ncu-rep
