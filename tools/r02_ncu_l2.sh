#!/bin/bash
# cross-launch L2 residency of the memory stream: ncu WITHOUT cache flushing between passes
mkdir -p gpurun_out
for cfg in "32 30000 0" "32 30000 2,0.5" "32 30000 2,1.0" "32 30000 1" "64 5000 0" "64 5000 1" "16 30000 0" "16 30000 1" "1 30000 0"; do
  set -- $cfg
  tag="${1}x${2}_l2_${3//,/_}"
  NGU_EPI_L2=$3 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sector_op_read_hit_rate.pct -k regex:scan_topk -s 30 -c 3 --csv --log-file gpurun_out/r02_l2res_$tag.csv python tools/ncu_shape.py $1 $2 > /dev/null 2>&1
  echo "$tag rc=$?"
done
