#!/bin/bash
# e2e host step vs the number of tiles a pre-armed scan prefetches into L2 while it waits for the doorbell
mkdir -p gpurun_out
for rep in 1 2; do
for P in 4 8 12 16 24 40; do
echo "== NGU_EPI_ARM_PREFETCH_TILES=$P"
NGU_EPI_ARM_PREFETCH_TILES=$P timeout 300 python tools/host_profile.py 256x30000 32x30000 64x5000 1x30000 2>&1 | grep -v "OMP\|device us"
done
done | tee gpurun_out/r02_armpf.txt
