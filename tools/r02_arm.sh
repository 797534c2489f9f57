#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -8
for arm in 0 1; do
echo "== NGU_EPI_HOST_ARM=$arm"
NGU_EPI_HOST_ARM=$arm timeout 300 python tools/host_profile.py 256x30000 64x5000 1x30000 32x30000 2>&1 | grep -v OMP | tee gpurun_out/r02_host_profile_arm$arm.txt
done
