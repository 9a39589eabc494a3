#!/bin/bash
# ncu --set full of the compositing kernels inside one resident bench step; RADE_RASTER_FLAGS selects the variant.
# usage: scripts/gpu_ncu_bwd.sh <tag>
TAG=${1:-default}
mkdir -p gpurun_out
CMD="python bench.py --profile-step --steps 1 --warmup 3"
$CMD > gpurun_out/r2_ncu_${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:rasterize_' \
    -o gpurun_out/r2_prof_${TAG} -f $CMD > gpurun_out/r2_ncu_${TAG}.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/r2_ncu_${TAG}.log
