#!/bin/bash
# Round-2 evidence run, part B: ncu --set full of every library kernel of one resident step (after the same command has
# exited 0 without the profiler).
mkdir -p gpurun_out
CMD="python bench.py --profile-step --steps 1 --warmup 3"
$CMD > gpurun_out/r2_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k 'regex:rasterize_|radix_scatter|radix_hist|project_|sh_colors|isect_|scan_kernel|pack_geom|unpack_geom|rade_loss|offset_encode' \
    -o gpurun_out/r2_prof_step -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full exit $?"
