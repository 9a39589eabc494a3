#!/bin/bash
# ncu evidence for one resident bench step: (1) launch list with per-kernel device time, (2) full capture of the
# rasterize kernels.  Each ncu run is preceded by the same command without ncu (must exit 0).
mkdir -p gpurun_out
CMD="python bench.py --profile-step --steps 1 --warmup 3"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:rasterize_|radix_scatter|radix_hist|project_|sh_colors|isect_|scan_kernel|pack_geom|unpack_geom|rade_loss|offset_encode' \
    -o gpurun_out/prof_raster -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
tail -n 3 gpurun_out/ncu_launches.log; tail -n 3 gpurun_out/ncu_full.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench exit $?; cat gpurun_out/bench.json | head -c 3000
