#!/bin/bash
# Runs the GPU parity tests group by group, each group in its own process so that one CUDA fault
# cannot poison the others.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
for grp in projection spherical cumsum radix_sort isect_bit rasterize_to_pixels rasterize_absgrad end_to_end retain_grad unsupported edge_cases full_size fused_loss sh_colors golden adjoint multi_camera tile_partitioned; do
  echo "=== $grp" | tee -a gpurun_out/pytest_gpu.log
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -p no:cacheprovider -k "$grp" 2>&1 | tail -n 60 >> gpurun_out/pytest_gpu.log
  rc=${PIPESTATUS[0]}
  echo "--- $grp exit $rc" | tee -a gpurun_out/pytest_gpu.log
  if [ $rc -ne 0 ]; then status=1; fi
done
grep -E "^(===|---)|passed|failed|error" gpurun_out/pytest_gpu.log | tail -n 60
exit $status
