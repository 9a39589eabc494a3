#!/bin/bash
# tests (isolated groups) + smoke + bench, logs to gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/pytest_gpu.log
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_summary.txt 2>&1
echo "tests exit $?" >> gpurun_out/tests_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/tests_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/tests_summary.txt
cat gpurun_out/tests_summary.txt | tail -50
tail -5 gpurun_out/smoke.log
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
