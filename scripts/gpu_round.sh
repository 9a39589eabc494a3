#!/bin/bash
# tests (isolated groups) + smoke + sort A/B + bench, logs to gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/pytest_gpu.log
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_summary.txt 2>&1
echo "tests exit $?" >> gpurun_out/tests_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/tests_summary.txt
timeout 300 python scripts/sort_ab.py > gpurun_out/sort_ab.log 2>&1; echo "sort_ab exit $?" >> gpurun_out/tests_summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/tests_summary.txt
grep -E "exit|failed" gpurun_out/tests_summary.txt | tail -40
cat gpurun_out/sort_ab.log | tail -6
tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.json | head -c 3000
