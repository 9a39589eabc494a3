#!/bin/bash
# Round-2 evidence run on one B200, part A: GPU tests (with the gradient-parity report), smoke, bench N=1 (both arms) and
# the ncu launch list of one resident step.  Part B (scripts/gpu_round2_ncu.sh) is the full ncu capture: one profiler
# tool per gpurun call.  Logs to gpurun_out/r2_*.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_full.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2_pytest_full.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "reference arm exit $?"
CMD="python bench.py --profile-step --steps 1 --warmup 3"
$CMD > gpurun_out/r2_profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches exit $?"
head -c 1500 gpurun_out/r2_bench_n1.json; echo; head -c 1200 gpurun_out/r2_bench_ref.json
