#!/bin/bash
# tests + smoke + bench at N=1 and N=2 (torchrun, NCCL)
mkdir -p gpurun_out
rm -f gpurun_out/pytest_gpu.log
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_summary.txt 2>&1
echo "tests exit $?" >> gpurun_out/tests_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/tests_summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/tests_summary.txt
NG=${NG:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench N=$NG exit $?" >> gpurun_out/tests_summary.txt
grep -E "exit|failed" gpurun_out/tests_summary.txt | tail -30
tail -3 gpurun_out/bench.err gpurun_out/bench_n$NG.err
cat gpurun_out/bench.json | head -c 2500; echo; cat gpurun_out/bench_n$NG.json | head -c 1500
