#!/bin/bash
# lean multi-GPU check: sort tests + A/B, bench N=1, bench N=$NG
mkdir -p gpurun_out
NG=${NG:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "radix_sort or isect_bit or tile_partitioned or end_to_end" 2>&1 | tail -n 5 > gpurun_out/quick_tests.txt
timeout 300 python scripts/sort_ab.py > gpurun_out/sort_ab.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/quick_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench N=$NG exit $?" >> gpurun_out/quick_tests.txt
cat gpurun_out/quick_tests.txt; tail -n 4 gpurun_out/sort_ab.log
tail -n 3 gpurun_out/bench_n$NG.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench*.json")):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
        if "stage_ms" in d: print("   ", d["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
