#!/bin/bash
# bench at N=$NG with the three SH-gradient exchange schemes (p2p gather kernel, all-gather + kernel, all-reduce)
mkdir -p gpurun_out
NG=${NG:-2}
for mode in ${MODES:-p2p allgather allreduce}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 --grad-exchange $mode > gpurun_out/bench_n${NG}_$mode.json 2> gpurun_out/bench_n${NG}_$mode.err; echo "bench N=$NG $mode exit $?"
  tail -n 2 gpurun_out/bench_n${NG}_$mode.err | cut -c1-300
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_n*_*.json")):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(f, "ERR", e)
PY
