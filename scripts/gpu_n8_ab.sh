#!/bin/bash
# N-GPU A/B of the gradient-exchange options on the full bench line (config 2 weak scaling + config 4 strong scaling)
NG=${NG:-8}
mkdir -p gpurun_out
run() {  # tag, extra args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $NG --steps 20 --warmup 3 $2 > gpurun_out/r2_bench_n${NG}_$1.json 2> gpurun_out/r2_bench_n${NG}_$1.err
  echo "$1 rc $?"
}
for v in ${VARIANTS:-nccl_dma peer_dma peer_sm}; do
  case $v in
    nccl_dma) run $v "--small-allreduce nccl --push-engine dma" ;;
    peer_dma) run $v "--small-allreduce peer --push-engine dma" ;;
    peer_sm) run $v "--small-allreduce peer --push-engine sm --push-ctas 8" ;;
    nccl_sm4) run $v "--small-allreduce nccl --push-engine sm --push-ctas 4" ;;
    nccl_sm8) run $v "--small-allreduce nccl --push-engine sm --push-ctas 8" ;;
    nccl_sm16) run $v "--small-allreduce nccl --push-engine sm --push-ctas 16" ;;
  esac
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_n*_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        c4 = d.get("config4", {})
        print(f, "cfg2 ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "| cfg4 ms/step", round(c4.get("ms_per_step", 0), 3),
              "views/s", round(c4.get("views_per_s", 0), 1))
        print("   cfg2 spans", d["multi_gpu"]["exchange_spans_ms_rank0"], "compute", d["multi_gpu"]["compute_only_ms_per_rank"])
        print("   cfg4 spans", c4["multi_gpu"]["exchange_spans_ms_rank0"], "compute", c4["multi_gpu"]["compute_only_ms_per_rank"])
    except Exception as e:
        print(f, "ERR", e)
PY
