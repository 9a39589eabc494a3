#!/bin/bash
# N-GPU quick A/B of the push CTAs (config 2 only), then one full default bench line
NG=${NG:-8}
mkdir -p gpurun_out
for f in "--push-ctas 4" "--push-ctas 6" "--push-ctas 8" "--push-ctas 4" "--push-engine dma"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $NG --steps 30 --warmup 5 --no-extras $f 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['multi_gpu']['exchange_spans_ms_rank0'], d['multi_gpu']['compute_only_ms_per_rank'][:3])"
done
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r2_bench_n${NG}_final.json 2> gpurun_out/r2_bench_n${NG}_final.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n${NG}_final.json').read().strip().splitlines()[-1]); print('full', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['multi_gpu']['replica_gradients_bit_identical'], d['config4']['ms_per_step'], d['config4']['views_per_s'], d['config4']['replica_gradients_bit_identical'])"
