#!/bin/bash
# One full default bench line on NG GPUs of one box (config 2 weak scaling + the config-4 block) -> gpurun_out/r2_bench_n${NG}_final.json
NG=${NG:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $NG --steps 30 --warmup 5 > gpurun_out/r2_bench_n${NG}_final.json 2> gpurun_out/r2_bench_n${NG}_final.err
echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n${NG}_final.json').read().strip().splitlines()[-1]); print('N=${NG}', round(d['ms_per_step'],4), round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],4), d['multi_gpu']['replica_gradients_bit_identical'], d['multi_gpu']['main_stream_phases_ms_rank0'], 'cfg4', round(d['config4']['ms_per_step'],3), round(d['config4']['views_per_s'],1), d['config4']['replica_gradients_bit_identical'])"
