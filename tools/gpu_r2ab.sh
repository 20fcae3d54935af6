#!/bin/bash
# round 2, call ab (EIGHT GPUs): bench.py under torchrun at N = 8 with the final build (weak scaling line + product gather over NCCL + config5 + strong)
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --steps 5 --warmup 3 \
    > gpurun_out/r2ab_bench_8gpu.json 2> gpurun_out/r2ab_bench_8gpu.err; echo "bench rc=$?"
grep -E "nranks|Init COMPLETE" gpurun_out/r2ab_bench_8gpu.err | head -3
python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench_8gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']); print('product', d['product']['value'], d['product']['bit_identical_to_single_gpu_product'], d['product']['ms_per_step']); print('config5', d['config5']['value'], d['config5']['product']); print('strong', d['strong']['value'], d['strong']['ms_per_step'])"
