#!/bin/bash
# round 2, call v (TWO GPUs, the final build with lazy reduction in the Miller unit): the multi-device tests (in-library threads + peer-copy gather) and bench.py under torchrun (NCCL gather)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2v_gpus.txt
python -m pytest tests -m gpu -x -q -k "multi_device or config5 or multi_miller_product or mode_with_free" > gpurun_out/r2v_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest_2gpu.log
tail -6 gpurun_out/r2v_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err; echo "bench rc=$?"
grep -c "NCCL INFO" gpurun_out/r2v_bench_2gpu.err; grep -E "nranks|Connected all|ncclCommInit" gpurun_out/r2v_bench_2gpu.err | head -5
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_2gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value']); print('product', d['product']); print('config5', d['config5']); print('strong', d['strong'])"
tail -c 1500 gpurun_out/r2v_bench_2gpu.err
