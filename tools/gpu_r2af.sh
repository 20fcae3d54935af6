#!/bin/bash
# round 2, call af: the final tree -- driver-shaped bench (all sub-records), reference arm, launch list of the bench
set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2af_bench_reference_arm.json 2> gpurun_out/r2af_bench_reference_arm.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2af_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2af_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2af_ncu_bench.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/r2af_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['peak'])
print({k:v['value'] for k,v in d['configs'].items()}, d['product']['value'], d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/r2af_bench_reference_arm.json')); print('reference arm', r['value'], r['cpu_baseline'])"
