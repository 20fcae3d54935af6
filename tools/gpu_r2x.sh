#!/bin/bash
# round 2, call x: the FINAL tree (r2u device arithmetic; k_pairing takes a start index, host-side split hooks off by default) --
# full parity suite, driver-shaped bench, reference arm, launch list
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -4 gpurun_out/r2x_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference_arm.json 2> gpurun_out/r2x_bench_reference_arm.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2x_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2x_ncu_bench.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['peak'])
print({k:v['value'] for k,v in d['configs'].items()}, d['product']['value'], d['cpu_baseline']['value'])"
