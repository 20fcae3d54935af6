#!/bin/bash
# round 2, third GPU call: parity suite on the new default (binary-GCD inversion, f and R in shared memory), variants, small batches
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
bash tools/run_variants.sh 20 1 2 3 > gpurun_out/r2c_variants_2p20.log 2>&1; cat gpurun_out/r2c_variants_2p20.log
for l in 14 16 17; do echo "== 2^$l"; python tools/prof_pairing.py $l 3 3 3; done > gpurun_out/r2c_small.log 2>&1; cat gpurun_out/r2c_small.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['peak'])
print({k:v['value'] for k,v in d['configs'].items()}, d['product']['value'])"
./build/dfma_probe > gpurun_out/r2c_dfma_probe.txt 2>&1; cat gpurun_out/r2c_dfma_probe.txt
