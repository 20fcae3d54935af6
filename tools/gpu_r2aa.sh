#!/bin/bash
# round 2, call aa: the shipped build (lazy reduction + Fp2-level rendezvous points + reordered lazy sums in the Miller unit) -- full parity suite, driver-shaped bench, reference arm, launch list, ncu captures
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
tail -4 gpurun_out/r2aa_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2aa_bench_reference_arm.json 2> gpurun_out/r2aa_bench_reference_arm.err; echo "ref rc=$?"
python tools/prof_product.py > gpurun_out/r2aa_product.log 2>&1; tail -3 gpurun_out/r2aa_product.log
for l in 16 17; do python tools/prof_pairing.py $l 3 3; done > gpurun_out/r2aa_small.log 2>&1; cat gpurun_out/r2aa_small.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2aa_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2aa_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2aa_ncu_bench.log 2>&1
export ZKP_PROF_SMALL_WARMUP=1
ZKPAIR_LIB=$PWD/build/libzkpair_nosplit.so python tools/prof_pairing.py 16 3 > gpurun_out/r2aa_plain_step.log 2>&1 &&
ZKPAIR_LIB=$PWD/build/libzkpair_nosplit.so ncu --set full --clock-control none --import-source on \
    -k regex:"k_pairing|k_fe_stage|k_fe_batch_inv" --launch-skip 13 --launch-count 13 -o gpurun_out/r2aa_step -f \
    python tools/prof_pairing.py 16 3 > gpurun_out/r2aa_ncu_step.log 2>&1
bash tools/ncu_export.sh gpurun_out/r2aa_step.ncu-rep 1
python tools/prof_checks4.py 16 > gpurun_out/r2aa_plain_checks4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairing" --launch-skip 2 --launch-count 2 -o gpurun_out/r2aa_checks4 -f \
    python tools/prof_checks4.py 16 > gpurun_out/r2aa_ncu_checks4.log 2>&1
bash tools/ncu_export.sh gpurun_out/r2aa_checks4.ncu-rep 1
python -c "
import json; d=json.load(open('gpurun_out/r2aa_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['peak'])
print({k:v['value'] for k,v in d['configs'].items()}, d['product']['value'], d['cpu_baseline']['value'])"
du -sh gpurun_out
