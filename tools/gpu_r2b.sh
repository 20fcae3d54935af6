#!/bin/bash
# round 2, second GPU call: shared-memory state variants, bench line to a file, ncu summaries exported on the box
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
bash tools/run_variants.sh 20 1 3 > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
python tools/prof_product.py > gpurun_out/r2b_product.log 2>&1; tail -4 gpurun_out/r2b_product.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2b_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2b_ncu_bench.log 2>&1
export ZKP_PROF_SMALL_WARMUP=1
for v in nosplit smem2ns; do
ZKPAIR_LIB=$PWD/build/libzkpair_$v.so python tools/prof_pairing.py 16 3 > gpurun_out/r2b_plain_step_$v.log 2>&1 &&
ZKPAIR_LIB=$PWD/build/libzkpair_$v.so ncu --set full --clock-control none --import-source on \
    -k regex:"k_pairing|k_fe_stage" --launch-skip 7 --launch-count 7 -o gpurun_out/r2b_step_$v -f \
    python tools/prof_pairing.py 16 3 > gpurun_out/r2b_ncu_step_$v.log 2>&1
bash tools/ncu_export.sh gpurun_out/r2b_step_$v.ncu-rep 1
done
python tools/prof_checks4.py 16 > gpurun_out/r2b_plain_checks4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairing" --launch-skip 2 --launch-count 2 -o gpurun_out/r2b_checks4 -f \
    python tools/prof_checks4.py 16 > gpurun_out/r2b_ncu_checks4.log 2>&1
bash tools/ncu_export.sh gpurun_out/r2b_checks4.ncu-rep 1
du -sh gpurun_out
