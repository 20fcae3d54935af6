#!/bin/bash
# round 2, first GPU call: parity suite, default bench, launch list, ncu captures (one step at 2^16 without the FE split; k_pairing<4>)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
python tools/prof_configs.py 18 > gpurun_out/r2a_configs.log 2>&1; cat gpurun_out/r2a_configs.log
python tools/prof_product.py > gpurun_out/r2a_product.log 2>&1; tail -5 gpurun_out/r2a_product.log
# launch list of the bench command (cold-cache, serialised: shares only)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2a_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2a_ncu_bench.log 2>&1
export ZKP_PROF_SMALL_WARMUP=1
ZKPAIR_LIB=$PWD/build/libzkpair_nosplit.so python tools/prof_pairing.py 16 3 > gpurun_out/r2a_plain_step.log 2>&1 &&
ZKPAIR_LIB=$PWD/build/libzkpair_nosplit.so ncu --set full --clock-control none --import-source on \
    -k regex:"k_pairing|k_fe_stage|k_fe_batch_inv" --launch-skip 13 --launch-count 13 -o gpurun_out/r2a_step -f \
    python tools/prof_pairing.py 16 3 > gpurun_out/r2a_ncu_step.log 2>&1
python tools/prof_checks4.py 16 > gpurun_out/r2a_plain_checks4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairing" --launch-skip 2 --launch-count 2 -o gpurun_out/r2a_checks4 -f \
    python tools/prof_checks4.py 16 > gpurun_out/r2a_ncu_checks4.log 2>&1
cat gpurun_out/r2a_plain_checks4.log
ls -la gpurun_out/*.ncu-rep | tail -3
