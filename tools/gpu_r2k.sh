#!/bin/bash
# round 2, call k: why is the lazy-reduction build not faster?  ncu --set full of one 2^16 step (k_pairing<1> + six k_fe_stage,
# final exponentiation as one piece) for the lazy build and for the same tree with -DZKP_LAZY=0
mkdir -p gpurun_out
export ZKP_PROF_SMALL_WARMUP=1
for v in nosplit nolazyns; do
  export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so
  python tools/prof_pairing.py 16 3 > gpurun_out/r2k_plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"k_pairing|k_fe_stage" --launch-skip 7 --launch-count 3 \
      -o gpurun_out/r2k_$v -f python tools/prof_pairing.py 16 3 > gpurun_out/r2k_ncu_$v.log 2>&1
  bash tools/ncu_export.sh gpurun_out/r2k_$v.ncu-rep 1
done
du -sh gpurun_out
