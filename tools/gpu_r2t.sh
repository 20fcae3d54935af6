#!/bin/bash
# round 2, call t: lazy reduction in the Miller unit only (mlazy3 = Fp6 products + fp6_mul_by_01, mlazy2 = fp6_mul_by_01 only), against the shipped build:
# pairings at 2^20 (modes 1, 3), 2^16, and the 4-pair checks of BASELINE config 3 at 2^18
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in default mlazy3 mlazy2; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 3 | awk '{printf "%s %s ms | ", $1 $2, $4}') $(python tools/prof_pairing.py 16 3 3 | tail -1 | awk '{printf "2^16 %s ms | ", $4}') $(python tools/prof_checks4.py 18 | awk '{printf "%s %s ms | ", $2, $6}')"
  done
done > gpurun_out/r2t_miller_lazy.log 2>&1
cat gpurun_out/r2t_miller_lazy.log
