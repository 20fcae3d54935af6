#!/bin/bash
# round 2, call m: fp6_mul_sums through a thread-local operand sum + fp6_mul (ZKP_SUMS_SIMPLE=1: 26 KB less hot code, 21 Fp2
# additions and 4 xi-multiplications fewer per Fp12 squaring) against the shipped on-the-fly form; lazy variants for reference.
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in nolazy sums lazy3; do
    export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 3 | awk '{printf "%s %s ms | ", $1 $2, $4}')"
  done
done > gpurun_out/r2m_sums.log 2>&1
cat gpurun_out/r2m_sums.log
