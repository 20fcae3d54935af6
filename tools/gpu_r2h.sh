#!/bin/bash
# round 2, call h: final-exponentiation variants, interleaved A/B at 2^20 (mode 2 = final exponentiation only)
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in default csqinl fesmem; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep"; python tools/prof_pairing.py 20 2
  done
done > gpurun_out/r2h_fe_ab.log 2>&1
grep -A1 variant gpurun_out/r2h_fe_ab.log | grep -v "^--" | paste - - | sort
