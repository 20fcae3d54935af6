#!/bin/bash
# round 2, call o: rendezvous density / branch-free subtraction / one-piece final exponentiation re-measured on the smaller code
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default msync3 msync5 subnb fsync3 fsync5 nosplit; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 2 3 | awk '{printf "%s %s ms | ", $1 $2, $4}')"
  done
done > gpurun_out/r2o_variants.log 2>&1
cat gpurun_out/r2o_variants.log
