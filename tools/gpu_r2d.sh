#!/bin/bash
# round 2, call d: interleaved A/B of the Miller-kernel candidates (noise between identical builds is +-0.5 %)
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in default nofused inpl3 inpl3s2 inpl3s3 sync2; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep"; python tools/prof_pairing.py 20 1 3
  done
done > gpurun_out/r2d_ab.log 2>&1
grep -A2 variant gpurun_out/r2d_ab.log | grep -v "^--" | paste - - - | sort | awk '{print $1, $2, $6, $7, $12, $13}'
