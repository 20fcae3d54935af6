#!/bin/bash
# round 2, call ai: host pipeline chunk size re-measured with the round-2 kernels (2^17 against the shipped 2^18), e2e at 2^20, interleaved
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default chunk17; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(timeout 120 python tools/prof_e2e.py 20 6 2>&1 | tail -1)"
  done
done > gpurun_out/r2ai_e2e_chunk17.log 2>&1
cat gpurun_out/r2ai_e2e_chunk17.log
