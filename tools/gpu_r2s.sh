#!/bin/bash
# round 2, call s: the lazy-reduction variants on BASELINE config 3 (4-pair checks, 2^18; plain and prepared) and on the product path
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default lazy2 lazy3 lazy7; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_checks4.py 18 | tr '\n' '|')"
  done
done > gpurun_out/r2s_lazy_checks4.log 2>&1
cat gpurun_out/r2s_lazy_checks4.log
