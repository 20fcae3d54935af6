#!/bin/bash
# round 2, call z: lorder (lazy fp6_mul with fewer unreduced values alive across the last calls), msync5 (rendezvous per Fp2-level body
# in the Miller unit), lorder5 = both; interleaved A/B against the shipped build, 2^20 modes 1 and 3, three repetitions
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in default lorder msync5 lorder5; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 3 | awk '{printf "%s %s ms | ", $1 $2, $4}') $(python tools/prof_pairing.py 16 3 3 | tail -1 | awk '{printf "2^16 %s ms | ", $4}')"
  done
done > gpurun_out/r2z_variants.log 2>&1
cat gpurun_out/r2z_variants.log
