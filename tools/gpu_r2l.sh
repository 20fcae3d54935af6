#!/bin/bash
# round 2, call l: which lazy-reduction pieces pay?  ZKP_LAZY bit 0 = Fp6 products, bit 1 = fp6_mul_by_01, bit 2 = Fp4 squares.
# Interleaved A/B at 2^20 (modes: 1 Miller loop, 2 final exponentiation, 3 pairing)
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default nolazy lazy1 lazy2 lazy3 lazy4; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 2 3 | awk '{printf "%s %s ms | ", $1 $2, $4}')"
  done
done > gpurun_out/r2l_lazy_pieces.log 2>&1
cat gpurun_out/r2l_lazy_pieces.log
