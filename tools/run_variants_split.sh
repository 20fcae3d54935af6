#!/bin/bash
# Miller-kernel variants (build/libzkpair_mb*.so) are timed in mode 1, final-exponentiation variants (feb*) in mode 2.
log2=${1:-20}
echo "variant=default"; python tools/prof_pairing.py $log2 1 2
for so in build/libzkpair_mb*.so; do [ -e "$so" ] || continue; echo "variant=$so"; ZKPAIR_LIB=$PWD/$so python tools/prof_pairing.py $log2 1; done
for so in build/libzkpair_feb*.so; do [ -e "$so" ] || continue; echo "variant=$so"; ZKPAIR_LIB=$PWD/$so python tools/prof_pairing.py $log2 2; done
