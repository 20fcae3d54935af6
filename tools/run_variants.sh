#!/bin/bash
# Times the pairing kernels (tools/prof_pairing.py) for every build variant under build/ plus the
# in-tree library.  Usage (on the GPU box): tools/run_variants.sh LOG2 [modes...]
log2=${1:-18}; shift
modes=${@:-"1 3"}
echo "variant=default"; python tools/prof_pairing.py $log2 $modes
for so in build/libzkpair_*.so; do
  [ -e "$so" ] || continue
  echo "variant=$so"; ZKPAIR_LIB=$PWD/$so python tools/prof_pairing.py $log2 $modes
done
