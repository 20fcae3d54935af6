#!/bin/bash
# round 2, call j: lazy reduction (unreduced Fp2 products recombined as 768-bit integers) -- GPU parity suite, then
# interleaved A/B at 2^20 against the same tree built with -DZKP_LAZY=0 (modes: 1 Miller loop, 2 final exponentiation, 3 pairing)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
for rep in 1 2; do
  for v in default nolazy; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep"; python tools/prof_pairing.py 20 1 2 3
  done
done > gpurun_out/r2j_lazy_ab.log 2>&1
cat gpurun_out/r2j_lazy_ab.log
