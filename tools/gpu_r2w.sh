#!/bin/bash
# round 2, call w: small variants on the shipped build, interleaved A/B at 2^20 (modes 1, 2, 3) and 2^16 --
# msplit (two-stream halves start at the Miller kernel), sqrsel (fp2_sqr: select instead of a broadcast shuffle),
# carve0 / carve50 (explicit L1 / shared split for the stage kernels), xopt (ptxas expensive optimizations),
# msync3 / msync5 (rendezvous density of the Miller unit on top of the lazy forms), both = msplit + sqrsel;
# then the parity tests that cover the split path with the candidates
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default msplit sqrsel both carve0 carve50 xopt msync3 msync5; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(python tools/prof_pairing.py 20 1 2 3 | awk '{printf "%s %s ms | ", $1 $2, $4}') $(python tools/prof_pairing.py 16 3 3 | tail -1 | awk '{printf "2^16 %s ms | ", $4}')"
  done
done > gpurun_out/r2w_variants.log 2>&1
cat gpurun_out/r2w_variants.log
for v in both; do
  ZKPAIR_LIB=$PWD/build/libzkpair_$v.so python -m pytest tests -m gpu -x -q -k "config1 or 2p20 or config3 or config2 or golden or seeded or device_resident or prepared or final_exponentiation_edge" > gpurun_out/r2w_pytest_$v.log 2>&1
  echo "pytest($v) rc=$?" >> gpurun_out/r2w_pytest_$v.log; tail -3 gpurun_out/r2w_pytest_$v.log
done
