#!/bin/bash
# round 2, call ac: (1) compute-sanitizer memcheck / racecheck / synccheck of the shipped build over every kernel (tools/sanitize_small.py:
# the Miller kernel now keeps f + the in-place temporary in shared memory and meets at block-wide rendezvous points);
# (2) occupancy / shared-memory-state flags re-measured ON TOP of the lazy forms of the Miller unit (they were last measured with the
# reduced forms): mb3 = 3 blocks per SM at 168 registers, smem1 = only f in shared memory (more L1 for the unreduced values' spills),
# smem2 = f + R; interleaved A/B against the shipped build, 2^20 modes 1 and 3, two repetitions
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2ac_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok' gpurun_out/r2ac_sanitizer_$tool.log | tr '\n' ' ')"
done
for rep in 1 2; do
  for v in default mb3 smem1 smem2; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    echo "variant=$v rep=$rep $(timeout 120 python tools/prof_pairing.py 20 1 3 | awk '{printf "%s %s ms | ", $1 $2, $4}')"
  done
done > gpurun_out/r2ac_variants.log 2>&1
cat gpurun_out/r2ac_variants.log
