#!/bin/bash
# round 2, call ae: end-to-end (pinned host buffers) chunk schedules -- uniform 2^18 chunks (shipped) against a tapered schedule
# (chunk/4 first; chunk/2, chunk/4 last), each also with 2^19 chunks; tools/prof_e2e.py at 2^20, interleaved, two repetitions
mkdir -p gpurun_out
rm -f build/libzkpair_mb3.so build/libzkpair_smem1.so build/libzkpair_smem2.so
for rep in 1 2; do
  for v in default chunk19; do
    if [ $v = default ]; then unset ZKPAIR_LIB; else export ZKPAIR_LIB=$PWD/build/libzkpair_$v.so; fi
    for t in 0 1; do
      echo "variant=$v rep=$rep $(ZKP_TAPER=$t timeout 120 python tools/prof_e2e.py 20 4 2>&1 | tail -1)"
    done
  done
done > gpurun_out/r2ae_e2e_schedules.log 2>&1
cat gpurun_out/r2ae_e2e_schedules.log
