#!/bin/bash
# round 2, call i: multiply-pipe token probe (tools/ring_probe.cu), one process per configuration under its own timeout
mkdir -p gpurun_out
for cfg in "4 128 0" "4 512 0" "4 512 1" "3 128 0" "3 384 0" "3 384 1" "2 128 0" "2 256 0" "2 256 1"; do
  timeout 40 build/ring_probe $cfg || echo "cfg $cfg: rc=$?"
done > gpurun_out/r2i_ring_probe.txt 2>&1
cat gpurun_out/r2i_ring_probe.txt
