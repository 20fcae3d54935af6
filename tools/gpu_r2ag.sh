#!/bin/bash
# round 2, call ag (TWO GPUs, final tree): the multi-device tests (in-library threads + peer-copy gather, config 5 at 2^24) on the final device code
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2ag_gpus.txt
timeout 200 python -m pytest tests -m gpu -x -q -k "multi_device or config5" > gpurun_out/r2ag_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ag_pytest_2gpu.log
tail -4 gpurun_out/r2ag_pytest_2gpu.log
