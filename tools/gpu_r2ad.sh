#!/bin/bash
# round 2, call ad: the final tree after the host-side clean-up (event handling on error paths) -- full GPU suite, smoke, a short bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.log
tail -4 gpurun_out/r2ad_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2ad_bench.json')); print(d['value'], d['e2e']['value'])"
