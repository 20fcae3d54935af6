#!/bin/bash
# round 2, call y: the final tree with the precompile-shaped scalar entry points -- full GPU suite, smoke, a short bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['value'], d['e2e']['value'])"
