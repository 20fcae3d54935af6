#!/bin/bash
# round 2, call ah: early per-half device->host copy of the results (pinned destinations): e2e A/B at 2^20 (ZKP_EARLY_D2H=0 = after both
# halves, as before), three interleaved repetitions; then the GPU parity suite and a short bench (e2e must match the device path)
mkdir -p gpurun_out
for rep in 1 2 3; do
  for t in 0 1; do
    echo "early_d2h=$t rep=$rep $(ZKP_EARLY_D2H=$t timeout 120 python tools/prof_e2e.py 20 6 2>&1 | tail -1)"
  done
done > gpurun_out/r2ah_e2e_early_d2h.log 2>&1
cat gpurun_out/r2ah_e2e_early_d2h.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
tail -3 gpurun_out/r2ah_pytest.log
python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2ah_bench.json')); print(d['value'], d['e2e'])"
