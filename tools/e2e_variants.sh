#!/bin/bash
show() { python -c "
import json,sys
d=json.load(sys.stdin)
print(d['config']['engine'][-16:], round(d['value']), round(d['e2e']['value']))"; }
python bench.py --steps 3 --warmup 3 2>/dev/null | show
for so in build/libzkpair_*.so; do ZKPAIR_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 2>/dev/null | show; done
