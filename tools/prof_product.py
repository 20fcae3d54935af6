#!/usr/bin/env python3
"""BASELINE config 4 (multi-GPU): one product of Miller loops over n = 2^LOG2 pairs sharded in contiguous
slices over all visible GPUs, 576-byte Fp12 partial per GPU gathered on the first one, one final
exponentiation.  Checks that the sharded result is bit-identical to the single-GPU one.
Usage: python tools/prof_product.py [LOG2=20]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log2
ndev = z.device_count()
all_eng = z.PairingEngine()            # every visible device
g1, i1, g2, i2 = all_eng.gen_points(0xFACADE, 0, n)
res = {}
for name, eng in (("%d GPU(s)" % ndev, all_eng), ("1 GPU", z.PairingEngine([0]))):
    eng.multi_miller_product(g1[:4096], g2[:4096])          # warm-up
    for label in ("first full-size call (allocates the pipeline buffers and pinned staging)", "steady state"):
        t0 = time.perf_counter()
        ml, gt = eng.multi_miller_product(g1, g2)
        dt = time.perf_counter() - t0
        print("multi_miller_product n=2^%d on %-9s %.1f ms  %.3f M pairs/s (pageable host buffers, copies included; %s)"
              % (log2, name, dt * 1e3, n / dt / 1e6, label))
    res[name] = (ml, gt)
    if ndev == 1:
        break
vals = list(res.values())
if len(vals) == 2:
    assert np.array_equal(vals[0][0], vals[1][0]) and np.array_equal(vals[0][1], vals[1][1])
    print("sharded product and Gt are bit-identical to the single-GPU result")
