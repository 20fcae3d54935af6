#!/usr/bin/env python3
"""Small pass over every kernel for compute-sanitizer (memcheck): python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import zkvm_pairings_b200 as z

eng = z.PairingEngine([0])
n = 37
g1, i1, g2, i2 = eng.gen_points(3, 0, n)
gt = eng.pairing_batch(g1, g2)
ml = eng.miller_loop_batch(g1, g2)
fe = eng.final_exponentiation_batch(ml)
assert np.array_equal(fe, gt)
gt4, one = eng.multi_pairing_batch(g1[:36], g2[:36], 4)
tab = eng.g2_prepare_batch(g2[:3])
g2v = g2[:36].reshape(9, 4, 24).copy()
g2v[:, 1:, :] = g2[:3]
a, _ = eng.multi_pairing_prepared_batch(g1[:36], np.ascontiguousarray(g2v[:, 0, :]), 4, tab)
b, _ = eng.multi_pairing_batch(g1[:36], g2v.reshape(-1, 24), 4)
assert np.array_equal(a, b)
eng.multi_miller_product(g1, g2)
assert not eng.g1_check_batch(g1).any() and not eng.g2_check_batch(g2).any()
k = np.arange(4 * n, dtype=np.uint64).reshape(n, 4)
eng.g1_mul_batch(g1, k)
eng.g2_mul_batch(g2, k)
limbs, ok = eng.fp_from_bytes_batch(np.arange(48 * 5, dtype=np.uint8).reshape(5, 48))
eng.fp_to_bytes_batch(limbs)
x = eng.tower_op("fp12_mul", gt[:5], gt[5:10])
eng.tower_op("fp12_inv", gt[:3])
eng.tower_op("fp_inv", gt[:3, :6])
eng.close()
print("sanitize_small ok")
