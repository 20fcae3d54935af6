#!/usr/bin/env python3
"""Drives the bound-tracking build of the CPU dev simulation (libzkpair_sim_tb.so) through every
code path: all tower ops on edge + random operands, single / multi-pair Miller loops, final
exponentiation, pairing, point generation.  The tracker aborts the process on a violation."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import coracle  # noqa: E402
import util  # noqa: E402
from zkvm_pairings_b200.engine import TOWER_OPS, op_widths  # noqa: E402

sim = ctypes.CDLL(sys.argv[1])


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


for name, code in TOWER_OPS.items():
    na, nb, nr = op_widths(name)
    n = 14
    a = util.random_fp_matrix(n, na, seed=code + 1)
    b = util.random_fp_matrix(n, nb, seed=code + 101) if nb else None
    out = np.zeros((n, 6 * nr), np.uint64)
    sim.sim_tower_op(code, P(a), P(b), P(out), None, ctypes.c_size_t(n))
    if name in ("fp12_frob2", "fp12_frob3"):
        exp = a
        for _ in range(int(name[-1])):
            exp = coracle.tower_op("fp12_frob", exp)
    else:
        exp = coracle.tower_op(name, a, b)
    assert np.array_equal(out, exp), name
for k in (1, 2, 3, 4, 8):
    nchk = 2
    g1, i1, g2, i2 = util.oracle_points(coracle, 50 + k, 0, nchk * k)
    if k > 1:
        i1[1] = 1
    out, one = np.zeros((nchk, 72), np.uint64), np.zeros(nchk, np.uint8)
    assert sim.sim_pairing(3, P(g1), P(i1), P(g2), P(i2), ctypes.c_size_t(nchk), k, None, P(out), P(one)) == 0
    exp, eone = coracle.multi_pairing_batch(g1, i1, g2, i2, k)
    assert np.array_equal(out, exp) and np.array_equal(one, eone), k
    ml = np.zeros((nchk, 72), np.uint64)
    assert sim.sim_pairing(1, P(g1), P(i1), P(g2), P(i2), ctypes.c_size_t(nchk), k, None, P(ml), None) == 0
    fe = np.zeros((nchk, 72), np.uint64)
    assert sim.sim_pairing(2, None, None, None, None, ctypes.c_size_t(nchk), 1, P(ml), P(fe), None) == 0
    assert np.array_equal(fe, exp)
n = 4
a, b = util.scalars_for(9, 0, n)
k1, k2 = np.array(a, np.uint64), np.array(b, np.uint64)
g1, g2 = np.zeros((n, 12), np.uint64), np.zeros((n, 24), np.uint64)
f1, f2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
sim.sim_gen_points(P(k1), P(k2), ctypes.c_size_t(n), P(g1), P(f1), P(g2), P(f2))
e1, _, e2, _ = util.oracle_points(coracle, 9, 0, n)
assert np.array_equal(g1, e1) and np.array_equal(g2, e2)
mb = (ctypes.c_double * 4)()
sim.sim_max_bounds(mb)
print("BOUNDS OK  max|limb| = 2^%.2f  max|top limb| = 2^%.2f  max|value| = %.1f p  max column = 2^%.2f"
      % (np.log2(mb[0]), np.log2(mb[1]), mb[2], np.log2(mb[3])))
