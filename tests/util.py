"""Shared helpers for the tests: limb packing and seeded input generation."""
import json
import os
import random

import numpy as np

import pyref as o

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def fp_arr(vals):
    """list of ints -> flat uint64 limb array (6 per value)."""
    out = []
    for v in vals:
        out.extend(o.fp_to_u64(v))
    return np.array(out, dtype=np.uint64)


def arr_fp(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 6)
    return [o.fp_from_u64([int(x) for x in row]) for row in a]


def fp12_to_arr(f):
    return fp_arr(o.fp12_flatten(f))


def arr_to_fp12(a):
    return o.fp12_unflatten(arr_fp(a))


def g1_to_arr(p):
    return fp_arr([p[0], p[1]])


def g2_to_arr(q):
    return fp_arr([q[0][0], q[0][1], q[1][0], q[1][1]])


def hex_fp12(lst):
    return o.fp12_unflatten([int(h, 16) for h in lst])


def hex_g1(d):
    return (int(d["x"], 16), int(d["y"], 16), bool(d["inf"]))


def hex_g2(d):
    return ((int(d["x"][0], 16), int(d["x"][1], 16)), (int(d["y"][0], 16), int(d["y"][1], 16)), bool(d["inf"]))


def limbs_hex(l):
    """six '0x..' limb strings -> int"""
    return o.fp_from_u64([int(x, 16) for x in l])


EDGE = [0, 1, 2, o.P - 1, o.P - 2, (o.P + 1) // 2, o.R_MONT, (1 << 380), (1 << 32) - 1, 1 << 32, (1 << 64) - 1]


def random_fp_matrix(n, width, seed, edges=True):
    """(n, 6*width) uint64 of canonical field elements; the first rows mix in edge values."""
    rng = random.Random(seed)
    rows = []
    for i in range(n):
        if edges and i < len(EDGE):
            row = [EDGE[(i + j) % len(EDGE)] if (j % 3 != 2) else rng.randrange(o.P) for j in range(width)]
            if i < 3:
                row = [EDGE[i]] * width
        else:
            row = [rng.randrange(o.P) for _ in range(width)]
        rows.append(fp_arr(row))
    return np.stack(rows)


def scalars_for(seed, first, n):
    """The 64-bit scalars zkp_gen_points uses: SplitMix64 random access, zero mapped to one."""
    def at(idx):
        z = (seed + (idx + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    a = [at(2 * (first + i)) or 1 for i in range(n)]
    b = [at(2 * (first + i) + 1) or 1 for i in range(n)]
    return a, b


def scalar_matrix(ks):
    m = np.zeros((len(ks), 4), dtype=np.uint64)
    for i, k in enumerate(ks):
        for j in range(4):
            m[i, j] = (k >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return m


def oracle_points(coracle, seed, first, n):
    """Valid subgroup points a_i*G1, b_i*G2 from the C ORACLE (independent of the CUDA generator)."""
    a, b = scalars_for(seed, first, n)
    g1, i1 = coracle.g1_mul_batch(scalar_matrix(a))
    g2, i2 = coracle.g2_mul_batch(scalar_matrix(b))
    return g1, i1, g2, i2
