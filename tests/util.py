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


def group_check_cases(pyref, coracle, n_valid=4, seed=0x6F):
    """Points for the validity checks with the status the reference semantics give them
    (0 ok, 1 not on curve, 2 on the curve but outside the prime-order subgroup): valid multiples of
    the generators, the identity, random off-curve pairs (what the reference's random() produces,
    src/g1.rs:64-72), random curve points (cofactor != 1, so outside the subgroup) and the
    reference's own G2 torsion KAT point (src/g2.rs:401-443)."""
    import random
    rng = random.Random(seed)
    P = pyref.P
    g1v, _, g2v, _ = oracle_points(coracle, seed, 0, n_valid)
    g1, i1, e1 = [r for r in g1v], [0] * n_valid, [0] * n_valid
    g2, i2, e2 = [r for r in g2v], [0] * n_valid, [0] * n_valid
    g1.append(g1_to_arr((0, 1))); i1.append(1); e1.append(0)
    g2.append(g2_to_arr(((0, 0), (1, 0)))); i2.append(1); e2.append(0)
    for _ in range(3):
        g1.append(g1_to_arr((rng.randrange(P), rng.randrange(P)))); i1.append(0); e1.append(1)
        g2.append(g2_to_arr(((rng.randrange(P), rng.randrange(P)), (rng.randrange(P), rng.randrange(P))))); i2.append(0); e2.append(1)
    cnt = 0
    while cnt < 3:
        x = rng.randrange(P)
        y = pyref.fp_sqrt(pyref.fp_add(pyref.fp_mul(pyref.fp_square(x), x), pyref.B1))
        if y is None:
            continue
        assert pyref.g1_is_on_curve((x, y, False)) and not pyref.g1_is_torsion_free((x, y, False))
        g1.append(g1_to_arr((x, y))); i1.append(0); e1.append(2); cnt += 1
    cnt = 0
    while cnt < 3:
        x = (rng.randrange(P), rng.randrange(P))
        y = pyref.fp2_sqrt(pyref.fp2_add(pyref.fp2_mul(pyref.fp2_square(x), x), pyref.B2))
        if y is None:
            continue
        assert pyref.g2_is_on_curve((x, y, False)) and not pyref.g2_is_torsion_free((x, y, False))
        g2.append(g2_to_arr((x, y))); i2.append(0); e2.append(2); cnt += 1
    # the reference's torsion KAT (src/g2.rs:401-443) holds zkcrypto's MONTGOMERY limbs, which in this
    # crate's canonical storage are not a curve point at all: is_valid fails at the on-curve test
    kat = golden("reference_kats.json")["g2_not_torsion_free"]["p"]     # x.c0, x.c1, y.c0, y.c1 as u64 limbs
    kpt = [sum(int(h, 16) << (64 * i) for i, h in enumerate(limbs)) for limbs in kat]
    assert not pyref.g2_is_on_curve(((kpt[0], kpt[1]), (kpt[2], kpt[3]), False))
    assert not pyref.g2_is_torsion_free(((kpt[0], kpt[1]), (kpt[2], kpt[3]), False))
    g2.append(np.array([int(h, 16) for limbs in kat for h in limbs], dtype=np.uint64)); i2.append(0); e2.append(1)
    return (np.stack(g1), np.array(i1, np.uint8), np.array(e1, np.uint8),
            np.stack(g2), np.array(i2, np.uint8), np.array(e2, np.uint8))


def random_scalars(n, seed, edges=True):
    """(n,4) little-endian u64 limbs of 256-bit scalars (edge cases first: 0, 1, 2, r-1, r, 2^256-1)."""
    import random
    rng = random.Random(seed)
    R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    vals = [0, 1, 2, 3, R_ORDER - 1, R_ORDER, (1 << 256) - 1] if edges else []
    vals = vals[:n] + [rng.getrandbits(256) for _ in range(max(0, n - len(vals)))]
    return np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals[:n]], dtype=np.uint64)


POW_OPS = ("fp_pow", "fp2_pow", "fp12_pow", "fp_sqrt")


def pow_sqrt_case(pyref, name, n, seed):
    """Operands and expected results (via the Python oracle) for the ops the C oracle has no entry for:
    pow_vartime (src/fp.rs:264-276, src/fp2.rs:301-313, src/fp12.rs:127-139; exponent = six raw u64 limbs,
    edge exponents 0, 1, 2, p, 2^384-1 first) and Fp::sqrt (src/fp.rs:280-300; the reference's KAT
    sqrt(300855555557) and its non-residue 72057594037927816 first)."""
    import random
    rng = random.Random(seed)
    P = pyref.P
    if name == "fp_sqrt":
        vals = [300855555557, 72057594037927816, 0, 1, 4] + [rng.randrange(P) for _ in range(max(0, n - 5))]
        vals = vals[:n]
        a = fp_arr(vals).reshape(-1, 6)
        roots = [pyref.fp_sqrt(v) for v in vals]
        exp_status = np.array([2 if r is None else 0 for r in roots], np.uint8)
        return a, None, roots, exp_status
    width = {"fp_pow": 1, "fp2_pow": 2, "fp12_pow": 12}[name]
    a = random_fp_matrix(n, width, seed=seed, edges=False)
    es = [0, 1, 2, P, (1 << 384) - 1] + [rng.getrandbits(384) for _ in range(max(0, n - 5))]
    es = es[:n]
    b = np.array([[(e >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)] for e in es], dtype=np.uint64)
    exp = []
    for row, e in zip(a, es):
        by = [(e >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)]
        if name == "fp_pow":
            exp.append(fp_arr([pyref.fp_pow_vartime(arr_fp(row)[0], by)]).reshape(-1))
        elif name == "fp2_pow":
            v = arr_fp(row)
            r = pyref.fp2_pow_vartime((v[0], v[1]), by)
            exp.append(fp_arr([r[0], r[1]]).reshape(-1))
        else:
            exp.append(fp12_to_arr(pyref.fp12_pow_vartime(arr_to_fp12(row), by)))
    return a, b, np.stack(exp), np.zeros(n, np.uint8)


def final_exp_edge_inputs(seed=11, n_random=6):
    """Fp12 inputs for the final exponentiation that are NOT Miller-loop outputs: zero (maps to zero), one, -1,
    elements of the subfields Fp / Fp2 / Fp6 (the easy part sends them to one, so every f^x of the hard part
    runs the degenerate z2 = z3 = 0 decompression), a pure-w element, and random field elements."""
    rng = random.Random(seed)
    z = [0] * 12
    rows = [z[:], [1] + z[1:], [o.P - 1] + z[1:], [rng.randrange(1, o.P)] + z[1:],
            [rng.randrange(o.P), rng.randrange(o.P)] + z[2:],                       # Fp2
            [rng.randrange(o.P) for _ in range(6)] + z[6:],                         # Fp6 (c1 = 0)
            z[:6] + [rng.randrange(o.P) for _ in range(6)],                         # c0 = 0
            z[:6] + [1] + z[7:]]                                                    # w
    rows += [[rng.randrange(o.P) for _ in range(12)] for _ in range(n_random)]
    return np.stack([fp_arr(r) for r in rows])


def group_add_cases(coracle, n_random=6, seed=0xADD):
    """(a, a_inf, b, b_inf) rows for G1 and G2 covering every branch of the reference's affine `add`
    (src/g1.rs:155-187, src/g2.rs:210-242): generic chord, P + P (tangent), P + (-P) (the reference panics),
    identity + Q, P + identity, identity + identity."""
    g1, _, g2, _ = oracle_points(coracle, seed, 0, 2 * n_random + 2)

    def build(pts, w):
        half = w // 2
        a = [pts[i] for i in range(n_random)]
        b = [pts[n_random + i] for i in range(n_random)]
        ai, bi = [0] * n_random, [0] * n_random
        p = pts[2 * n_random]
        q = pts[2 * n_random + 1]
        neg = p.copy()
        yl = [o.fp_from_u64([int(v) for v in p[half + 6 * k: half + 6 * k + 6]]) for k in range(half // 6)]
        neg[half:] = fp_arr([(o.P - y) % o.P for y in yl])
        for x, y, xi, yi in ((p, p, 0, 0), (p, neg, 0, 0), (p, q, 1, 0), (p, q, 0, 1), (p, q, 1, 1)):
            a.append(x); b.append(y); ai.append(xi); bi.append(yi)
        return np.stack(a), np.array(ai, np.uint8), np.stack(b), np.array(bi, np.uint8)
    return build(g1, 12), build(g2, 24)


def oracle_group_add(coracle, group, a, ai, b, bi):
    """Element-wise oracle sums -> (points, identity flags, panics) ; `panics` marks P + (-P)."""
    w = a.shape[1]
    out, inf, pan = np.zeros_like(a), np.zeros(len(a), np.uint8), np.zeros(len(a), bool)
    for j in range(len(a)):
        if not ai[j] and not bi[j] and np.array_equal(a[j][: w // 2], b[j][: w // 2]) and not np.array_equal(a[j][w // 2:], b[j][w // 2:]):
            pan[j] = True      # equal x, different y: the reference divides by zero (src/g1.rs:177)
            inf[j] = 1
            continue
        out[j], inf[j] = coracle.group_op(group, "add", a[j], int(ai[j]), b[j], int(bi[j]))
    return out, inf, pan
