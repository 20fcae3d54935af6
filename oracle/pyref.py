"""CPU ORACLE (test infrastructure, NOT product code) -- Python big-int restatement.

This file restates, with exact Python integers, the arithmetic of the reference crate
0xWOLAND/zkvm-pairings (``/root/reference/src``) that sits on the pairing hot path, plus the
pairing itself (which the reference declares in ``src/lib.rs:12`` but leaves EMPTY:
``src/pairings.rs`` is 0 bytes).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``zkvm_pairings_b200``) never does.

Parity status
-------------
* Tower + groups (Fp, Fp2, Fp6, Fp12, G1Affine, G2Affine): PINNED against every known-answer
  vector the reference's own tests hold (``tests/golden/reference_kats.json``, extracted from
  ``src/fp.rs:577-588``, ``src/g1.rs:263-301``, ``src/g2.rs:349-443``, ``src/fp6.rs:562-757``,
  ``src/fp12.rs:414-799``) -- see ``tests/test_oracle.py``.
* Miller loop / final exponentiation / Gt: **PARITY UNPINNED by the reference** (no
  implementation, no vector).  They follow the zkcrypto ``bls12_381`` lineage the tower was
  copied from and are pinned by the standard ``e(G1,G2)`` Gt-generator
  vector, bilinearity and ``e^r = 1``.

The reference's host arithmetic is ``num-bigint 0.4.6`` exact integer ``*``/``+`` followed by
``% p`` (``src/fp.rs:351-368``, ``src/fp.rs:415-434``), so Python ``int`` arithmetic ``% P`` is
bit-identical by construction.  Elements are canonical (non-Montgomery) integers in ``[0,p)``
(``src/fp.rs:154-156``).

Representation: Fp = int; Fp2 = (c0, c1); Fp6 = (c0, c1, c2) of Fp2; Fp12 = (c0, c1) of Fp6;
affine points = (x, y, is_infinity).
"""
from __future__ import annotations

import hashlib

# --------------------------------------------------------------------------------------------
# Constants -- src/common.rs:68-157


def _limbs(l):
    v = 0
    for i, w in enumerate(l):
        v |= w << (64 * i)
    return v


# src/common.rs:74-81
P = _limbs([0xb9feffffffffaaab, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624,
            0x64774b84f38512bf, 0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a])
# src/common.rs:72 (|x|; the curve parameter is negative)
X = 0xd201000000010000
# src/common.rs:162-167 (scalar field modulus)
R_ORDER = _limbs([0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48])
# src/common.rs:83-90
BETA = _limbs([0x2e01fffffffefffe, 0xde17d813620a0002, 0xddb3a93be6f89688,
               0xba69c6076a0f77ea, 0x5f19672fdf76ce51, 0x0])
# src/common.rs:92-108
G1_X = _limbs([0xfb3af00adb22c6bb, 0x6c55e83ff97a1aef, 0xa14e3a3f171bac58,
               0xc3688c4f9774b905, 0x2695638c4fa9ac0f, 0x17f1d3a73197d794])
G1_Y = _limbs([0x0caa232946c5e7e1, 0xd03cc744a2888ae4, 0x00db18cb2c04b3ed,
               0xfcf5e095d5d00af6, 0xa09e30ed741d8ae4, 0x08b3f481e3aaa0f1])
# src/common.rs:110-144
G2_X0 = _limbs([0xd48056c8c121bdb8, 0x0bac0326a805bbef, 0xb4510b647ae3d177,
                0xc6e47ad4fa403b02, 0x260805272dc51051, 0x024aa2b2f08f0a91])
G2_X1 = _limbs([0xe5ac7d055d042b7e, 0x334cf11213945d57, 0xb5da61bbdc7f5049,
                0x596bd0d09920b61a, 0x7dacd3a088274f65, 0x13e02b6052719f60])
G2_Y0 = _limbs([0xe193548608b82801, 0x923ac9cc3baca289, 0x6d429a695160d12c,
                0xadfd9baa8cbdd3a7, 0x8cc9cdc6da2e351a, 0x0ce5d527727d6e11])
G2_Y1 = _limbs([0xaaa9075ff05f79be, 0x3f370d275cec1da1, 0x267492ab572e99ab,
                0xcb3e287e85a763af, 0x32acd2b02bc28b99, 0x0606c4a02ea734cc])
# src/common.rs:147-157 (Montgomery parameters; the reference stores canonical values and never
# uses them on the host path, the CUDA engine does)
INV = 0x89f3fffcfffcfffd
R_MONT = _limbs([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba,
                 0x77ce585370525745, 0x5c071a97a256ec6d, 0x15f65ec3fa80e493])
B1 = 4            # src/common.rs:69
B2 = (4, 4)       # src/common.rs:70-71

assert R_MONT == (1 << 384) % P and (INV * P) % (1 << 64) == (1 << 64) - 1
assert R_ORDER == X**4 - X**2 + 1

# Instrumentation: number of Fp multiplications executed (squarings count as multiplications).
FP_MULS = 0


def reset_counter():
    global FP_MULS
    FP_MULS = 0


# --------------------------------------------------------------------------------------------
# Fp -- src/fp.rs


def fp_add(a, b):          # src/fp.rs:351-368
    return (a + b) % P


def fp_neg(a):             # src/fp.rs:381-405
    return (P - a) % P


def fp_sub(a, b):          # src/fp.rs:407-411  (= (-b) + a)
    return (a - b) % P


def fp_mul(a, b):          # src/fp.rs:413-434
    global FP_MULS
    FP_MULS += 1
    return (a * b) % P


def fp_square(a):          # src/fp.rs:452-455
    return fp_mul(a, a)


def fp_pow_vartime(a, by):  # src/fp.rs:264-276 ; by = six little-endian u64 limbs
    res = 1
    for e in reversed(by):
        for i in range(63, -1, -1):
            res = fp_square(res)
            if (e >> i) & 1:
                res = fp_mul(res, a)
    return res


def _to_limbs(v, n=6):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def fp_invert(a):          # src/fp.rs:306-319 ; None for zero
    inv = fp_pow_vartime(a, _to_limbs(P - 2))
    return None if a == 0 else inv


def fp_sqrt(a):            # src/fp.rs:280-300 ; None when not a residue
    s = fp_pow_vartime(a, _to_limbs((P + 1) // 4))
    return s if fp_square(s) == a else None


def fp_div(a, b):          # src/fp.rs:448-450 ; the reference panics on zero
    inv = fp_invert(b)
    if inv is None:
        raise ZeroDivisionError("Fp::div by zero (reference unwrap() panics, src/fp.rs:449)")
    return fp_mul(a, inv)


def fp_from_bytes(b: bytes):  # src/fp.rs:165-191 ; 48 big-endian bytes, canonical or None
    assert len(b) == 48
    v = int.from_bytes(b, "big")
    return v if v < P else None


def fp_to_bytes(a) -> bytes:  # src/fp.rs:195-207
    return a.to_bytes(48, "big")


def fp_from_u768(limbs):   # src/fp.rs:218-232 ; twelve big-endian-ordered u64 limbs
    d1 = _limbs([limbs[11], limbs[10], limbs[9], limbs[8], limbs[7], limbs[6]])
    d0 = _limbs([limbs[5], limbs[4], limbs[3], limbs[2], limbs[1], limbs[0]])
    # NOTE: the reference's from_raw_unchecked does not reduce d0/d1; add and mul reduce.
    return fp_add(d0, fp_mul(d1, R_MONT))


# --------------------------------------------------------------------------------------------
# Fp2 = Fp[u]/(u^2+1) -- src/fp2.rs

FP2_ZERO = (0, 0)
FP2_ONE = (1, 0)


def fp2_add(a, b):         # src/fp2.rs:216-218
    return (fp_add(a[0], b[0]), fp_add(a[1], b[1]))


def fp2_sub(a, b):         # src/fp2.rs:221-223
    return (fp_sub(a[0], b[0]), fp_sub(a[1], b[1]))


def fp2_neg(a):            # src/fp2.rs:226-228
    return (fp_neg(a[0]), fp_neg(a[1]))


def fp2_conjugate(a):      # src/fp2.rs:155-157
    return (a[0], fp_neg(a[1]))


fp2_frobenius_map = fp2_conjugate   # src/fp2.rs:147-151


def fp2_mul_by_nonresidue(a):       # src/fp2.rs:161-168
    return (fp_sub(a[0], a[1]), fp_add(a[0], a[1]))


KARATSUBA = False   # work-model switch: count Fp-muls as the CUDA engine performs them


def fp2_square(a):         # src/fp2.rs:171-189
    s = fp_add(a[0], a[1])
    d = fp_sub(a[0], a[1])
    c = fp_add(a[0], a[0])
    return (fp_mul(s, d), fp_mul(c, a[1]))


def fp2_mul(a, b):         # src/fp2.rs:192-209 (schoolbook, 4 Fp-mul)
    if KARATSUBA:          # same value, 3 Fp-mul (what the CUDA tower does)
        t0 = fp_mul(a[0], b[0])
        t1 = fp_mul(a[1], b[1])
        t2 = fp_mul(fp_add(a[0], a[1]), fp_add(b[0], b[1]))
        return (fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1))
    return (fp_sub(fp_mul(a[0], b[0]), fp_mul(a[1], b[1])),
            fp_add(fp_mul(a[0], b[1]), fp_mul(a[1], b[0])))


def fp2_mul_fp(a, k):      # src/fp2.rs:95-102
    return (fp_mul(a[0], k), fp_mul(a[1], k))


def fp2_invert(a):         # src/fp2.rs:278-296
    t = fp_invert(fp_add(fp_square(a[0]), fp_square(a[1])))
    if t is None:
        return None
    return (fp_mul(a[0], t), fp_mul(a[1], fp_neg(t)))


def fp2_div(a, b):         # src/fp2.rs:211-213
    inv = fp2_invert(b)
    if inv is None:
        raise ZeroDivisionError("Fp2::div by zero (reference panics)")
    return fp2_mul(a, inv)


def fp2_is_zero(a):
    return a[0] == 0 and a[1] == 0


def fp2_pow_vartime(a, by):  # src/fp2.rs:301-313
    res = FP2_ONE
    for e in reversed(by):
        for i in range(63, -1, -1):
            res = fp2_square(res)
            if (e >> i) & 1:
                res = fp2_mul(res, a)
    return res


def fp2_sqrt(a):           # src/fp2.rs:231-273
    if fp2_is_zero(a):
        return FP2_ZERO
    a1 = fp2_pow_vartime(a, _to_limbs((P - 3) // 4))
    alpha = fp2_mul(fp2_square(a1), a)
    x0 = fp2_mul(a1, a)
    if alpha == fp2_neg(FP2_ONE):
        return (fp_neg(x0[1]), x0[0])
    s = fp2_mul(fp2_pow_vartime(fp2_add(alpha, FP2_ONE), _to_limbs((P - 1) // 2)), x0)
    return s if fp2_square(s) == a else None


# --------------------------------------------------------------------------------------------
# Fp6 = Fp2[v]/(v^3-(u+1)) -- src/fp6.rs

FP6_ZERO = (FP2_ZERO, FP2_ZERO, FP2_ZERO)
FP6_ONE = (FP2_ONE, FP2_ZERO, FP2_ZERO)

# True Frobenius coefficients, computed -- not copied -- and asserted below.
FROB6_C1 = fp2_pow_vartime((1, 1), _to_limbs((P - 1) // 3))
FROB6_C2 = fp2_pow_vartime((1, 1), _to_limbs((2 * P - 2) // 3))
FROB12_C1 = fp2_pow_vartime((1, 1), _to_limbs((P - 1) // 6))
assert FROB6_C1 == (0, 0x1a0111ea397fe699ec02408663d4de85aa0d857d89759ad4897d29650fb85f9b409427eb4f49fffd8bfd00000000aaac)
assert FROB6_C2 == (0x1a0111ea397fe699ec02408663d4de85aa0d857d89759ad4897d29650fb85f9b409427eb4f49fffd8bfd00000000aaad, 0)
# src/fp12.rs:150-165 (this one is correct in the reference)
assert FROB12_C1 == (
    _limbs([0x8d0775ed92235fb8, 0xf67ea53d63e7813d, 0x7b2443d784bab9c4,
            0x0fd603fd3cbd5f4f, 0xc231beb4202c0d1f, 0x1904d3bf02bb0667]),
    _limbs([0x2cf78a126ddc4af3, 0x282d5ac14d6c7ec2, 0xec0c8ec971f63c5f,
            0x54a14787b6c7b36f, 0x88e9e902231f9fb8, 0x00fc3e2b36c4e032]))
# The constants the reference actually uses in Fp6::frobenius_map (src/fp6.rs:150-171): WRONG
# for a p-power map (they are the p^2 coefficients); kept only for ref_compat=True.
_REF_FROB6_C1 = (BETA, 0)
_REF_FROB6_C2 = (_limbs([0x8bfd00000000aaac, 0x409427eb4f49fffd, 0x897d29650fb85f9b,
                         0xaa0d857d89759ad4, 0xec02408663d4de85, 0x1a0111ea397fe699]), 0)
reset_counter()


def fp6_add(a, b):         # src/fp6.rs:322-333
    return (fp2_add(a[0], b[0]), fp2_add(a[1], b[1]), fp2_add(a[2], b[2]))


def fp6_sub(a, b):         # src/fp6.rs:358-367
    return (fp2_sub(a[0], b[0]), fp2_sub(a[1], b[1]), fp2_sub(a[2], b[2]))


def fp6_neg(a):            # src/fp6.rs:335-346
    return (fp2_neg(a[0]), fp2_neg(a[1]), fp2_neg(a[2]))


def fp6_mul_by_nonresidue(a):   # src/fp6.rs:128-139
    return (fp2_mul_by_nonresidue(a[2]), a[0], a[1])


def fp6_mul_by_1(a, c1):   # src/fp6.rs:102-108
    return (fp2_mul_by_nonresidue(fp2_mul(a[2], c1)), fp2_mul(a[0], c1), fp2_mul(a[1], c1))


def fp6_mul_by_01(a, c0, c1):   # src/fp6.rs:110-125
    a_a = fp2_mul(a[0], c0)
    b_b = fp2_mul(a[1], c1)
    t1 = fp2_add(fp2_mul_by_nonresidue(fp2_mul(a[2], c1)), a_a)
    t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(c0, c1), fp2_add(a[0], a[1])), a_a), b_b)
    t3 = fp2_add(fp2_mul(a[2], c0), b_b)
    return (t1, t2, t3)


def fp6_mul(a, b):         # src/fp6.rs:188-267 (mul_interleaved, the Mul impl :312-319)
    if KARATSUBA:          # same value, 6 Fp2-mul Karatsuba (what the CUDA tower does)
        v0 = fp2_mul(a[0], b[0])
        v1 = fp2_mul(a[1], b[1])
        v2 = fp2_mul(a[2], b[2])
        t0 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a[1], a[2]), fp2_add(b[1], b[2])), v1), v2)
        t1 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a[0], a[1]), fp2_add(b[0], b[1])), v0), v1)
        t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a[0], a[2]), fp2_add(b[0], b[2])), v0), v2)
        return (fp2_add(v0, fp2_mul_by_nonresidue(t0)),
                fp2_add(t1, fp2_mul_by_nonresidue(v2)),
                fp2_add(t2, v1))
    m, ad, sb = fp_mul, fp_add, fp_sub
    (a00, a01), (a10, a11), (a20, a21) = a
    (b00, b01), (b10, b11), (b20, b21) = b
    b10p, b10m = ad(b10, b11), sb(b10, b11)
    b20p, b20m = ad(b20, b21), sb(b20, b21)

    def sop(terms):
        acc = 0
        for sign, x, y in terms:
            acc = ad(acc, m(x, y)) if sign > 0 else sb(acc, m(x, y))
        return acc

    c00 = sop([(1, a00, b00), (-1, a01, b01), (1, a10, b20m), (-1, a11, b20p), (1, a20, b10m), (-1, a21, b10p)])
    c01 = sop([(1, a00, b01), (1, a01, b00), (1, a10, b20p), (1, a11, b20m), (1, a20, b10p), (1, a21, b10m)])
    c10 = sop([(1, a00, b10), (-1, a01, b11), (1, a10, b00), (-1, a11, b01), (1, a20, b20m), (-1, a21, b20p)])
    c11 = sop([(1, a00, b11), (1, a01, b10), (1, a10, b01), (1, a11, b00), (1, a20, b20p), (1, a21, b20m)])
    c20 = sop([(1, a00, b20), (-1, a01, b21), (1, a10, b10), (-1, a11, b11), (1, a20, b00), (-1, a21, b01)])
    c21 = sop([(1, a00, b21), (1, a01, b20), (1, a10, b11), (1, a11, b10), (1, a20, b01), (1, a21, b00)])
    return ((c00, c01), (c10, c11), (c20, c21))


def fp6_square(a):         # src/fp6.rs:274-288
    s0 = fp2_square(a[0])
    ab = fp2_mul(a[0], a[1])
    s1 = fp2_add(ab, ab)
    s2 = fp2_square(fp2_add(fp2_sub(a[0], a[1]), a[2]))
    bc = fp2_mul(a[1], a[2])
    s3 = fp2_add(bc, bc)
    s4 = fp2_square(a[2])
    return (fp2_add(fp2_mul_by_nonresidue(s3), s0),
            fp2_add(fp2_mul_by_nonresidue(s4), s1),
            fp2_sub(fp2_sub(fp2_add(fp2_add(s1, s2), s3), s0), s4))


def fp6_invert(a):         # src/fp6.rs:291-309
    c0 = fp2_sub(fp2_square(a[0]), fp2_mul_by_nonresidue(fp2_mul(a[1], a[2])))
    c1 = fp2_sub(fp2_mul_by_nonresidue(fp2_square(a[2])), fp2_mul(a[0], a[1]))
    c2 = fp2_sub(fp2_square(a[1]), fp2_mul(a[0], a[2]))
    tmp = fp2_mul_by_nonresidue(fp2_add(fp2_mul(a[1], c2), fp2_mul(a[2], c1)))
    tmp = fp2_add(tmp, fp2_mul(a[0], c0))
    t = fp2_invert(tmp)
    if t is None:
        return None
    return (fp2_mul(t, c0), fp2_mul(t, c1), fp2_mul(t, c2))


def fp6_mul_fp(a, k):      # src/fp6.rs:369-380
    return (fp2_mul_fp(a[0], k), fp2_mul_fp(a[1], k), fp2_mul_fp(a[2], k))


def fp6_frobenius_map(a, ref_compat=False):
    """a^p.  src/fp6.rs:142-176 conjugates the three coefficients and multiplies c1, c2 by
    constants -- but the constants there are the p^2 ones, so the reference map is NOT a^p
    (SURVEY.md 0.5).  ref_compat=True reproduces the reference's (wrong) output; the default is
    the true map, which is what the CUDA engine implements."""
    k1, k2 = (_REF_FROB6_C1, _REF_FROB6_C2) if ref_compat else (FROB6_C1, FROB6_C2)
    return (fp2_frobenius_map(a[0]),
            fp2_mul(fp2_frobenius_map(a[1]), k1),
            fp2_mul(fp2_frobenius_map(a[2]), k2))


def fp6_is_zero(a):
    return all(fp2_is_zero(c) for c in a)


# --------------------------------------------------------------------------------------------
# Fp12 = Fp6[w]/(w^2-v) -- src/fp12.rs

FP12_ZERO = (FP6_ZERO, FP6_ZERO)
FP12_ONE = (FP6_ONE, FP6_ZERO)


def fp12_add(a, b):        # src/fp12.rs:212-220
    return (fp6_add(a[0], b[0]), fp6_add(a[1], b[1]))


def fp12_sub(a, b):        # src/fp12.rs:240-247
    return (fp6_sub(a[0], b[0]), fp6_sub(a[1], b[1]))


def fp12_neg(a):           # src/fp12.rs:222-229
    return (fp6_neg(a[0]), fp6_neg(a[1]))


def fp12_conjugate(a):     # src/fp12.rs:123-125
    return (a[0], fp6_neg(a[1]))


def fp12_mul(a, b):        # src/fp12.rs:193-210
    aa = fp6_mul(a[0], b[0])
    bb = fp6_mul(a[1], b[1])
    o = fp6_add(b[0], b[1])
    c1 = fp6_mul(fp6_add(a[1], a[0]), o)
    c1 = fp6_sub(fp6_sub(c1, aa), bb)
    c0 = fp6_add(fp6_mul_by_nonresidue(bb), aa)
    return (c0, c1)


def fp12_square(a):        # src/fp12.rs:173-184
    ab = fp6_mul(a[0], a[1])
    c0c1 = fp6_add(a[0], a[1])
    c0 = fp6_add(fp6_mul_by_nonresidue(a[1]), a[0])
    c0 = fp6_sub(fp6_mul(c0, c0c1), ab)
    c1 = fp6_add(ab, ab)
    c0 = fp6_sub(c0, fp6_mul_by_nonresidue(ab))
    return (c0, c1)


def fp12_mul_by_014(a, c0, c1, c4):   # src/fp12.rs:99-111
    aa = fp6_mul_by_01(a[0], c0, c1)
    bb = fp6_mul_by_1(a[1], c4)
    o = fp2_add(c1, c4)
    r1 = fp6_mul_by_01(fp6_add(a[1], a[0]), c0, o)
    r1 = fp6_sub(fp6_sub(r1, aa), bb)
    r0 = fp6_add(fp6_mul_by_nonresidue(bb), aa)
    return (r0, r1)


def fp12_invert(a):        # src/fp12.rs:186-190
    t = fp6_invert(fp6_sub(fp6_square(a[0]), fp6_mul_by_nonresidue(fp6_square(a[1]))))
    if t is None:
        return None
    return (fp6_mul(a[0], t), fp6_mul(a[1], fp6_neg(t)))


def fp12_div(a, b):        # src/fp12.rs:113-115
    inv = fp12_invert(b)
    if inv is None:
        raise ZeroDivisionError("Fp12::div by zero (reference panics)")
    return fp12_mul(inv, a)


def fp12_mul_fp(a, k):     # src/fp12.rs:249-256
    return (fp6_mul_fp(a[0], k), fp6_mul_fp(a[1], k))


def fp12_frobenius_map(a, ref_compat=False):   # src/fp12.rs:143-170
    c0 = fp6_frobenius_map(a[0], ref_compat)
    c1 = fp6_frobenius_map(a[1], ref_compat)
    # the reference multiplies by Fp6::from(Fp2 const) = (k,0,0) with a FULL Fp6 mul
    c1 = fp6_mul(c1, (FROB12_C1, FP2_ZERO, FP2_ZERO))
    return (c0, c1)


def fp12_pow_vartime(a, by):   # src/fp12.rs:127-139 ; ``by`` little-endian u64 limbs (any count)
    res = FP12_ONE
    for e in reversed(by):
        for i in range(63, -1, -1):
            res = fp12_square(res)
            if (e >> i) & 1:
                res = fp12_mul(res, a)
    return res


def fp12_pow_int(a, e: int):
    n = max(1, (e.bit_length() + 63) // 64)
    return fp12_pow_vartime(a, _to_limbs(e, n))


def fp12_is_zero(a):
    return fp6_is_zero(a[0]) and fp6_is_zero(a[1])


# --------------------------------------------------------------------------------------------
# G1Affine -- src/g1.rs ; points are (x, y, is_infinity)

G1_IDENTITY = (0, 1, True)          # src/g1.rs:25-31
G1_GENERATOR = (G1_X, G1_Y, False)  # src/g1.rs:41-47


def g1_neg(p):             # src/g1.rs:118-128
    return (p[0], fp_neg(p[1]), p[2])


def g1_double(p):          # src/g1.rs:74-91
    x, y, inf = p
    if inf:
        return G1_IDENTITY
    slope = fp_div(fp_mul(3, fp_square(x)), fp_mul(2, y))
    xr = fp_sub(fp_square(slope), fp_mul(2, x))
    yr = fp_sub(fp_mul(slope, fp_sub(x, xr)), y)
    return (xr, yr, False)


def g1_add(p, q):          # src/g1.rs:155-187 ; P + (-P) panics in the reference (:177)
    if p[2]:
        return q
    if q[2]:
        return p
    x1, y1, _ = p
    x2, y2, _ = q
    if x1 == x2 and y1 == y2:
        return g1_double(p)
    if x1 == x2:               # P + (-P): the reference divides by zero and panics (src/g1.rs:177);
        return G1_IDENTITY     # the oracle returns the identity so r*G = O can be exercised
    slope = fp_div(fp_sub(y2, y1), fp_sub(x2, x1))
    xr = fp_sub(fp_sub(fp_square(slope), x1), x2)
    yr = fp_sub(fp_mul(slope, fp_sub(x1, xr)), y1)
    return (xr, yr, False)


def g1_mul(p, k: int, ref_compat=False):
    """Scalar multiplication.  ref_compat=True reproduces src/g1.rs:130-153 literally: LSB-first
    over the 256 scalar bits with ``.skip(1)`` -- which DROPS bit 0 and doubles before adding, so
    it returns (k & ~1)*P.  Default: correct k*P."""
    if ref_compat:
        xself, acc = G1_IDENTITY, p
        for i in range(1, 256):
            acc = g1_double(acc)
            if (k >> i) & 1:
                xself = g1_add(xself, acc)
        return xself
    acc = G1_IDENTITY
    for i in range(k.bit_length() - 1, -1, -1):
        acc = g1_double(acc)
        if (k >> i) & 1:
            acc = g1_add(acc, p)
    return acc


def g1_is_on_curve(p):     # src/g1.rs:95-101
    return fp_square(p[1]) == fp_add(fp_mul(fp_square(p[0]), p[0]), B1)


def g1_is_torsion_free(p):  # src/g1.rs:103-115 (semantics: -[x^2]P == (beta*x, y)); uses correct mul
    lhs = g1_neg(g1_mul(g1_mul(p, X), X))
    rhs = (fp_mul(p[0], BETA), p[1], False)
    return lhs[0] == rhs[0] and lhs[1] == rhs[1]


def g1_is_valid(p):        # src/g1.rs:49-62
    if p[2]:
        return True
    return g1_is_on_curve(p) and g1_is_torsion_free(p)


# --------------------------------------------------------------------------------------------
# G2Affine -- src/g2.rs

G2_IDENTITY = (FP2_ZERO, FP2_ONE, True)                       # src/g2.rs:27-33
G2_GENERATOR = ((G2_X0, G2_X1), (G2_Y0, G2_Y1), False)        # src/g2.rs:43-55


def g2_neg(p):             # src/g2.rs:172-182
    return (p[0], fp2_neg(p[1]), p[2])


def g2_double(p):          # src/g2.rs:81-105
    x, y, inf = p
    if inf or fp2_is_zero(y):
        return G2_IDENTITY
    slope = fp2_div(fp2_mul_fp(fp2_square(x), 3), fp2_mul_fp(y, 2))
    xn = fp2_sub(fp2_square(slope), fp2_mul_fp(x, 2))
    yn = fp2_sub(fp2_mul(slope, fp2_sub(x, xn)), y)
    return (xn, yn, False)


def g2_add(p, q):          # src/g2.rs:210-242
    if p[2]:
        return q
    if q[2]:
        return p
    x1, y1, _ = p
    x2, y2, _ = q
    if x1 == x2 and y1 == y2:
        return g2_double(p)
    if x1 == x2:               # P + (-P): reference panics (src/g2.rs:232); oracle returns identity
        return G2_IDENTITY
    slope = fp2_div(fp2_sub(y2, y1), fp2_sub(x2, x1))
    xr = fp2_sub(fp2_sub(fp2_square(slope), x1), x2)
    yr = fp2_sub(fp2_mul(slope, fp2_sub(x1, xr)), y1)
    return (xr, yr, False)


def g2_mul(p, k: int):     # src/g2.rs:185-208 (MSB-first, skips bit 255; correct for k < 2^255)
    acc = G2_IDENTITY
    for i in range(254, -1, -1):
        acc = g2_double(acc)
        if (k >> i) & 1:
            acc = g2_add(acc, p)
    return acc


def g2_is_on_curve(p):     # src/g2.rs:109-120
    return fp2_square(p[1]) == fp2_add(fp2_mul(fp2_square(p[0]), p[0]), B2)


# src/g2.rs:128-157
PSI_COEFF_X = (0, _limbs([0x8bfd00000000aaad, 0x409427eb4f49fffd, 0x897d29650fb85f9b,
                          0xaa0d857d89759ad4, 0xec02408663d4de85, 0x1a0111ea397fe699]))
PSI_COEFF_Y = (_limbs([0xf1ee7b04121bdea2, 0x304466cf3e67fa0a, 0xef396489f61eb45e,
                       0x1c3dedd930b1cf60, 0xe2e9c448d77a2cd9, 0x135203e60180a68e]),
               _limbs([0xc81084fbede3cc09, 0xee67992f72ec05f4, 0x77f76e17009241c5,
                       0x48395dabc2d3435e, 0x6831e36d6bd17ffe, 0x06af0e0437ff400b]))


def g2_psi(p):             # src/g2.rs:126-164
    return (fp2_mul(fp2_frobenius_map(p[0]), PSI_COEFF_X),
            fp2_mul(fp2_frobenius_map(p[1]), PSI_COEFF_Y), False)


def g2_is_torsion_free(p):  # src/g2.rs:166-170 : psi(P) == -[|x|]P
    lhs = g2_psi(p)
    rhs = g2_neg(g2_mul(p, X))
    return lhs[0] == rhs[0] and lhs[1] == rhs[1]


def g2_is_valid(p):        # src/g2.rs:57-69
    if p[2]:
        return True
    return g2_is_on_curve(p) and g2_is_torsion_free(p)


# --------------------------------------------------------------------------------------------
# Pairing -- src/pairings.rs is EMPTY in the reference; algorithm = zkcrypto bls12_381 lineage,
# restated from SURVEY.md section 9 (PARITY UNPINNED by the reference; pinned by 9.4 vectors).


def _doubling_step(r):     # SURVEY 9.1 ; r = (x, y, z) projective over Fp2
    x, y, z = r
    t0 = fp2_square(x)
    t1 = fp2_square(y)
    t2 = fp2_square(t1)
    t3 = fp2_sub(fp2_sub(fp2_square(fp2_add(t1, x)), t0), t2)
    t3 = fp2_add(t3, t3)
    t4 = fp2_add(fp2_add(t0, t0), t0)
    t6 = fp2_add(x, t4)
    t5 = fp2_square(t4)
    zz = fp2_square(z)
    xn = fp2_sub(fp2_sub(t5, t3), t3)
    zn = fp2_sub(fp2_sub(fp2_square(fp2_add(z, y)), t1), zz)
    yn = fp2_mul(fp2_sub(t3, xn), t4)
    t2 = fp2_add(t2, t2)
    t2 = fp2_add(t2, t2)
    t2 = fp2_add(t2, t2)
    yn = fp2_sub(yn, t2)
    c1 = fp2_mul(t4, zz)
    c1 = fp2_neg(fp2_add(c1, c1))
    c2 = fp2_sub(fp2_sub(fp2_square(t6), t0), t5)
    t1 = fp2_add(t1, t1)
    t1 = fp2_add(t1, t1)
    c2 = fp2_sub(c2, t1)
    c0 = fp2_mul(zn, zz)
    c0 = fp2_add(c0, c0)
    return (xn, yn, zn), (c0, c1, c2)


def _addition_step(r, q):  # SURVEY 9.1 ; q = affine (x, y)
    x, y, z = r
    qx, qy = q
    zz = fp2_square(z)
    yy = fp2_square(qy)
    t0 = fp2_mul(zz, qx)
    t1 = fp2_mul(fp2_sub(fp2_sub(fp2_square(fp2_add(qy, z)), yy), zz), zz)
    t2 = fp2_sub(t0, x)
    t3 = fp2_square(t2)
    t4 = fp2_add(t3, t3)
    t4 = fp2_add(t4, t4)
    t5 = fp2_mul(t4, t2)
    t6 = fp2_sub(fp2_sub(t1, y), y)
    t9 = fp2_mul(t6, qx)
    t7 = fp2_mul(t4, x)
    xn = fp2_sub(fp2_sub(fp2_sub(fp2_square(t6), t5), t7), t7)
    zn = fp2_sub(fp2_sub(fp2_square(fp2_add(z, t2)), zz), t3)
    t10 = fp2_add(qy, zn)
    t8 = fp2_mul(fp2_sub(t7, xn), t6)
    t0 = fp2_mul(y, t5)
    t0 = fp2_add(t0, t0)
    yn = fp2_sub(t8, t0)
    t10 = fp2_sub(fp2_square(t10), yy)
    zt2 = fp2_square(zn)
    t10 = fp2_sub(t10, zt2)
    t9 = fp2_sub(fp2_add(t9, t9), t10)
    c0 = fp2_add(zn, zn)
    t6 = fp2_neg(t6)
    c1 = fp2_add(t6, t6)
    return (xn, yn, zn), (c0, c1, t9)


def _ell(f, co, p):        # SURVEY 9.1 ; uses Fp12::mul_by_014 (src/fp12.rs:99-111)
    c0, c1, c2 = co
    a = fp2_mul_fp(c0, p[1])
    b = fp2_mul_fp(c1, p[0])
    return fp12_mul_by_014(f, c2, b, a)


def multi_miller_loop(pairs):
    """pairs: list of (G1 (x,y,inf), G2 (x,y,inf)).  All pairs share the accumulator f; every
    pair's line step runs before each squaring.  Pairs with either point at infinity are skipped
    (contribute one).  Returns the UNexponentiated Fp12 (conjugated, x<0)."""
    live = [(p, q) for (p, q) in pairs if not p[2] and not q[2]]
    rs = [(q[0], q[1], FP2_ONE) for (_, q) in live]
    f = FP12_ONE
    found = False
    for b in range(63, -1, -1):
        i = ((X >> 1) >> b) & 1
        if not found:
            found = bool(i)
            continue
        for k, (p, q) in enumerate(live):
            rs[k], co = _doubling_step(rs[k])
            f = _ell(f, co, p)
        if i:
            for k, (p, q) in enumerate(live):
                rs[k], co = _addition_step(rs[k], (q[0], q[1]))
                f = _ell(f, co, p)
        f = fp12_square(f)
    for k, (p, q) in enumerate(live):
        rs[k], co = _doubling_step(rs[k])
        f = _ell(f, co, p)
    return fp12_conjugate(f)


def miller_loop(p, q):
    return multi_miller_loop([(p, q)])


def _fp4_square(a, b):
    t0 = fp2_square(a)
    t1 = fp2_square(b)
    c0 = fp2_add(fp2_mul_by_nonresidue(t1), t0)
    c1 = fp2_sub(fp2_sub(fp2_square(fp2_add(a, b)), t0), t1)
    return c0, c1


def cyclotomic_square(f):  # SURVEY 9.2 (Granger-Scott)
    z0, z4, z3 = f[0]
    z2, z1, z5 = f[1]
    t0, t1 = _fp4_square(z0, z1)
    z0 = fp2_sub(t0, z0)
    z0 = fp2_add(fp2_add(z0, z0), t0)
    z1 = fp2_add(t1, z1)
    z1 = fp2_add(fp2_add(z1, z1), t1)
    t0, t1 = _fp4_square(z2, z3)
    t2, t3 = _fp4_square(z4, z5)
    z4 = fp2_sub(t0, z4)
    z4 = fp2_add(fp2_add(z4, z4), t0)
    z5 = fp2_add(t1, z5)
    z5 = fp2_add(fp2_add(z5, z5), t1)
    t0 = fp2_mul_by_nonresidue(t3)
    z2 = fp2_add(t0, z2)
    z2 = fp2_add(fp2_add(z2, z2), t0)
    z3 = fp2_sub(t2, z3)
    z3 = fp2_add(fp2_add(z3, z3), t2)
    return ((z0, z4, z3), (z2, z1, z5))


def cyclotomic_exp(f):     # SURVEY 9.2 : f^|x| then conjugate (x<0)
    tmp = FP12_ONE
    found = False
    for b in range(63, -1, -1):
        i = (X >> b) & 1
        if found:
            tmp = cyclotomic_square(tmp)
        else:
            found = bool(i)
        if i:
            tmp = fp12_mul(tmp, f)
    return fp12_conjugate(tmp)


def final_exponentiation(f):   # SURVEY 9.2 ; f^((p^12-1)/r) up to the lineage's fixed cofactor
    fr = fp12_frobenius_map
    t0 = fr(fr(fr(fr(fr(fr(f))))))
    t1 = fp12_invert(f)
    if t1 is None:
        raise ZeroDivisionError("final_exponentiation of zero")
    t2 = fp12_mul(t0, t1)
    t1 = t2
    t2 = fp12_mul(fr(fr(t2)), t1)
    t1 = fp12_conjugate(cyclotomic_square(t2))
    t3 = cyclotomic_exp(t2)
    t4 = cyclotomic_square(t3)
    t5 = fp12_mul(t1, t3)
    t1 = cyclotomic_exp(t5)
    t0 = cyclotomic_exp(t1)
    t6 = fp12_mul(cyclotomic_exp(t0), t4)
    t4 = cyclotomic_exp(t6)
    t5 = fp12_conjugate(t5)
    t4 = fp12_mul(t4, fp12_mul(t5, t2))
    t5 = fp12_conjugate(t2)
    t1 = fp12_mul(t1, t2)
    t1 = fr(fr(fr(t1)))
    t6 = fp12_mul(t6, t5)
    t6 = fr(t6)
    t3 = fp12_mul(t3, t0)
    t3 = fr(fr(t3))
    t3 = fp12_mul(t3, t1)
    t3 = fp12_mul(t3, t6)
    return fp12_mul(t3, t4)


def pairing(p, q):
    return final_exponentiation(miller_loop(p, q))


def multi_pairing(pairs):
    return final_exponentiation(multi_miller_loop(pairs))


# --------------------------------------------------------------------------------------------
# Boundary (de)serialisation: canonical little-endian u64 limbs, exactly ``Fp.0`` (src/fp.rs:24)


def fp_to_u64(a):
    return _to_limbs(a, 6)


def fp_from_u64(l):
    return _limbs(l)


def fp12_flatten(a):
    """Fp12 -> twelve Fp in boundary order c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1."""
    return [c for six in a for two in six for c in two]


def fp12_unflatten(l):
    assert len(l) == 12
    return (((l[0], l[1]), (l[2], l[3]), (l[4], l[5])), ((l[6], l[7]), (l[8], l[9]), (l[10], l[11])))


def fp12_to_u64(a):
    out = []
    for c in fp12_flatten(a):
        out.extend(fp_to_u64(c))
    return out


def fp12_from_u64(l):
    return fp12_unflatten([fp_from_u64(l[6 * i:6 * i + 6]) for i in range(12)])


def fp12_sha256(a) -> str:
    """SHA-256 over the 12 x 48-byte big-endian concatenation."""
    return hashlib.sha256(b"".join(fp_to_bytes(c) for c in fp12_flatten(a))).hexdigest()


def splitmix64(state):
    """SplitMix64 step: returns (new_state, output).  Used for all seeded synthetic inputs so the
    C oracle, the CUDA generator and the tests agree bit-for-bit."""
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)
