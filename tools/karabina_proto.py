"""Prototype (big-int, oracle arithmetic) of the compressed cyclotomic squaring used by the staged final exponentiation (pairing.cuh cexp_begin / cexp_end):
B/C-only Granger-Scott squarings, decompression of (z0, z1) from (z2..z5), and the f^|x| chain with three
decompression points.  Checks every step against oracle/pyref.py's cyclotomic_square / cyclotomic_exp."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pyref as o

X = o.X if o.X > 0 else -o.X


def z_of(f):
    (c00, c01, c02), (c10, c11, c12) = f
    return [c00, c11, c10, c02, c01, c12]          # z0..z5


def f_of(z):
    z0, z1, z2, z3, z4, z5 = z
    return ((z0, z4, z3), (z2, z1, z5))


def comp_sqr(z2, z3, z4, z5):
    t0, t1 = o._fp4_square(z2, z3)
    t2, t3 = o._fp4_square(z4, z5)
    three = lambda t: o.fp2_add(o.fp2_add(t, t), t)
    dbl = lambda t: o.fp2_add(t, t)
    nz4 = o.fp2_sub(three(t0), dbl(z4))
    nz5 = o.fp2_add(three(t1), dbl(z5))
    nz2 = o.fp2_add(three(o.fp2_mul_by_nonresidue(t3)), dbl(z2))
    nz3 = o.fp2_sub(three(t2), dbl(z3))
    return nz2, nz3, nz4, nz5


def decompress(z2, z3, z4, z5):
    dbl = lambda t: o.fp2_add(t, t)
    if not o.fp2_is_zero(z2):
        num = o.fp2_add(o.fp2_mul_by_nonresidue(o.fp2_square(z5)), o.fp2_sub(o.fp2_add(dbl(o.fp2_square(z4)), o.fp2_square(z4)), dbl(z3)))
        den = dbl(dbl(z2))
        z1 = o.fp2_mul(num, o.fp2_invert(den))
        t = o.fp2_add(dbl(o.fp2_square(z1)), o.fp2_mul(z2, z5))
    else:
        num = dbl(o.fp2_mul(z4, z5))
        inv = o.fp2_invert(z3) if not o.fp2_is_zero(z3) else o.FP2_ZERO
        z1 = o.fp2_mul(num, inv)
        t = dbl(o.fp2_square(z1))
    m = o.fp2_mul(z3, z4)
    t = o.fp2_sub(t, o.fp2_add(dbl(m), m))
    z0 = o.fp2_add(o.fp2_mul_by_nonresidue(t), o.FP2_ONE)
    return z0, z1


def main():
    e = o.pairing(o.G1_GENERATOR, o.G2_GENERATOR)
    for f in (e, o.fp12_square(e), o.FP12_ONE):
        z = z_of(f)
        assert f_of(z) == f
        c = tuple(z[2:])
        g = f
        snaps = {}
        for i in range(1, 58):
            c = comp_sqr(*c)
            g = o.cyclotomic_square(g)
            assert tuple(z_of(g)[2:]) == c, i
            if i in (16, 48, 57):
                z0, z1 = decompress(*c)
                assert (z0, z1) == tuple(z_of(g)[:2]), ("decompress", i)
                snaps[i] = g
        # bits of |x|: 63 62 60 57 48 16
        assert X == sum(1 << b for b in (63, 62, 60, 57, 48, 16))
        a = snaps[57]
        acc = o.fp12_mul(snaps[16], snaps[48])
        acc = o.fp12_mul(acc, a)
        for _ in range(3):
            a = o.cyclotomic_square(a)
        acc = o.fp12_mul(acc, a)      # 2^60
        for _ in range(2):
            a = o.cyclotomic_square(a)
        acc = o.fp12_mul(acc, a)      # 2^62
        a = o.cyclotomic_square(a)
        acc = o.fp12_mul(acc, a)      # 2^63
        assert o.fp12_conjugate(acc) == o.cyclotomic_exp(f)
    print("karabina prototype ok")


if __name__ == "__main__":
    main()
