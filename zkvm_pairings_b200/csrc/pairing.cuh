// Optimal-ate pairing on BLS12-381: Miller loop + final exponentiation (device code).
//
// The reference declares this module (/root/reference/src/lib.rs:12) but src/pairings.rs is a
// 0-byte file; the algorithm is the zkcrypto bls12_381 lineage the reference tower was copied
// from, as specified in SURVEY.md section 9: projective G2 doubling/addition line steps, sparse
// line * Fp12 via Fp12::mul_by_014 (src/fp12.rs:99-111), Frobenius easy part + one inversion,
// Granger-Scott cyclotomic squarings in the hard part.
#pragma once
#include "tower.cuh"

namespace zkp {

struct G1A { Fp x, y; };          // affine G1 (src/g1.rs:7-11), Montgomery coordinates, same in both lanes
struct G2A { Fp2 x, y; };         // affine G2 (src/g2.rs:8-12), lane-split like every Fp2
struct G2P { Fp2 x, y, z; };      // Jacobian-style projective G2 used by the line steps

// SURVEY 9.1 doubling_step: 8 Fp2 sqr + 3 Fp2 mul.  co = (c0, c1, c2).  r stays normalized.
ZKP_NOINLINE void doubling_step(G2P &r, Fp2 *co) {
    Fp2 t0 = fp2_sqr(r.x);
    Fp2 t1 = fp2_sqr(r.y);
    Fp2 t2 = fp2_sqr(t1);
    Fp2 t3 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(t1, r.x)), t0), t2);
    t3 = fp2_dbl(t3);
    Fp2 t4 = fp2_add(fp2_dbl(t0), t0);
    Fp2 t6 = fp2_add(r.x, t4);
    Fp2 t5 = fp2_sqr(t4);
    Fp2 zz = fp2_sqr(r.z);
    Fp2 xn = fp2_sub(fp2_sub(t5, t3), t3);
    Fp2 zn = fp2_sub(fp2_sub(fp2_sqr(fp2_add(r.z, r.y)), t1), zz);
    Fp2 yn = fp2_mul(fp2_sub(t3, xn), t4);
    t2 = fp2_dbl(fp2_dbl(t2));
    r.y = fp2_sub(yn, fp2_dbl(t2));
    r.x = xn;
    r.z = zn;
    co[1] = fp2_neg(fp2_dbl(fp2_mul(t4, zz)));
    t6 = fp2_sub(fp2_sub(fp2_sqr(t6), t0), t5);
    co[2] = fp2_sub(t6, fp2_dbl(fp2_dbl(t1)));
    co[0] = fp2_dbl(fp2_mul(zn, zz));
}

// SURVEY 9.1 addition_step: 8 Fp2 sqr + 7 Fp2 mul.  q normalized; r stays normalized.
ZKP_NOINLINE void addition_step(G2P &r, const G2A &q, Fp2 *co) {
    Fp2 zz = fp2_sqr(r.z);
    Fp2 yy = fp2_sqr(q.y);
    Fp2 t0 = fp2_mul(zz, q.x);
    Fp2 t1 = fp2_mul(fp2_sub(fp2_sub(fp2_sqr(fp2_add(q.y, r.z)), yy), zz), zz);
    Fp2 t2 = fp2_sub(t0, r.x);
    Fp2 t3 = fp2_sqr(t2);
    Fp2 t4 = fp2_dbl(fp2_dbl(t3));
    Fp2 t5 = fp2_mul(t4, t2);
    Fp2 t6 = fp2_sub(fp2_sub(t1, r.y), r.y);
    Fp2 t9 = fp2_mul(t6, q.x);
    Fp2 t7 = fp2_mul(t4, r.x);
    Fp2 xn = fp2_sub(fp2_sub(fp2_sub(fp2_sqr(t6), t5), t7), t7);
    Fp2 zn = fp2_sub(fp2_sub(fp2_sqr(fp2_add(r.z, t2)), zz), t3);
    Fp2 t10 = fp2_add(q.y, zn);
    Fp2 t8 = fp2_mul(fp2_sub(t7, xn), t6);
    t0 = fp2_dbl(fp2_mul(r.y, t5));
    r.y = fp2_sub(t8, t0);
    r.x = xn;
    r.z = zn;
    t10 = fp2_sub(fp2_sub(fp2_sqr(t10), yy), fp2_sqr(zn));
    co[2] = fp2_sub(fp2_dbl(t9), t10);
    co[0] = fp2_dbl(zn);
    co[1] = fp2_dbl(fp2_neg(t6));
}

// SURVEY 9.1 ell: scale the line by P and fold it into f.  A pair flagged `skip` (a point at
// infinity) multiplies f by the line (1, 0, 0) = one instead -- by selects, not by a branch, so all
// lanes of a warp stay on one path.
ZKP_HD Fp2 fp2_select(bool c, const Fp2 &a, const Fp2 &b) { Fp2 r; r.c = fp_select(c, a.c, b.c); return r; }
ZKP_HD void ell(Fp12 &f, const Fp2 *co, const G1A &p, bool skip) {
    Fp2 a = fp2_select(skip, fp2_zero(), fp2_mul_fp(co[0], p.y));
    Fp2 b = fp2_select(skip, fp2_zero(), fp2_mul_fp(co[1], p.x));
    fp12_mul_by_014(f, fp2_select(skip, fp2_one(), co[2]), b, a);
}

// Bits of |x| >> 1 below its leading one (bit 62), MSB first: 62 iterations, additions where set.
#define ZKP_X_HALF (ZKP_BLS_X >> 1)

// Prepared G2 points ("G2Prepared" of the zkcrypto lineage): the 68 line-coefficient
// triples the loop below derives from Q (63 doubling + 5 addition steps), computed once for points
// that stay fixed across checks (verifying-key points).  Device format: Montgomery limbs,
// [point][step][coefficient][lane parity] Fp, so each lane fetches its own half with three 128-bit
// loads and all checks of a warp read the same addresses (one broadcast per load).
#define ZKP_LINE_STEPS 68
ZKP_HD Fp2 line_tab_load(const Fp *tab, int point, int step, int c) {
    Fp2 r;
    r.c = tab[((point * ZKP_LINE_STEPS + step) * 3 + c) * 2 + lane_par()];
    return r;
}

// Miller loop over kv + kf pairs sharing the accumulator f (one pair: single pairing).  The first kv
// pairs bring their own G2 point (qs, scratch rs), the last kf use prepared line tables.  Pairs
// flagged `skip` (a point at infinity) contribute one.  Output is conjugated (x < 0).
ZKP_HD void miller_loop(Fp12 &f, const G1A *ps, const G2A *qs, const bool *skip, G2P *rs, int kv,
                        const Fp *tab = nullptr, int kf = 0) {
    Fp2 co[3];
    fp12_set_one(f);
    for (int j = 0; j < kv; j++) {
        rs[j].x = qs[j].x;
        rs[j].y = qs[j].y;
        rs[j].z = fp2_one();
    }
    int step = 0;
#pragma unroll 1
    for (int b = 61; b >= -1; b--) {   // b = -1: the final doubling step, no squaring after it
        bool bit = b >= 0 && ((ZKP_X_HALF >> b) & 1);
        for (int j = 0; j < kv; j++) {
            doubling_step(rs[j], co);
            ell(f, co, ps[j], skip[j]);
        }
        for (int j = 0; j < kf; j++) {
            for (int c = 0; c < 3; c++) co[c] = line_tab_load(tab, j, step, c);
            ell(f, co, ps[kv + j], skip[kv + j]);
        }
        step++;
        if (bit) {
            for (int j = 0; j < kv; j++) {
                addition_step(rs[j], qs[j], co);
                ell(f, co, ps[j], skip[j]);
            }
            for (int j = 0; j < kf; j++) {
                for (int c = 0; c < 3; c++) co[c] = line_tab_load(tab, j, step, c);
                ell(f, co, ps[kv + j], skip[kv + j]);
            }
            step++;
        }
        if (b >= 0) fp12_sqr(f, f);
    }
    fp12_conj(f, f);
}
// the line table of one G2 point: this lane's half of the 68 x 3 coefficients, in loop order
ZKP_HD void g2_prepare(const G2A &q, Fp *out_lane /* stride 2 Fp per coefficient */) {
    Fp2 co[3];
    G2P r;
    r.x = q.x; r.y = q.y; r.z = fp2_one();
    int step = 0;
#pragma unroll 1
    for (int b = 61; b >= -1; b--) {
        bool bit = b >= 0 && ((ZKP_X_HALF >> b) & 1);
        doubling_step(r, co);
        for (int c = 0; c < 3; c++) out_lane[(step * 3 + c) * 2] = co[c].c;
        step++;
        if (bit) {
            addition_step(r, q, co);
            for (int c = 0; c < 3; c++) out_lane[(step * 3 + c) * 2] = co[c].c;
            step++;
        }
    }
}

// f^|x| followed by conjugation (x < 0); f in the cyclotomic subgroup.  63 squarings + 5 muls.
ZKP_NOINLINE void cyclotomic_exp(Fp12 &r, const Fp12 &f) {
    Fp12 t = f;   // leading bit 63
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        fp12_cyclotomic_sqr(t, t);
        if ((ZKP_BLS_X >> b) & 1) fp12_mul(t, t, f);
    }
    fp12_conj(r, t);
}

// SURVEY 9.2, split around the single Fp inversion of the easy part (f^-1): fe_prepare leaves the
// cofactors and the norm n, fe_finish continues from ninv = 1/n.  The pairing kernels run the two
// halves as separate launches with a batched inversion kernel in between (Montgomery's trick across
// pairings: ~41 Fp products per inverse instead of a 609-product Fermat ladder per lane).
// f must be non-zero (a Miller-loop output always is); zero maps to zero.
struct FeState {
    Fp6 c;   // cofactors of the Fp6 inverse
    Fp2 t;   // the Fp2 whose norm is inverted
};
ZKP_HD Fp fe_prepare(FeState &s, const Fp12 &f) { return fp12_inv_prepare(s.c, s.t, f); }
ZKP_HD void fe_finish(Fp12 &r, const Fp12 &f, const FeState &s, const Fp &ninv) {
    Fp12 t0, t1, t2, t3, t4, t5, t6;
    fp12_conj(t0, f);                 // f^(p^6)
    fp12_inv_finish(t1, f, s.c, s.t, ninv);
    fp12_mul(t2, t0, t1);             // f^(p^6-1)
    t1 = t2;
    fp12_frobenius(t2, t2, 2);
    fp12_mul(t2, t2, t1);             // easy part done
    fp12_cyclotomic_sqr(t1, t2);
    fp12_conj(t1, t1);
    cyclotomic_exp(t3, t2);
    fp12_cyclotomic_sqr(t4, t3);
    fp12_mul(t5, t1, t3);
    cyclotomic_exp(t1, t5);
    cyclotomic_exp(t0, t1);
    cyclotomic_exp(t6, t0);
    fp12_mul(t6, t6, t4);
    cyclotomic_exp(t4, t6);
    fp12_conj(t5, t5);
    fp12_mul(t5, t5, t2);
    fp12_mul(t4, t4, t5);
    fp12_conj(t5, t2);
    fp12_mul(t1, t1, t2);
    fp12_frobenius(t1, t1, 3);
    fp12_mul(t6, t6, t5);
    fp12_frobenius(t6, t6, 1);
    fp12_mul(t3, t3, t0);
    fp12_frobenius(t3, t3, 2);
    fp12_mul(t3, t3, t1);
    fp12_mul(t3, t3, t6);
    fp12_mul(r, t3, t4);
}
ZKP_HD void final_exponentiation(Fp12 &r, const Fp12 &f) {
    FeState s;
    Fp n = fe_prepare(s, f);
    fe_finish(r, f, s, fp_inv(n));
}

// Montgomery's trick over a run of values held by ONE thread: v[i] <- 1/v[i] for i < cnt with a
// single Fermat inversion and 3 (cnt - 1) products.  Zeros (a zero norm: only for a zero Fp12
// input, which maps to zero) are skipped and stay zero.  `pre` is scratch of cnt elements.
ZKP_HD void fp_batch_inv(Fp *v, Fp *pre, int cnt) {
    Fp acc = fp_one();
    for (int i = 0; i < cnt; i++) {
        pre[i] = acc;
        if (!fp_is_zero(v[i])) acc = fmul(acc, v[i]);
    }
    acc = fp_inv(acc);
    for (int i = cnt - 1; i >= 0; i--) {
        if (fp_is_zero(v[i])) continue;
        Fp inv = fmul(acc, pre[i]);
        acc = fmul(acc, v[i]);
        v[i] = inv;
    }
}

// ------------------------------------------------------------------ group helpers (input prep)
//
// [k]P in Jacobian coordinates (a = 0 curves), k a 64-bit scalar, then one Fermat inversion back
// to affine.  Used to synthesise valid subgroup points on the device (k*G1gen, k*G2gen); the
// reference's own random() points are off-curve (src/g1.rs:64-72) and its G1 scalar mul drops
// bit 0 (src/g1.rs:130-153) -- this is the correct double-and-add of src/g2.rs:185-208.
// Field-generic via small traits.
struct OpsFp {
    typedef Fp T;
    static ZKP_MEMBER T add(const T &a, const T &b) { return fp_add(a, b); }
    static ZKP_MEMBER T sub(const T &a, const T &b) { return fp_sub(a, b); }
    static ZKP_MEMBER T mul(const T &a, const T &b) { return fmul(a, b); }
    static ZKP_MEMBER T sqr(const T &a) { return fsqr(a); }
    static ZKP_MEMBER T inv(const T &a) { return fp_inv(a); }
    static ZKP_MEMBER T one() { return fp_one(); }
    static ZKP_MEMBER T zero() { return fp_zero(); }
    static ZKP_MEMBER bool is_zero(const T &a) { return fp_is_zero(a); }
};
struct OpsFp2 {
    typedef Fp2 T;
    static ZKP_MEMBER T add(const T &a, const T &b) { return fp2_add(a, b); }
    static ZKP_MEMBER T sub(const T &a, const T &b) { return fp2_sub(a, b); }
    static ZKP_MEMBER T mul(const T &a, const T &b) { return fp2_mul(a, b); }
    static ZKP_MEMBER T sqr(const T &a) { return fp2_sqr(a); }
    static ZKP_MEMBER T inv(const T &a) { return fp2_inv(a); }
    static ZKP_MEMBER T one() { return fp2_one(); }
    static ZKP_MEMBER T zero() { return fp2_zero(); }
    static ZKP_MEMBER bool is_zero(const T &a) { return fp2_is_zero(a); }
};

template <class O>
struct Jac {
    typename O::T x, y, z;   // z == 0 <=> infinity
};
// dbl-2009-l
template <class O>
ZKP_NOINLINE void jac_double(Jac<O> &r, const Jac<O> &p) {
    typedef typename O::T T;
    T a = O::sqr(p.x), b = O::sqr(p.y), c = O::sqr(b);
    T d = O::sub(O::sub(O::sqr(O::add(p.x, b)), a), c);
    d = O::add(d, d);
    T e = O::add(O::add(a, a), a);
    T f = O::sqr(e);
    T z3 = O::mul(p.y, p.z);
    z3 = O::add(z3, z3);
    T x3 = O::sub(f, O::add(d, d));
    T c8 = O::add(c, c);
    c8 = O::add(c8, c8);
    c8 = O::add(c8, c8);
    r.y = O::sub(O::mul(e, O::sub(d, x3)), c8);
    r.x = x3;
    r.z = z3;
}
// mixed addition (madd-2007-bl); q affine, not infinity.  Handles p = inf, p = q, p = -q.
template <class O>
ZKP_NOINLINE void jac_add_affine(Jac<O> &r, const Jac<O> &p, const typename O::T &qx, const typename O::T &qy) {
    typedef typename O::T T;
    if (O::is_zero(p.z)) { r.x = qx; r.y = qy; r.z = O::one(); return; }
    T z1z1 = O::sqr(p.z);
    T u2 = O::mul(qx, z1z1);
    T s2 = O::mul(O::mul(qy, p.z), z1z1);
    T h = O::sub(u2, p.x);
    T rr = O::sub(s2, p.y);
    if (O::is_zero(h)) {
        if (O::is_zero(rr)) { Jac<O> t; t.x = qx; t.y = qy; t.z = O::one(); jac_double(r, t); return; }
        r.x = O::zero(); r.y = O::one(); r.z = O::zero(); return;
    }
    rr = O::add(rr, rr);
    T hh = O::sqr(h);
    T i = O::add(hh, hh);
    i = O::add(i, i);
    T j = O::mul(h, i);
    T v = O::mul(p.x, i);
    T x3 = O::sub(O::sub(O::sqr(rr), j), O::add(v, v));
    T t = O::mul(p.y, j);
    r.y = O::sub(O::mul(rr, O::sub(v, x3)), O::add(t, t));
    r.z = O::sub(O::sub(O::sqr(O::add(p.z, h)), z1z1), hh);
    r.x = x3;
}
// [k]Q in Jacobian coordinates, MSB-first double-and-add over `nbits` scalar bits
template <class O>
ZKP_HD void scalar_mul_jac(Jac<O> &acc, const typename O::T &qx, const typename O::T &qy, const uint64_t *k, int nbits) {
    acc.x = O::zero(); acc.y = O::one(); acc.z = O::zero();
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        jac_double(acc, acc);
        if ((k[i >> 6] >> (i & 63)) & 1) jac_add_affine(acc, acc, qx, qy);
    }
}
// [k]Q as an affine point (ax, ay); returns the infinity flag
template <class O>
ZKP_HD bool scalar_mul_affine(typename O::T &ax, typename O::T &ay, const typename O::T &qx, const typename O::T &qy,
                              const uint64_t *k, int nbits) {
    typedef typename O::T T;
    Jac<O> acc;
    scalar_mul_jac<O>(acc, qx, qy, k, nbits);
    if (O::is_zero(acc.z)) { ax = O::zero(); ay = O::one(); return true; }   // identity = (0,1,inf) src/g1.rs:25-31
    T zi = O::inv(acc.z);
    T zi2 = O::sqr(zi);
    ax = O::mul(acc.x, zi2);
    ay = O::mul(acc.y, O::mul(zi2, zi));
    return false;
}
// Jacobian p == -(affine (x, y)) without an inversion: X == x Z^2 and Y == -y Z^3 (false at infinity)
template <class O>
ZKP_HD bool jac_equals_neg_affine(const Jac<O> &p, const typename O::T &x, const typename O::T &y) {
    typedef typename O::T T;
    if (O::is_zero(p.z)) return false;
    T zz = O::sqr(p.z);
    bool ex = O::is_zero(O::sub(p.x, O::mul(x, zz)));
    bool ey = O::is_zero(O::add(p.y, O::mul(y, O::mul(zz, p.z))));
    return ex & ey;
}

}  // namespace zkp
