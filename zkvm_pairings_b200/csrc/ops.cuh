// Boundary marshalling + per-element operation bodies shared by the CUDA kernels (kernels.cu) and
// the CPU dev-simulation used by the not-gpu tests (tests/host_sim/sim.cpp).
//
// Every routine here is executed by a LANE PAIR (two adjacent threads, tower.cuh): both lanes make
// the same calls; Fp2-typed data is split between them, Fp-typed data is replicated.
//
// Boundary layout (host and device buffers): canonical little-endian u64 limbs, array-of-structs,
// exactly `Fp.0` of the reference (/root/reference/src/fp.rs:24): Fp = 6 u64, Fp2 = (c0,c1),
// Fp6 = (c0,c1,c2), Fp12 = (c0,c1) = 72 u64; G1 = x|y, G2 = x.c0|x.c1|y.c0|y.c1.
#pragma once
#include "pairing.cuh"

namespace zkp {

// op codes of the element-wise tower entry point (zkp_tower_op_batch)
enum TowerOp {
    OP_FP_ADD = 0, OP_FP_SUB, OP_FP_NEG, OP_FP_MUL, OP_FP_SQR, OP_FP_INV, OP_FP_POW, OP_FP_SQRT,
    OP_FP2_ADD = 16, OP_FP2_SUB, OP_FP2_NEG, OP_FP2_MUL, OP_FP2_SQR, OP_FP2_INV, OP_FP2_MUL_NR, OP_FP2_CONJ, OP_FP2_POW,
    OP_FP6_ADD = 32, OP_FP6_SUB, OP_FP6_NEG, OP_FP6_MUL, OP_FP6_SQR, OP_FP6_INV, OP_FP6_MUL_NR, OP_FP6_FROB, OP_FP6_MUL_BY_1, OP_FP6_MUL_BY_01,
    OP_FP12_ADD = 48, OP_FP12_SUB, OP_FP12_NEG, OP_FP12_MUL, OP_FP12_SQR, OP_FP12_INV, OP_FP12_CONJ, OP_FP12_FROB, OP_FP12_MUL_BY_014, OP_FP12_CYC_SQR, OP_FP12_CYC_EXP,
    OP_FP12_FROB2 = 59, OP_FP12_FROB3 = 60, OP_FP12_POW = 61
};

// number of Fp in operand a / operand b / result for an op (0 = operand unused)
ZKP_HOSTDEV void tower_op_shape(int op, int &na, int &nb, int &nr) {
    int w = op < 16 ? 1 : op < 32 ? 2 : op < 48 ? 6 : 12;
    na = w; nr = w; nb = 0;
    switch (op) {
        case OP_FP_ADD: case OP_FP_SUB: case OP_FP_MUL:
        case OP_FP2_ADD: case OP_FP2_SUB: case OP_FP2_MUL:
        case OP_FP6_ADD: case OP_FP6_SUB: case OP_FP6_MUL:
        case OP_FP12_ADD: case OP_FP12_SUB: case OP_FP12_MUL: nb = w; break;
        case OP_FP_POW: case OP_FP2_POW: case OP_FP12_POW: nb = 1; break;   // b = the exponent: six RAW u64 limbs
        case OP_FP6_MUL_BY_1: nb = 2; break;
        case OP_FP6_MUL_BY_01: nb = 4; break;
        case OP_FP12_MUL_BY_014: nb = 6; break;
        default: break;
    }
}

// canonical u64 limbs -> Montgomery Fp.  Sets bad when the value is >= p.
ZKP_NOINLINE Fp load_fp(const uint64_t *src, bool &bad) {
    uint32_t w[12];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint64_t x = src[i];
        w[2 * i] = (uint32_t)x;
        w[2 * i + 1] = (uint32_t)(x >> 32);
    }
    return fp_from_words(w, bad);
}
// Montgomery Fp -> canonical u64 limbs; returns the OR of all output words except the lowest and
// writes the lowest to *low (so callers can test for 0 / 1 without another conversion)
ZKP_NOINLINE uint32_t store_fp(uint64_t *dst, Fp m, uint32_t *low, bool live = true) {
    uint32_t w[12];
    fp_to_words(w, m);
    uint32_t rest = 0;
    if (live) {
#pragma unroll
        for (int i = 0; i < 6; i++) dst[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    }
#pragma unroll
    for (int i = 1; i < 12; i++) rest |= w[i];
    if (low) *low = w[0];
    return rest;
}
// Lane-split marshalling: an Fp2 occupies 12 u64 (c0 | c1); the even lane moves c0, the odd lane c1.
ZKP_HD Fp2 load_fp2(const uint64_t *src, bool &bad) { Fp2 r; r.c = load_fp(src + 6 * lane_par(), bad); return r; }
ZKP_HD uint32_t store_fp2(uint64_t *dst, const Fp2 &a, uint32_t *low, bool live = true) { return store_fp(dst + 6 * lane_par(), a.c, low, live); }
ZKP_HD void load_fp2s(Fp2 *dst, const uint64_t *src, int n, bool &bad) {
    for (int i = 0; i < n; i++) dst[i] = load_fp2(src + 12 * i, bad);
}
ZKP_HD void store_fp2s(uint64_t *dst, const Fp2 *src, int n) {
    for (int i = 0; i < n; i++) store_fp2(dst + 12 * i, src[i], nullptr);
}

// One element of a batched tower op, executed by a lane pair.  a/b/out point at this element's
// limbs.  Returns a status byte (identical in both lanes): bit0 = non-canonical input, bit1 =
// inverse of zero requested (result is zero).
ZKP_HD uint8_t tower_op_one(int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    int na, nb, nr;
    tower_op_shape(op, na, nb, nr);
    bool bad = false, noinv = false;
    if (op < 16) {   // Fp-level: both lanes compute the same value, the even lane stores it
        Fp x = load_fp(a, bad), y = fp_zero(), r = fp_zero();
        if (nb && op != OP_FP_POW) y = load_fp(b, bad);
        switch (op) {
            case OP_FP_POW: r = fp_pow(x, b); break;
            case OP_FP_SQRT: {   // src/fp.rs:280-300: a^((p+1)/4), Err when that is not a root
                r = fp_pow_const(x, ZKP_SQRT_EXP);
                noinv = !fp_is_zero(fp_sub(fsqr(r), x));
            } break;
            case OP_FP_ADD: r = fp_add(x, y); break;
            case OP_FP_SUB: r = fp_sub(x, y); break;
            case OP_FP_NEG: r = fp_neg(x); break;
            case OP_FP_MUL: r = fmul(x, y); break;
            case OP_FP_SQR: r = fsqr(x); break;
            case OP_FP_INV: noinv = fp_is_zero(x); r = fp_inv(x); break;
            default: bad = true; break;
        }
        if (lane_par() == 0) store_fp(out, r, nullptr);
        return (uint8_t)((bad ? 1 : 0) | (noinv ? 2 : 0));
    }
    Fp12 A, B, R;
    Fp2 *a2 = &A.c0.c0, *b2 = &B.c0.c0, *r2 = &R.c0.c0;
    load_fp2s(a2, a, na / 2, bad);
    const bool is_pow = op == OP_FP2_POW || op == OP_FP12_POW;
    if (nb && !is_pow) load_fp2s(b2, b, nb / 2, bad);
    switch (op) {
        case OP_FP2_POW: {       // src/fp2.rs:301-313
            Fp2 res = fp2_one();
            for (int i = 383; i >= 0; i--) {
                res = fp2_sqr(res);
                if ((b[i >> 6] >> (i & 63)) & 1) res = fp2_mul(res, a2[0]);
            }
            r2[0] = res;
        } break;
        case OP_FP12_POW: {      // src/fp12.rs:127-139
            fp12_set_one(R);
            for (int i = 383; i >= 0; i--) {
                fp12_sqr(R, R);
                if ((b[i >> 6] >> (i & 63)) & 1) fp12_mul(R, R, A);
            }
        } break;
        case OP_FP2_ADD: r2[0] = fp2_add(a2[0], b2[0]); break;
        case OP_FP2_SUB: r2[0] = fp2_sub(a2[0], b2[0]); break;
        case OP_FP2_NEG: r2[0] = fp2_neg(a2[0]); break;
        case OP_FP2_MUL: r2[0] = fp2_mul(a2[0], b2[0]); break;
        case OP_FP2_SQR: r2[0] = fp2_sqr(a2[0]); break;
        case OP_FP2_INV: noinv = fp2_is_zero(a2[0]); r2[0] = fp2_inv(a2[0]); break;
        case OP_FP2_MUL_NR: r2[0] = fp2_mul_nr(a2[0]); break;
        case OP_FP2_CONJ: r2[0] = fp2_conj(a2[0]); break;
        case OP_FP6_ADD: fp6_add(R.c0, A.c0, B.c0); break;
        case OP_FP6_SUB: fp6_sub(R.c0, A.c0, B.c0); break;
        case OP_FP6_NEG: fp6_neg(R.c0, A.c0); break;
        case OP_FP6_MUL: fp6_mul(R.c0, A.c0, B.c0); break;
        case OP_FP6_SQR: fp6_sqr(R.c0, A.c0); break;
        case OP_FP6_INV: noinv = fp2_is_zero(A.c0.c0) & fp2_is_zero(A.c0.c1) & fp2_is_zero(A.c0.c2); fp6_inv(R.c0, A.c0); break;
        case OP_FP6_MUL_NR: fp6_mul_nr(R.c0, A.c0); break;
        case OP_FP6_FROB: {   // Fp6 embedded as c0 of an Fp12: frobenius acts coefficient-wise
            fp6_set_zero(A.c1);
            fp12_frobenius(R, A, 1);
        } break;
        case OP_FP6_MUL_BY_1: fp6_mul_by_1(R.c0, A.c0, b2[0]); break;
        case OP_FP6_MUL_BY_01: fp6_mul_by_01(R.c0, A.c0, b2[0], b2[1]); break;
        case OP_FP12_ADD: fp6_add(R.c0, A.c0, B.c0); fp6_add(R.c1, A.c1, B.c1); break;
        case OP_FP12_SUB: fp6_sub(R.c0, A.c0, B.c0); fp6_sub(R.c1, A.c1, B.c1); break;
        case OP_FP12_NEG: fp6_neg(R.c0, A.c0); fp6_neg(R.c1, A.c1); break;
        case OP_FP12_MUL: fp12_mul(R, A, B); break;
        case OP_FP12_SQR: fp12_sqr(R, A); break;
        case OP_FP12_INV: {
            bool z = true;
            for (int i = 0; i < 6; i++) z = z & fp2_is_zero(a2[i]);
            noinv = z;
            fp12_inv(R, A);
        } break;
        case OP_FP12_CONJ: fp12_conj(R, A); break;
        case OP_FP12_FROB: fp12_frobenius(R, A, 1); break;
        case OP_FP12_FROB2: fp12_frobenius(R, A, 2); break;
        case OP_FP12_FROB3: fp12_frobenius(R, A, 3); break;
        case OP_FP12_MUL_BY_014: R = A; fp12_mul_by_014(R, b2[0], b2[1], b2[2]); break;
        case OP_FP12_CYC_SQR: fp12_cyclotomic_sqr(R, A); break;
        case OP_FP12_CYC_EXP: cyclotomic_exp(R, A); break;
        default: bad = true; nr = 0; break;
    }
    store_fp2s(out, r2, nr / 2);
    bad = lane_or(bad);
    return (uint8_t)((bad ? 1 : 0) | (noinv ? 2 : 0));
}

ZKP_HD void load_g1(G1A &p, const uint64_t *xy, bool &bad) { p.x = load_fp(xy, bad); p.y = load_fp(xy + 6, bad); }
ZKP_HD void load_g2(G2A &q, const uint64_t *xy, bool &bad) { q.x = load_fp2(xy, bad); q.y = load_fp2(xy + 12, bad); }
// stores f and returns true (in both lanes) when it equals Fp12::one() (canonical 1, 0, ..., 0)
ZKP_HD bool store_fp12(uint64_t *dst, const Fp12 &f, bool live = true) {
    const Fp2 *c = &f.c0.c0;
    uint32_t low = 0, rest = 0;
    rest |= store_fp2(dst, c[0], &low, live);
    bool first = lane_par() == 0 ? (low == 1u) : (low == 0u);
    for (int i = 1; i < 6; i++) {
        uint32_t l2 = 0;
        rest |= store_fp2(dst + 12 * i, c[i], &l2, live);
        rest |= l2;
    }
    return lane_and(first & (rest == 0));
}
ZKP_HD void load_fp12(Fp12 &f, const uint64_t *src, bool &bad) { load_fp2s(&f.c0.c0, src, 6, bad); }

// mode bits for pairing_one
// (ZKP_FREE_LINE_SCALING: the Miller output is only consumed by a later final exponentiation, so its lines may be scaled)
enum { ZKP_DO_MILLER = 1, ZKP_DO_FINAL_EXP = 2, ZKP_FREE_LINE_SCALING = 4 };

// One "check": k pairs -> shared-accumulator Miller loop (-> final exponentiation).  g1/g2/inf
// point at this check's first pair.  Returns status bit0 = non-canonical input.  Control flow is
// lane-uniform (k and mode are kernel arguments); `live` = false suppresses the stores only.
// is_one (optional) receives 1 when the result equals Fp12::one().  K = compile-time capacity
// (k <= K) so the per-thread scratch is sized for the common k = 1 case.
template <int K>
ZKP_HD void pairing_front(Fp12 &f, bool &bad, int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2,
                          const uint8_t *g2inf, int k, const uint64_t *in12, const Fp *tab = nullptr,
                          const uint8_t *tabinf = nullptr, int kf = 0, G2P *rs_ext = nullptr, Fp6 *tmp = nullptr) {
    // k pairs in total; the last kf of them take their G2 lines from the prepared tables `tab`
    // (then g2 / g2inf hold only the k - kf per-check points)
    if (mode & ZKP_DO_MILLER) {
        G1A ps[K];
        G2A qs[K];
        G2P rs_local[K];
        G2P *rs = rs_ext ? rs_ext : rs_local;   // the caller may keep the G2 accumulators elsewhere (shared memory)
        bool skip[K];
        const int kv = k - kf;
        for (int j = 0; j < k; j++) {
            load_g1(ps[j], g1 + 12 * j, bad);
            skip[j] = g1inf && g1inf[j];
        }
        for (int j = 0; j < kv; j++) {
            load_g2(qs[j], g2 + 24 * j, bad);
            skip[j] = skip[j] | (g2inf && g2inf[j]);
        }
        for (int j = 0; j < kf; j++) skip[kv + j] = skip[kv + j] | (tabinf && tabinf[j]);
#if ZKP_INPLACE12
        Fp6 tmp_local;   // the one Fp6 temporary of the in-place Fp12 operations when the caller does not supply one
        if (!tmp) tmp = &tmp_local;
#endif
#ifdef ZKP_NO_FUSED_LINES
        const bool fused = false;
#else
        const bool fused = (mode & (ZKP_DO_FINAL_EXP | ZKP_FREE_LINE_SCALING)) != 0;   // only Gt is observable: line scaling is free (pairing.cuh)
#endif
        miller_loop(f, ps, qs, skip, rs, kv, tab, kf, tmp, fused);
    } else {
        load_fp12(f, in12, bad);
    }
}
// the whole check in one call (dev simulation and small helpers; the GPU path splits the final
// exponentiation over 13 launches, pairing_kernel.cu)
template <int K>
ZKP_HD uint8_t pairing_one(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                           int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one, bool live = true) {
    bool bad = false;
    Fp12 f;
    pairing_front<K>(f, bad, mode, g1, g1inf, g2, g2inf, k, in12);
    if (mode & ZKP_DO_FINAL_EXP) final_exponentiation(f, f);
    bool one = store_fp12(out, f, live);
    if (is_one && live && lane_par() == 0) *is_one = one ? 1 : 0;
    return lane_or(bad) ? 1 : 0;
}

// SplitMix64 output number idx+1 of the stream seeded with `seed` (random access)
ZKP_HD uint64_t splitmix64_at(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// [k]G for the G1 / G2 generators, k = 64-bit scalar (synthetic input generation, untimed)
ZKP_HD void gen_g1_one(uint64_t k, uint64_t *xy, uint8_t *inf) {
    Fp gx = fp_const(ZKP_G1_GEN), gy = fp_const(ZKP_G1_GEN + ZKP_NL), ax, ay;
    bool is_inf = scalar_mul_affine<OpsFp>(ax, ay, gx, gy, &k, 64);   // both lanes hold the same point
    if (lane_par() == 0) {
        *inf = is_inf ? 1 : 0;
        store_fp(xy, ax, nullptr);
    } else {
        store_fp(xy + 6, ay, nullptr);
    }
}
ZKP_HD void gen_g2_one(uint64_t k, uint64_t *xy, uint8_t *inf) {
    Fp2 gx = fp2_const(ZKP_G2_GEN), gy = fp2_const(ZKP_G2_GEN + 2 * ZKP_NL), ax, ay;
    bool is_inf = scalar_mul_affine<OpsFp2>(ax, ay, gx, gy, &k, 64);
    if (lane_par() == 0) *inf = is_inf ? 1 : 0;
    store_fp2(xy, ax, nullptr);
    store_fp2(xy + 12, ay, nullptr);
}

// ------------------------------------------------------------------ group-level batch ops

enum GroupOp { GOP_G1_CHECK = 0, GOP_G2_CHECK = 1, GOP_G1_MUL = 2, GOP_G2_MUL = 3, GOP_G1_ADD = 4, GOP_G2_ADD = 5 };
#ifndef ZKP_POINT_OK   /* same values as include/zkpair.h */
#define ZKP_POINT_OK 0
#define ZKP_POINT_NOT_ON_CURVE 1
#define ZKP_POINT_NOT_TORSION_FREE 2
#endif

// |x|^2 (128 bits): G1 subgroup test -[x^2]P == (beta x, y), src/g1.rs:103-115
ZKP_HD uint8_t g1_check_one(const uint64_t *xy, uint8_t inf, bool &bad) {
    Fp x = load_fp(xy, bad), y = load_fp(xy + 6, bad);
    if (inf) return ZKP_POINT_OK;                                             // src/g1.rs:50-52
    Fp rhs = fp_add(fmul(fsqr(x), x), fp_const(ZKP_B1));                      // y^2 = x^3 + 4, src/g1.rs:95-101
    if (!fp_is_zero(fp_sub(fsqr(y), rhs))) return ZKP_POINT_NOT_ON_CURVE;
    const uint64_t xx[2] = {(uint64_t)(ZKP_BLS_X * ZKP_BLS_X), (uint64_t)(((unsigned __int128)ZKP_BLS_X * ZKP_BLS_X) >> 64)};
    Jac<OpsFp> acc;
    scalar_mul_jac<OpsFp>(acc, x, y, xx, 128);
    // -[x^2]P == (beta x, y)  <=>  [x^2]P == -(beta x, y)
    return jac_equals_neg_affine<OpsFp>(acc, fmul(x, fp_const(ZKP_BETA)), y) ? ZKP_POINT_OK : ZKP_POINT_NOT_TORSION_FREE;
}
// G2 subgroup test psi(Q) == -[|x|]Q, src/g2.rs:126-170
ZKP_HD uint8_t g2_check_one(const uint64_t *xy, uint8_t inf, bool &bad) {
    Fp2 x = load_fp2(xy, bad), y = load_fp2(xy + 12, bad);
    if (inf) return ZKP_POINT_OK;                                             // src/g2.rs:58-60
    Fp2 b2;
    b2.c = fp_const(ZKP_B1);                                                  // b' = 4 + 4u, src/common.rs:70-71
    Fp2 rhs = fp2_add(fp2_mul(fp2_sqr(x), x), b2);
    if (!fp2_is_zero(fp2_sub(fp2_sqr(y), rhs))) return ZKP_POINT_NOT_ON_CURVE;
    const uint64_t k[1] = {ZKP_BLS_X};
    Jac<OpsFp2> acc;
    scalar_mul_jac<OpsFp2>(acc, x, y, k, 64);
    Fp2 px = fp2_mul(fp2_conj(x), fp2_const(ZKP_PSI));
    Fp2 py = fp2_mul(fp2_conj(y), fp2_const(ZKP_PSI + 2 * ZKP_NL));
    return jac_equals_neg_affine<OpsFp2>(acc, px, py) ? ZKP_POINT_OK : ZKP_POINT_NOT_TORSION_FREE;
}
// [k]P for a 256-bit scalar k (4 little-endian u64, the limbs of an Fr, src/fr.rs).  Correct
// double-and-add over all 256 bits (the reference's G1 `Mul<&Fr>` drops bit 0, src/g1.rs:138-142).
ZKP_HD void g1_mul_one(const uint64_t *xy, uint8_t inf, const uint64_t *k, uint64_t *out_xy, uint8_t *out_inf, bool &bad) {
    Fp x = load_fp(xy, bad), y = load_fp(xy + 6, bad), ax, ay;
    bool is_inf = true;
    if (inf) { ax = fp_zero(); ay = fp_one(); }
    else is_inf = scalar_mul_affine<OpsFp>(ax, ay, x, y, k, 256);
    if (lane_par() == 0) {
        *out_inf = is_inf ? 1 : 0;
        store_fp(out_xy, ax, nullptr);
    } else {
        store_fp(out_xy + 6, ay, nullptr);
    }
}
ZKP_HD void g2_mul_one(const uint64_t *xy, uint8_t inf, const uint64_t *k, uint64_t *out_xy, uint8_t *out_inf, bool &bad) {
    Fp2 x = load_fp2(xy, bad), y = load_fp2(xy + 12, bad), ax, ay;
    bool is_inf = true;
    if (inf) { ax = fp2_zero(); ay = fp2_one(); }
    else is_inf = scalar_mul_affine<OpsFp2>(ax, ay, x, y, k, 256);
    if (lane_par() == 0) *out_inf = is_inf ? 1 : 0;
    store_fp2(out_xy, ax, nullptr);
    store_fp2(out_xy + 12, ay, nullptr);
}

// P + Q with the affine chord-and-tangent law of `&G1Affine + &G1Affine` / `&G2Affine + &G2Affine`
// (src/g1.rs:155-187 with double() :74-91; src/g2.rs:210-242 with :81-105): identity operands pass the
// other one through, equal points double, slope = (y2 - y1) / (x2 - x1) or 3 x^2 / (2 y).  Where the
// reference divides by zero and panics (P + (-P), or doubling a point with y = 0) the result is the
// identity (0, 1) and bit1 of the flag is set.  flag bit0 = the result is the identity.
template <class O>
ZKP_HD uint8_t affine_add(typename O::T &xr, typename O::T &yr, const typename O::T &x1, const typename O::T &y1, bool inf1,
                          const typename O::T &x2, const typename O::T &y2, bool inf2) {
    typedef typename O::T T;
    if (inf1 | inf2) {
        bool both = inf1 & inf2;
        xr = both ? O::zero() : (inf1 ? x2 : x1);
        yr = both ? O::one() : (inf1 ? y2 : y1);
        return both ? 1 : 0;
    }
    T dx = O::sub(x2, x1), dy = O::sub(y2, y1), num, den;
    if (O::is_zero(dx) && O::is_zero(dy)) {
        T xx = O::sqr(x1);
        num = O::add(O::add(xx, xx), xx);
        den = O::add(y1, y1);
    } else {
        num = dy;
        den = dx;
    }
    if (O::is_zero(den)) {
        xr = O::zero();
        yr = O::one();
        return 3;
    }
    T slope = O::mul(num, O::inv(den));
    xr = O::sub(O::sub(O::sqr(slope), x1), x2);
    yr = O::sub(O::mul(slope, O::sub(x1, xr)), y1);
    return 0;
}
ZKP_HD void g1_add_one(const uint64_t *a, uint8_t ainf, const uint64_t *b, uint8_t binf, uint64_t *out_xy, uint8_t *flag, bool &bad) {
    Fp x1 = load_fp(a, bad), y1 = load_fp(a + 6, bad), x2 = load_fp(b, bad), y2 = load_fp(b + 6, bad), xr, yr;
    uint8_t f = affine_add<OpsFp>(xr, yr, x1, y1, ainf != 0, x2, y2, binf != 0);
    if (lane_par() == 0) {
        *flag = f;
        store_fp(out_xy, xr, nullptr);
    } else {
        store_fp(out_xy + 6, yr, nullptr);
    }
}
ZKP_HD void g2_add_one(const uint64_t *a, uint8_t ainf, const uint64_t *b, uint8_t binf, uint64_t *out_xy, uint8_t *flag, bool &bad) {
    Fp2 x1 = load_fp2(a, bad), y1 = load_fp2(a + 12, bad), x2 = load_fp2(b, bad), y2 = load_fp2(b + 12, bad), xr, yr;
    uint8_t f = affine_add<OpsFp2>(xr, yr, x1, y1, ainf != 0, x2, y2, binf != 0);
    if (lane_par() == 0) *flag = f;
    store_fp2(out_xy, xr, nullptr);
    store_fp2(out_xy + 12, yr, nullptr);
}

}  // namespace zkp
