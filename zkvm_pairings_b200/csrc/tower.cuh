// Fp2 / Fp6 / Fp12 tower for BLS12-381 on sm_100a.
//
// Device re-design of the reference tower (/root/reference/src/fp2.rs, fp6.rs, fp12.rs).  Same
// field definitions -- Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(u+1)), Fp12 = Fp6[w]/(w^2-v) -- and
// therefore bit-identical canonical results, but Karatsuba at every level (Fp2 mul = 3 Fp mul,
// Fp6 mul = 6 Fp2 mul, Fp12 mul = 3 Fp6 mul) instead of the reference's schoolbook Fp2
// (src/fp2.rs:192-209) and 36-mul interleaved Fp6 (src/fp6.rs:188-267).
//
// Work split: TWO lanes per pairing.  Every Fp2 is split over an even/odd lane pair (see the Fp2
// section), so a thread carries half of the tower state (an Fp12 is 6 x 12 registers per lane)
// and the Fp6-level working sets fit the register file.
//
// Code-size / register strategy: the Fp2 product and square (444 / 300 wide MACs per lane) are
// out-of-line functions whose operands and result travel in registers (24 in, 12 out), so their
// bodies stay hot in the instruction cache.  Additions are inlined; Fp6-level and larger
// operations are out-of-line and exchange operands through thread-local memory.
//
// Reduction discipline: every Fp is kept 2p-redundant (fp.cuh): additions and subtractions correct
// by +-2p, products come back below 2p, so any two values can be multiplied without further
// thought.  The one lazy spot is the Fp2 squaring (one operand is an uncorrected sum <= 4p).
#pragma once
#include "fp.cuh"

// 1: the Miller loop squares f and multiplies it by a line IN PLACE with one Fp6 temporary (fp12_sqr_inplace,
// fp12_mul_by_014_inplace below) instead of three thread-local ones per operation; 0: the out-of-place forms only.
// Measured (profiles/r2d_miller_ab.txt): Miller loop at 2^20 276.4 -> 273.7 ms with f and the temporary in shared memory.
#ifndef ZKP_INPLACE12
#define ZKP_INPLACE12 1
#endif

// Multiply-pipe token (pairing_kernel.cu, tools/ring_probe.cu): the MAC-dense bodies below are bracketed by these two
// hooks; a kernel may define them so that the warps sharing a scheduler enter the bodies one after the other (FIFO)
// instead of time-slicing the multiply pipe.  Default: nothing.
#ifndef ZKP_PIPE_ACQUIRE
#define ZKP_PIPE_ACQUIRE() do { } while (0)
#define ZKP_PIPE_RELEASE() do { } while (0)
#endif

namespace zkp {

// ------------------------------------------------------------------ out-of-line Fp kernels
ZKP_NOINLINE Fp fmul(Fp a, Fp b) {
    ZKP_PIPE_ACQUIRE();
    Fp r = fp_mul(a, b);
    ZKP_PIPE_RELEASE();
    return r;
}
#define fsqr(a) fmul((a), (a))

// 1/a mod p (the value of Fp::invert, src/fp.rs:306-319, which walks the exponent p - 2, src/fp.rs:264-276; zero maps
// to zero, callers flag it) by the binary extended GCD instead of that Fermat ladder: 761 branch-free iterations of shifts, subtractions and selects
// on the ALU pipe -- no multiplications at all -- so the inversion neither occupies the multiply pipe nor
// waits on its latency: ~0.14 M ALU instructions with a ~0.15 k-cycle dependent chain per iteration against
// 608 dependent Montgomery products (0.18 M wide MACs) for the ladder.  This is what makes the six batched
// inversion launches of the staged final exponentiation cheap at small batches (fe_kernel.cu).
//   invariants: a = u y, b = v y (mod p); a odd step: a <- |a - b| / 2, b <- min(a, b); even step: a <- a / 2;
//   len(a) + len(b) drops every iteration, so 2 * 381 - 1 iterations end with a = 0, b = gcd = 1, v = 1/y.
// The operand is a Montgomery representative x = X R in [0, 2p]; y = x mod p, and (X R)^-1 R^3 / R = X^-1 R.
ZKP_NOINLINE Fp fp_inv(Fp x) {
    uint32_t a[ZKP_NL], b[ZKP_NL], u[ZKP_NL], v[ZKP_NL];
#pragma unroll 1
    for (int r = 0; r < 2; r++) {   // [0, 2p] -> [0, p)
        Fp d;
        d.l[0] = sub_cc(x.l[0], ZKP_P[0]);
#pragma unroll
        for (int i = 1; i < ZKP_NL; i++) d.l[i] = subc_cc(x.l[i], ZKP_P[i]);
        bool ge = subc(0, 0) == 0;
        x = fp_select(ge, d, x);
    }
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        a[i] = x.l[i];
        b[i] = ZKP_P[i];
        u[i] = i == 0 ? 1u : 0u;
        v[i] = 0;
    }
#pragma unroll 1
    for (int it = 0; it < 2 * 381 - 1; it++) {
        const bool odd = (a[0] & 1u) != 0;
        uint32_t d[ZKP_NL], nd[ZKP_NL], w[ZKP_NL];
        d[0] = sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < ZKP_NL; i++) d[i] = subc_cc(a[i], b[i]);
        const bool lt = subc(0, 0) != 0;   // a < b
        const bool sw = odd & lt;          // the pair (a, u) <-> (b, v) changes roles
        nd[0] = sub_cc(0, d[0]);
#pragma unroll
        for (int i = 1; i < ZKP_NL - 1; i++) nd[i] = subc_cc(0, d[i]);
        nd[ZKP_NL - 1] = subc(0, d[ZKP_NL - 1]);
#pragma unroll
        for (int i = 0; i < ZKP_NL; i++) {
            uint32_t na = odd ? (lt ? nd[i] : d[i]) : a[i];
            b[i] = sw ? a[i] : b[i];
            a[i] = na;
        }
#pragma unroll
        for (int i = 0; i < ZKP_NL - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[ZKP_NL - 1] >>= 1;
#pragma unroll
        for (int i = 0; i < ZKP_NL; i++) {
            uint32_t tu = sw ? v[i] : u[i];
            v[i] = sw ? u[i] : v[i];
            u[i] = tu;
        }
        // u <- odd ? (u - v mod p) : u
        w[0] = sub_cc(u[0], v[0]);
#pragma unroll
        for (int i = 1; i < ZKP_NL; i++) w[i] = subc_cc(u[i], v[i]);
        const uint32_t mneg = subc(0, 0) != 0 ? 0xffffffffu : 0u;
        w[0] = add_cc(w[0], ZKP_P[0] & mneg);
#pragma unroll
        for (int i = 1; i < ZKP_NL - 1; i++) w[i] = addc_cc(w[i], ZKP_P[i] & mneg);
        w[ZKP_NL - 1] = addc(w[ZKP_NL - 1], ZKP_P[ZKP_NL - 1] & mneg);
#pragma unroll
        for (int i = 0; i < ZKP_NL; i++) u[i] = odd ? w[i] : u[i];
        // u <- u / 2 mod p: (u + p) / 2 when u is odd; u + p < 2^382, no carry out of the top word
        const uint32_t modd = (u[0] & 1u) ? 0xffffffffu : 0u;
        w[0] = add_cc(u[0], ZKP_P[0] & modd);
#pragma unroll
        for (int i = 1; i < ZKP_NL - 1; i++) w[i] = addc_cc(u[i], ZKP_P[i] & modd);
        w[ZKP_NL - 1] = addc(u[ZKP_NL - 1], ZKP_P[ZKP_NL - 1] & modd);
#pragma unroll
        for (int i = 0; i < ZKP_NL - 1; i++) u[i] = (w[i] >> 1) | (w[i + 1] << 31);
        u[ZKP_NL - 1] = w[ZKP_NL - 1] >> 1;
    }
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = v[i];
    return fmul(r, fp_const(ZKP_R3));
}

// a^e for a 384-bit exponent given as six little-endian u64 limbs: Fp::pow_vartime, src/fp.rs:264-276
ZKP_NOINLINE Fp fp_pow(Fp a, const uint64_t *e) {
    Fp res = fp_one();
    for (int i = 383; i >= 0; i--) {
        res = fsqr(res);
        if ((e[i >> 6] >> (i & 63)) & 1) res = fmul(res, a);
    }
    return res;
}
// same with the exponent as twelve constant 32-bit words
ZKP_NOINLINE Fp fp_pow_const(Fp a, const uint32_t *e) {
    Fp res = fp_one();
    for (int i = 383; i >= 0; i--) {
        res = fsqr(res);
        if ((e[i >> 5] >> (i & 31)) & 1) res = fmul(res, a);
    }
    return res;
}

// ------------------------------------------------------------------ Fp2 (split over a lane pair)
//
// An Fp2 value a0 + a1*u lives in TWO adjacent lanes: the even lane holds a0, the odd lane a1.
// Additions are lane-local; a product costs each lane three 12-word shuffles and one lazy
// "two products, one reduction" (fp.cuh mont_mul2):
//     even lane:  c0 = a0*b0 - a1*b1            odd lane:  c1 = a0*b1 + a1*b0
// i.e. the schoolbook form of src/fp2.rs:192-209, which with a single reduction per lane costs
// 2*(288+156) = 888 wide MACs per Fp2 product against 3*300 = 900 for Karatsuba over full Fp
// products -- and both lanes do identical work, so there is no divergence and no idle lane.
struct Fp2 {
    Fp c;   // c0 in even lanes, c1 in odd lanes
};

ZKP_HD Fp2 fp2_zero() { Fp2 r; r.c = fp_zero(); return r; }
ZKP_HD Fp2 fp2_one() { Fp2 r; r.c = fp_select(lane_par() != 0, fp_zero(), fp_one()); return r; }
// constant stored as c0 | c1 (2 x 12 words, Montgomery form)
ZKP_HD Fp2 fp2_const(const uint32_t *k) { Fp2 r; r.c = fp_const(k + lane_par() * ZKP_NL); return r; }
ZKP_HD bool fp2_is_zero(const Fp2 &a) { return lane_and(fp_is_zero(a.c)); }
// The linear ops are inlined.  (Measured: making them out-of-line calls shrinks the code from 313 KB
// to 188 KB but costs 10% throughput -- 1.20 vs 1.34 M pairings/s -- in call overhead.)
#ifndef ZKP_LIN
#define ZKP_LIN ZKP_HD
#endif
#ifndef ZKP_SQR_SEL
#define ZKP_SQR_SEL 0   // 1: fp2_sqr forms a0 by a select on the exchanged value instead of a broadcast shuffle (12 SEL for 12 SHFL)
#endif
ZKP_LIN Fp2 fp2_add(Fp2 a, Fp2 b) { Fp2 r; r.c = fp_add(a.c, b.c); return r; }   // src/fp2.rs:216-218
ZKP_LIN Fp2 fp2_sub(Fp2 a, Fp2 b) { Fp2 r; r.c = fp_sub(a.c, b.c); return r; }   // src/fp2.rs:221-223
ZKP_LIN Fp2 fp2_neg(Fp2 a) { Fp2 r; r.c = fp_neg(a.c); return r; }               // src/fp2.rs:226-228
ZKP_HD Fp2 fp2_dbl(const Fp2 &a) { return fp2_add(a, a); }
// (a0, -a1)   -- src/fp2.rs:155-157
ZKP_LIN Fp2 fp2_conj(Fp2 a) { Fp2 r; r.c = fp_select(lane_par() != 0, fp_neg(a.c), a.c); return r; }
// (a + bu)(1 + u) = (a - b) + (a + b)u   -- src/fp2.rs:161-168
ZKP_LIN Fp2 fp2_mul_nr(Fp2 a) {
    Fp t = fp_xchg(a.c);
    Fp2 r;
    r.c = fp_add(a.c, fp_select(lane_par() != 0, t, fp_neg(t)));
    return r;
}
// Product; operands and result 2p-redundant.  Both lanes evaluate the SAME expression
//     own_a * b0 + t * b1,   t = the partner's a (the odd lane sends -a1, the even lane a0)
// which is a0*b0 - a1*b1 in the even lane and a1*b0 + a0*b1 in the odd lane: three 12-word
// shuffles (one exchange, two broadcasts), one lane-dependent negation, no selects.
ZKP_NOINLINE Fp2 fp2_mul(Fp2 a, Fp2 b) {
    ZKP_CODE_SYNC(5);
    bool odd = lane_par() != 0;
    Fp t = fp_xchg(fp_select(odd, fp_neg(a.c), a.c));
    Fp b0 = fp_bcast<0>(b.c), b1 = fp_bcast<1>(b.c);
    Fp2 r;
    ZKP_PIPE_ACQUIRE();
    r.c = mont_mul2(a.c, b0, t, b1);
    ZKP_PIPE_RELEASE();
    return r;
}
// Square (complex method, src/fp2.rs:171-189): even lane (a0+a1)(a0-a1), odd lane (2 a0) a1.
// x = a0 + partner's a is an uncorrected sum (<= 4p), y is 2p-redundant: x*y <= 8p^2 as mont_mul
// requires.
ZKP_NOINLINE Fp2 fp2_sqr(Fp2 a) {
    ZKP_CODE_SYNC(5);
    bool odd = lane_par() != 0;
    Fp pa = fp_xchg(a.c);
#if ZKP_SQR_SEL
    Fp x = fp_add_lazy(fp_select(odd, pa, a.c), pa);   // a0 in both lanes without a second shuffle
#else
    Fp x = fp_add_lazy(fp_bcast<0>(a.c), pa);
#endif
    Fp y = fp_select(odd, a.c, fp_sub(a.c, pa));
    Fp2 r;
    ZKP_PIPE_ACQUIRE();
    r.c = mont_mul(x, y);
    ZKP_PIPE_RELEASE();
    return r;
}
// ---- unreduced Fp2 products (fp.cuh FpW): this lane's component, at most 8 p^2, to be recombined and reduced later
// ZKP_LAZY -- lazy reduction (fp.cuh FpW): bit 0 Fp6 products, bit 1 fp6_mul_by_01, bit 2 Fp4 squares recombine UNREDUCED
// Fp2 products (3 / 3 / 2 reductions instead of 6 / 5 / 3 per lane: Miller loop -9.8 %, final exponentiation -16.1 % wide
// MACs, bit-identical results, GPU suite green with all three).  Measured at 2^20 (profiles/r2l_lazy_reduction.txt): the
// 768-bit recombinations, their spills and the larger code give back most of what the multiplier saves.  Everywhere (7):
// Miller loop 273.3 -> 269.7..271.7 ms, final exponentiation 260.5 -> 262.4..266.7 ms; in the Miller unit only (3, what
// pairing_kernel.cu ships): pairing 520.5 -> 516.4 ms, prepared 4-pair checks -2.7 %.  Default here (every other unit): 0.
#ifndef ZKP_LAZY
#define ZKP_LAZY 0
#endif
#ifndef ZKP_LAZY_ORDER
#define ZKP_LAZY_ORDER 1   // order of the sums in the lazy fp6_mul (below); 1 = fewer unreduced values alive across the last calls
#endif
ZKP_NOINLINE FpW fp2_mulw(Fp2 a, Fp2 b) {
    ZKP_CODE_SYNC(5);
    bool odd = lane_par() != 0;
    Fp t = fp_xchg(fp_select(odd, fp_neg(a.c), a.c));
    Fp b0 = fp_bcast<0>(b.c), b1 = fp_bcast<1>(b.c);
    ZKP_PIPE_ACQUIRE();
    FpW r = mul_wide2(a.c, b0, t, b1);
    ZKP_PIPE_RELEASE();
    return r;
}
ZKP_NOINLINE FpW fp2_sqrw(Fp2 a) {
    ZKP_CODE_SYNC(5);
    bool odd = lane_par() != 0;
    Fp pa = fp_xchg(a.c);
    Fp x = fp_add_lazy(fp_bcast<0>(a.c), pa);
    Fp y = fp_select(odd, a.c, fp_sub(a.c, pa));
    ZKP_PIPE_ACQUIRE();
    FpW r = mul_wide(x, y);
    ZKP_PIPE_RELEASE();
    return r;
}
// xi * X for an unreduced Fp2 value (X = this lane's component, Y = the partner's): X - Y in the even lane,
// X + Y in the odd lane, modulo 2^768; the subtraction is the two's complement (~Y + 1) so that both lanes run the
// same carry chain
ZKP_HD FpW fpw_mul_nr(const FpW &x) {
    FpW y = fpw_xchg(x);
    const uint32_t m = lane_par() != 0 ? 0u : 0xffffffffu;
    FpW r;
    add_cc(m, 1u);   // carry flag <- 1 in the even lane
#pragma unroll
    for (int i = 0; i < 2 * ZKP_NL - 1; i++) r.l[i] = addc_cc(x.l[i], y.l[i] ^ m);
    r.l[2 * ZKP_NL - 1] = addc(x.l[2 * ZKP_NL - 1], y.l[2 * ZKP_NL - 1] ^ m);
    return r;
}
// reduce a recombined value (the caller has added the multiple of p^2 that makes it non-negative).
// BIG = 0: value < 29 p^2 (result < 4p, one correction); 1: < 68 p^2 (result < 8p, two corrections)
template <int BIG>
ZKP_NOINLINE Fp2 fp2_redc(FpW t) {
    ZKP_PIPE_ACQUIRE();
    Fp r = mont_redc(t);
    ZKP_PIPE_RELEASE();
#ifndef ZKP_DEVICE_BUILD
    ZKP_SIM_ASSERT(BIG ? !fp_geq_const(r, ZKP_8P) : fp_leq_4p(r), "fp2_redc result bound");
#endif
    Fp2 o;
    o.c = BIG ? fp_correct8(r) : fp_correct(r);
    return o;
}

ZKP_HD Fp2 fp2_mul_fp(const Fp2 &a, const Fp &k) { Fp2 r; r.c = fmul(a.c, k); return r; }   // src/fp2.rs:95-102
// src/fp2.rs:278-296 ; zero maps to zero.  Split around the Fp inversion of the norm so that the
// pairing kernels can batch that inversion over many pairings (pairing_kernel.cu):
//   n = a0^2 + a1^2 (both lanes)   ->   ninv = 1/n   ->   (a0 * ninv, -a1 * ninv)
ZKP_HD Fp fp2_norm(const Fp2 &a) {
    Fp n = fsqr(a.c);
    return fp_add(n, fp_xchg(n));
}
ZKP_HD Fp2 fp2_inv_finish(const Fp2 &a, const Fp &ninv) {
    Fp m = fmul(a.c, ninv);
    Fp2 r;
    r.c = fp_select(lane_par() != 0, fp_neg(m), m);
    return r;
}
ZKP_HD Fp2 fp2_inv(const Fp2 &a) { return fp2_inv_finish(a, fp_inv(fp2_norm(a))); }

// ------------------------------------------------------------------ Fp6
struct Fp6 {
    Fp2 c0, c1, c2;
};

ZKP_HD void fp6_set_zero(Fp6 &r) { r.c0 = fp2_zero(); r.c1 = fp2_zero(); r.c2 = fp2_zero(); }
ZKP_HD void fp6_add(Fp6 &r, const Fp6 &a, const Fp6 &b) { r.c0 = fp2_add(a.c0, b.c0); r.c1 = fp2_add(a.c1, b.c1); r.c2 = fp2_add(a.c2, b.c2); }   // src/fp6.rs:322-333
ZKP_HD void fp6_sub(Fp6 &r, const Fp6 &a, const Fp6 &b) { r.c0 = fp2_sub(a.c0, b.c0); r.c1 = fp2_sub(a.c1, b.c1); r.c2 = fp2_sub(a.c2, b.c2); }   // src/fp6.rs:358-367
ZKP_HD void fp6_neg(Fp6 &r, const Fp6 &a) { r.c0 = fp2_neg(a.c0); r.c1 = fp2_neg(a.c1); r.c2 = fp2_neg(a.c2); }                                   // src/fp6.rs:335-346
// times v: (c0,c1,c2) -> (xi*c2, c0, c1)  -- src/fp6.rs:128-139 ; r may alias a
ZKP_HD void fp6_mul_nr(Fp6 &r, const Fp6 &a) {
    Fp2 t = fp2_mul_nr(a.c2);
    r.c2 = a.c1;
    r.c1 = a.c0;
    r.c0 = t;
}
// Karatsuba, 6 Fp2 mul (value-equal to mul_interleaved, src/fp6.rs:188-267); r may alias a or b
#if ZKP_LAZY & 1
// Lazy reduction: the six products stay unreduced (v, w <= 8 p^2 per lane) and are recombined as 768-bit integers;
// three reductions per lane instead of six (2196 wide MACs instead of 2664).  Bounds in units of p^2, true values:
//   c0 = v0 - xi (v1 + v2 - w0):  inner in [-8, 16]; xi: even lane [-24, 24], odd [-16, 32];  c0 in [-32, 32]  -> + 32: [0, 64]
//   c1 = w1 + xi v2 - v0 - v1:    xi v2: even [-8, 8], odd [0, 16];                           c1 in [-24, 24]  -> + 24: [0, 48]
//   c2 = w2 + v1 - v0 - v2:                                                                   c2 in [-16, 16]  -> + 16: [0, 32]
// all below 2^768 = 96.9 p^2 and below 68 p^2 (results < 8p: two correction steps).
#if ZKP_LAZY_ORDER
// same sums in an order that keeps fewer unreduced values alive across the last product and reduction calls (after c0 only
// d1 = xi v2 - v0 - v1 and d2 = v1 - v0 - v2 instead of v0, v1, v2; arithmetic modulo 2^768, so the reduced values are the
// same): neutral by itself, -0.1 % on top of the finer rendezvous points (profiles/r2z_sync_order_variants.txt) -- shipped
ZKP_NOINLINE void fp6_mul(Fp6 &r, const Fp6 &a, const Fp6 &b) {
    ZKP_CODE_SYNC(4);
    FpW w = fp2_mulw(fp2_add(a.c1, a.c2), fp2_add(b.c1, b.c2));
    FpW v1 = fp2_mulw(a.c1, b.c1);
    FpW v2 = fp2_mulw(a.c2, b.c2);
    FpW t0 = fpw_mul_nr(fpw_sub(fpw_add(v1, v2), w));
    FpW v0 = fp2_mulw(a.c0, b.c0);
    Fp2 c0 = fp2_redc<1>(fpw_add_const(fpw_sub(v0, t0), ZKP_P2X32));
    FpW d1 = fpw_sub(fpw_mul_nr(v2), fpw_add(v0, v1));
    FpW d2 = fpw_sub(v1, fpw_add(v0, v2));
    w = fp2_mulw(fp2_add(a.c0, a.c1), fp2_add(b.c0, b.c1));
    Fp2 c1 = fp2_redc<1>(fpw_add_const(fpw_add(w, d1), ZKP_P2X24));
    w = fp2_mulw(fp2_add(a.c0, a.c2), fp2_add(b.c0, b.c2));
    r.c2 = fp2_redc<1>(fpw_add_const(fpw_add(w, d2), ZKP_P2X16));
    r.c0 = c0;
    r.c1 = c1;
}
#else
ZKP_NOINLINE void fp6_mul(Fp6 &r, const Fp6 &a, const Fp6 &b) {
    ZKP_CODE_SYNC(4);
    FpW v1 = fp2_mulw(a.c1, b.c1);
    FpW v2 = fp2_mulw(a.c2, b.c2);
    FpW w = fp2_mulw(fp2_add(a.c1, a.c2), fp2_add(b.c1, b.c2));
    FpW t0 = fpw_mul_nr(fpw_sub(fpw_add(v1, v2), w));
    FpW v0 = fp2_mulw(a.c0, b.c0);
    Fp2 c0 = fp2_redc<1>(fpw_add_const(fpw_sub(v0, t0), ZKP_P2X32));
    w = fp2_mulw(fp2_add(a.c0, a.c1), fp2_add(b.c0, b.c1));
    FpW t1 = fpw_sub(fpw_add(w, fpw_mul_nr(v2)), fpw_add(v0, v1));
    Fp2 c1 = fp2_redc<1>(fpw_add_const(t1, ZKP_P2X24));
    w = fp2_mulw(fp2_add(a.c0, a.c2), fp2_add(b.c0, b.c2));
    FpW t2 = fpw_sub(fpw_add(w, v1), fpw_add(v0, v2));
    r.c2 = fp2_redc<1>(fpw_add_const(t2, ZKP_P2X16));
    r.c0 = c0;
    r.c1 = c1;
}
#endif
#else
ZKP_NOINLINE void fp6_mul(Fp6 &r, const Fp6 &a, const Fp6 &b) {
    ZKP_CODE_SYNC(4);
    Fp2 v0 = fp2_mul(a.c0, b.c0);
    Fp2 v1 = fp2_mul(a.c1, b.c1);
    Fp2 v2 = fp2_mul(a.c2, b.c2);
    Fp2 t0 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a.c1, a.c2), fp2_add(b.c1, b.c2)), v1), v2);
    Fp2 t1 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a.c0, a.c1), fp2_add(b.c0, b.c1)), v0), v1);
    Fp2 t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a.c0, a.c2), fp2_add(b.c0, b.c2)), v0), v2);
    r.c0 = fp2_add(v0, fp2_mul_nr(t0));
    r.c1 = fp2_add(t1, fp2_mul_nr(v2));
    r.c2 = fp2_add(t2, v1);
}
#endif
// src/fp6.rs:274-288 ; r may alias a
ZKP_NOINLINE void fp6_sqr(Fp6 &r, const Fp6 &a) {
    ZKP_CODE_SYNC(4);
    Fp2 s0 = fp2_sqr(a.c0);
    Fp2 ab = fp2_mul(a.c0, a.c1);
    Fp2 s1 = fp2_dbl(ab);
    Fp2 s2 = fp2_sqr(fp2_add(fp2_sub(a.c0, a.c1), a.c2));
    Fp2 bc = fp2_mul(a.c1, a.c2);
    Fp2 s3 = fp2_dbl(bc);
    Fp2 s4 = fp2_sqr(a.c2);
    r.c0 = fp2_add(fp2_mul_nr(s3), s0);
    r.c1 = fp2_add(fp2_mul_nr(s4), s1);
    r.c2 = fp2_sub(fp2_sub(fp2_add(fp2_add(s1, s2), s3), s0), s4);
}
// a * (0, c1, 0)  -- src/fp6.rs:102-108 ; r may alias a
ZKP_NOINLINE void fp6_mul_by_1(Fp6 &r, const Fp6 &a, const Fp2 &c1) {
    ZKP_CODE_SYNC(4);
    Fp2 t0 = fp2_mul_nr(fp2_mul(a.c2, c1));
    Fp2 t1 = fp2_mul(a.c0, c1);
    Fp2 t2 = fp2_mul(a.c1, c1);
    r.c0 = t0; r.c1 = t1; r.c2 = t2;
}
// a * (c0, c1, 0)  -- src/fp6.rs:110-125 ; r may alias a
#if ZKP_LAZY & 2
// Lazy reduction: 5 unreduced products, 3 reductions (1908 wide MACs per lane instead of 2220).  xi (a2 c1) is formed as
// a2 (xi c1), so no unreduced value crosses the lane pair.  Bounds (units of p^2, true values):
//   r0 = a0 c0 + a2 (xi c1) in [0, 16];  r1 = (c0 + c1)(a0 + a1) - a0 c0 - a1 c1 in [-16, 8] -> + 16: [0, 24];
//   r2 = a2 c0 + a1 c1 in [0, 16]: all below 29 p^2 (results < 4p, one correction step)
ZKP_NOINLINE void fp6_mul_by_01(Fp6 &r, const Fp6 &a, const Fp2 &c0, const Fp2 &c1) {
    ZKP_CODE_SYNC(4);
    FpW aa = fp2_mulw(a.c0, c0);
    FpW t = fp2_mulw(a.c2, fp2_mul_nr(c1));
    Fp2 r0 = fp2_redc<0>(fpw_add(aa, t));
    FpW bb = fp2_mulw(a.c1, c1);
    t = fp2_mulw(fp2_add(c0, c1), fp2_add(a.c0, a.c1));
    Fp2 r1 = fp2_redc<0>(fpw_add_const(fpw_sub(t, fpw_add(aa, bb)), ZKP_P2X16));
    t = fp2_mulw(a.c2, c0);
    r.c2 = fp2_redc<0>(fpw_add(t, bb));
    r.c0 = r0;
    r.c1 = r1;
}
#else
ZKP_NOINLINE void fp6_mul_by_01(Fp6 &r, const Fp6 &a, const Fp2 &c0, const Fp2 &c1) {
    ZKP_CODE_SYNC(4);
    Fp2 a_a = fp2_mul(a.c0, c0);
    Fp2 b_b = fp2_mul(a.c1, c1);
    Fp2 t1 = fp2_add(fp2_mul_nr(fp2_mul(a.c2, c1)), a_a);
    Fp2 t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(c0, c1), fp2_add(a.c0, a.c1)), a_a), b_b);
    Fp2 t3 = fp2_add(fp2_mul(a.c2, c0), b_b);
    r.c0 = t1; r.c1 = t2; r.c2 = t3;
}
#endif
// src/fp6.rs:291-309 ; zero maps to zero ; r may alias a.  Split like fp2_inv: `c` receives the three
// cofactors, the return value is the Fp2 whose inverse scales them.
ZKP_NOINLINE Fp2 fp6_inv_prepare(Fp6 &c, const Fp6 &a) {
    Fp2 c0 = fp2_sub(fp2_sqr(a.c0), fp2_mul_nr(fp2_mul(a.c1, a.c2)));
    Fp2 c1 = fp2_sub(fp2_mul_nr(fp2_sqr(a.c2)), fp2_mul(a.c0, a.c1));
    Fp2 c2 = fp2_sub(fp2_sqr(a.c1), fp2_mul(a.c0, a.c2));
    Fp2 t = fp2_mul_nr(fp2_add(fp2_mul(a.c1, c2), fp2_mul(a.c2, c1)));
    t = fp2_add(t, fp2_mul(a.c0, c0));
    c.c0 = c0; c.c1 = c1; c.c2 = c2;
    return t;
}
ZKP_NOINLINE void fp6_inv_finish(Fp6 &r, const Fp6 &c, const Fp2 &tinv) {
    r.c0 = fp2_mul(tinv, c.c0);
    r.c1 = fp2_mul(tinv, c.c1);
    r.c2 = fp2_mul(tinv, c.c2);
}
ZKP_HD void fp6_inv(Fp6 &r, const Fp6 &a) {
    Fp6 c;
    Fp2 t = fp6_inv_prepare(c, a);
    fp6_inv_finish(r, c, fp2_inv(t));
}

// ------------------------------------------------------------------ Fp12
struct Fp12 {
    Fp6 c0, c1;
};

ZKP_HD void fp12_set_one(Fp12 &r) { fp6_set_zero(r.c0); fp6_set_zero(r.c1); r.c0.c0 = fp2_one(); }
ZKP_HD void fp12_conj(Fp12 &r, const Fp12 &a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }   // src/fp12.rs:123-125
// Karatsuba over Fp6, 3 Fp6 mul  -- src/fp12.rs:193-210 ; r may alias a or b
ZKP_NOINLINE void fp12_mul(Fp12 &r, const Fp12 &a, const Fp12 &b) {
    Fp6 aa, bb, sa, sb;
    fp6_mul(aa, a.c0, b.c0);
    fp6_mul(bb, a.c1, b.c1);
    fp6_add(sa, a.c1, a.c0);
    
    fp6_add(sb, b.c0, b.c1);
    
    fp6_mul(sa, sa, sb);
    fp6_sub(sa, sa, aa);
    fp6_sub(sa, sa, bb);
    r.c1 = sa;
    fp6_mul_nr(bb, bb);
    fp6_add(bb, bb, aa);
    r.c0 = bb;
}
// complex squaring, 2 Fp6 mul  -- src/fp12.rs:173-184 ; r may alias a
ZKP_NOINLINE void fp12_sqr(Fp12 &r, const Fp12 &a) {
    Fp6 ab, s, t;
    fp6_mul(ab, a.c0, a.c1);
    fp6_add(s, a.c0, a.c1);
    
    fp6_mul_nr(t, a.c1);
    fp6_add(t, t, a.c0);
    
    fp6_mul(t, t, s);
    fp6_sub(t, t, ab);
    fp6_add(s, ab, ab);
    r.c1 = s;
    fp6_mul_nr(ab, ab);
    fp6_sub(t, t, ab);
    r.c0 = t;
}
#if ZKP_INPLACE12
// In-place forms of the two Fp12 operations of the Miller loop: ONE Fp6 temporary (handed in by the caller, so that
// it can live in shared memory next to f) instead of three thread-local ones each; same field elements.
// (a0 + a1)(a0 + v a1); r may alias a0 (not a1)
// ZKP_SUMS_SIMPLE = 1 (shipped): the second operand sum goes to a thread-local Fp6 and fp6_mul does the rest -- 26 KB less
// hot code, 21 Fp2 additions and 4 xi-multiplications fewer per Fp12 squaring than forming every sum on the fly (0):
// Miller loop at 2^20 273.3 -> 272.0 ms (profiles/r2l_lazy_reduction.txt)
#ifndef ZKP_SUMS_SIMPLE
#define ZKP_SUMS_SIMPLE 1
#endif
#if (ZKP_LAZY & 1) || ZKP_SUMS_SIMPLE
ZKP_NOINLINE void fp6_mul_sums(Fp6 &r, const Fp6 &a0, const Fp6 &a1) {
    Fp6 b;
    b.c0 = fp2_add(a0.c0, fp2_mul_nr(a1.c2));
    b.c1 = fp2_add(a0.c1, a1.c0);
    b.c2 = fp2_add(a0.c2, a1.c1);
    fp6_add(r, a0, a1);
    fp6_mul(r, r, b);
}
#else
// ... with the operand sums formed on the fly; r may alias a0 or a1
ZKP_NOINLINE void fp6_mul_sums(Fp6 &r, const Fp6 &a0, const Fp6 &a1) {
    ZKP_CODE_SYNC(4);
    Fp2 v0 = fp2_mul(fp2_add(a0.c0, a1.c0), fp2_add(a0.c0, fp2_mul_nr(a1.c2)));
    Fp2 v1 = fp2_mul(fp2_add(a0.c1, a1.c1), fp2_add(a0.c1, a1.c0));
    Fp2 v2 = fp2_mul(fp2_add(a0.c2, a1.c2), fp2_add(a0.c2, a1.c1));
    Fp2 t0, t1, t2;
    {
        Fp2 s1 = fp2_add(a0.c1, a1.c1), s2 = fp2_add(a0.c2, a1.c2);
        Fp2 u1 = fp2_add(a0.c1, a1.c0), u2 = fp2_add(a0.c2, a1.c1);
        t0 = fp2_sub(fp2_sub(fp2_mul(fp2_add(s1, s2), fp2_add(u1, u2)), v1), v2);
    }
    {
        Fp2 s0 = fp2_add(a0.c0, a1.c0), s1 = fp2_add(a0.c1, a1.c1);
        Fp2 u0 = fp2_add(a0.c0, fp2_mul_nr(a1.c2)), u1 = fp2_add(a0.c1, a1.c0);
        t1 = fp2_sub(fp2_sub(fp2_mul(fp2_add(s0, s1), fp2_add(u0, u1)), v0), v1);
    }
    {
        Fp2 s0 = fp2_add(a0.c0, a1.c0), s2 = fp2_add(a0.c2, a1.c2);
        Fp2 u0 = fp2_add(a0.c0, fp2_mul_nr(a1.c2)), u2 = fp2_add(a0.c2, a1.c1);
        t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(s0, s2), fp2_add(u0, u2)), v0), v2);
    }
    r.c0 = fp2_add(v0, fp2_mul_nr(t0));
    r.c1 = fp2_add(t1, fp2_mul_nr(v2));
    r.c2 = fp2_add(t2, v1);
}
#endif
// f <- f^2 (complex squaring, src/fp12.rs:173-184): T = c0 c1; c0 <- (c0 + c1)(c0 + v c1) - T - v T; c1 <- 2 T
ZKP_NOINLINE void fp12_sqr_inplace(Fp12 &f, Fp6 &T) {
    fp6_mul(T, f.c0, f.c1);
    fp6_mul_sums(f.c0, f.c0, f.c1);
    f.c1.c0 = fp2_dbl(T.c0);
    f.c1.c1 = fp2_dbl(T.c1);
    f.c1.c2 = fp2_dbl(T.c2);
    f.c0.c0 = fp2_sub(fp2_sub(f.c0.c0, T.c0), fp2_mul_nr(T.c2));
    f.c0.c1 = fp2_sub(fp2_sub(f.c0.c1, T.c1), T.c0);
    f.c0.c2 = fp2_sub(fp2_sub(f.c0.c2, T.c2), T.c1);
}
// f <- f * (c0 + c1 v + c4 v w) (src/fp12.rs:99-111): T = (f.c0 + f.c1)(c0, c1 + c4); f.c0 <- aa; f.c1 <- bb; combine
ZKP_NOINLINE void fp12_mul_by_014_inplace(Fp12 &f, Fp6 &T, const Fp2 &c0, const Fp2 &c1, const Fp2 &c4) {
    fp6_add(T, f.c1, f.c0);
    fp6_mul_by_01(T, T, c0, fp2_add(c1, c4));
    fp6_mul_by_01(f.c0, f.c0, c0, c1);          // aa
    fp6_mul_by_1(f.c1, f.c1, c4);               // bb
    Fp2 a, b;
    a = f.c0.c0; b = f.c1.c0;
    f.c0.c0 = fp2_add(a, fp2_mul_nr(f.c1.c2));
    Fp2 n10 = fp2_sub(fp2_sub(T.c0, a), b);
    a = f.c0.c1;
    f.c0.c1 = fp2_add(a, b);
    Fp2 b1 = f.c1.c1;
    Fp2 n11 = fp2_sub(fp2_sub(T.c1, a), b1);
    a = f.c0.c2;
    f.c0.c2 = fp2_add(a, b1);
    f.c1.c2 = fp2_sub(fp2_sub(T.c2, a), f.c1.c2);
    f.c1.c0 = n10;
    f.c1.c1 = n11;
}
#endif
// f * (c0 + c1 v + c4 v w): sparse line multiplication  -- src/fp12.rs:99-111 ; in place
ZKP_NOINLINE void fp12_mul_by_014(Fp12 &f, const Fp2 &c0, const Fp2 &c1, const Fp2 &c4) {
    Fp6 aa, bb, t;
    fp6_mul_by_01(aa, f.c0, c0, c1);
    fp6_mul_by_1(bb, f.c1, c4);
    Fp2 o = fp2_add(c1, c4);
    fp6_add(t, f.c1, f.c0);
    
    fp6_mul_by_01(t, t, c0, o);
    fp6_sub(t, t, aa);
    fp6_sub(t, t, bb);
    f.c1 = t;
    fp6_mul_nr(bb, bb);
    fp6_add(bb, bb, aa);
    f.c0 = bb;
}
// src/fp12.rs:186-190 ; zero maps to zero ; r may alias a.  Split around the one Fp inversion:
//   prepare: c = cofactors of (a.c0^2 - v a.c1^2)^-1, t = the Fp2 to invert, returns n = norm(t)
//   finish : given ninv = 1/n, r = a^-1
ZKP_NOINLINE Fp fp12_inv_prepare(Fp6 &c, Fp2 &t, const Fp12 &a) {
    Fp6 t0, t1;
    fp6_sqr(t0, a.c0);
    fp6_sqr(t1, a.c1);
    fp6_mul_nr(t1, t1);
    fp6_sub(t0, t0, t1);
    t = fp6_inv_prepare(c, t0);
    return fp2_norm(t);
}
ZKP_NOINLINE void fp12_inv_finish(Fp12 &r, const Fp12 &a, const Fp6 &c, const Fp2 &t, const Fp &ninv) {
    Fp6 t0;
    fp6_inv_finish(t0, c, fp2_inv_finish(t, ninv));
    fp6_mul(r.c0, a.c0, t0);
    fp6_neg(t0, t0);
    fp6_mul(r.c1, a.c1, t0);
}
ZKP_HD void fp12_inv(Fp12 &r, const Fp12 &a) {
    Fp6 c;
    Fp2 t;
    Fp n = fp12_inv_prepare(c, t, a);
    fp12_inv_finish(r, a, c, t, fp_inv(n));
}
// f^(p^k), k = 1..3, with the TRUE coefficients gamma_{k,i} = xi^(i (p^k-1)/6) on the w-power
// basis (c0.c0, c1.c0, c0.c1, c1.c1, c0.c2, c1.c2 <-> w^0..w^5).  The reference's
// Fp12::frobenius_map (src/fp12.rs:143-170) has this shape but inherits wrong Fp6 constants
// (src/fp6.rs:147-173); this is the mathematically correct map.  r may alias a.
ZKP_NOINLINE void fp12_frobenius(Fp12 &r, const Fp12 &a, int k) {
    ZKP_CODE_SYNC(4);
    const Fp2 *src[6] = {&a.c0.c0, &a.c1.c0, &a.c0.c1, &a.c1.c1, &a.c0.c2, &a.c1.c2};
    Fp2 *dst[6] = {&r.c0.c0, &r.c1.c0, &r.c0.c1, &r.c1.c1, &r.c0.c2, &r.c1.c2};
    const uint32_t *tab = ZKP_FROB + (k - 1) * (10 * ZKP_NL);
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        Fp2 c = *src[i];
        if (k & 1) c = fp2_conj(c);
        if (i > 0) {
            c = fp2_mul(c, fp2_const(tab + (i - 1) * 2 * ZKP_NL));
        }
        *dst[i] = c;
    }
}

// Granger-Scott squaring in the cyclotomic subgroup: 9 Fp2 squarings.  r may alias f.
#if ZKP_LAZY & 4
// Lazy reduction: 3 unreduced squares (<= 8 p^2 per lane), 2 reductions (744 wide MACs per lane instead of 900).
//   c0 = a^2 + xi b^2:  xi b^2 even lane [-8, 8], odd [0, 16];  c0 in [-8, 24] -> + 8: [0, 32]  (< 68 p^2: two corrections)
//   c1 = (a + b)^2 - a^2 - b^2 in [-16, 8] -> + 16: [0, 24]                                     (< 29 p^2: one correction)
// c0, c1 must not alias a, b.
ZKP_HD void fp4_square(Fp2 &c0, Fp2 &c1, const Fp2 &a, const Fp2 &b) {
    FpW pa = fp2_sqrw(a), pb = fp2_sqrw(b);
    c0 = fp2_redc<1>(fpw_add_const(fpw_add(pa, fpw_mul_nr(pb)), ZKP_P2X8));
    FpW ps = fp2_sqrw(fp2_add(a, b));
    c1 = fp2_redc<0>(fpw_add_const(fpw_sub(ps, fpw_add(pa, pb)), ZKP_P2X16));
}
#else
ZKP_HD void fp4_square(Fp2 &c0, Fp2 &c1, const Fp2 &a, const Fp2 &b) {
    Fp2 t0 = fp2_sqr(a);
    Fp2 t1 = fp2_sqr(b);
    c0 = fp2_add(fp2_mul_nr(t1), t0);
    c1 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(a, b)), t0), t1);
}
#endif
// 3t - 2z = 2(t - z) + t  and  3t + 2z = 2(t + z) + t
ZKP_HD Fp2 cyc_minus(const Fp2 &t, const Fp2 &z) {
    Fp2 w = fp2_sub(t, z);
    return fp2_add(fp2_dbl(w), t);
}
ZKP_HD Fp2 cyc_plus(const Fp2 &t, const Fp2 &z) {
    Fp2 w = fp2_add(t, z);
    return fp2_add(fp2_dbl(w), t);
}
ZKP_NOINLINE void fp12_cyclotomic_sqr(Fp12 &r, const Fp12 &f) {
    ZKP_CODE_SYNC(4);
    Fp2 t0, t1;
    // (z0, z1) = (c0.c0, c1.c1)
    {
        Fp2 z0 = f.c0.c0, z1 = f.c1.c1;
        fp4_square(t0, t1, z0, z1);
        r.c0.c0 = cyc_minus(t0, z0);
        r.c1.c1 = cyc_plus(t1, z1);
    }
    // (z2, z3) = (c1.c0, c0.c2) feeds z4', z5' ; (z4, z5) = (c0.c1, c1.c2) feeds z2', z3'
    {
        Fp2 z2 = f.c1.c0, z3 = f.c0.c2, z4 = f.c0.c1, z5 = f.c1.c2;
        Fp2 t2, t3;
        fp4_square(t0, t1, z2, z3);
        fp4_square(t2, t3, z4, z5);
        r.c0.c1 = cyc_minus(t0, z4);
        r.c1.c2 = cyc_plus(t1, z5);
        t0 = fp2_mul_nr(t3);
        r.c1.c0 = cyc_plus(t0, z2);
        r.c0.c2 = cyc_minus(t2, z3);
    }
}

}  // namespace zkp
