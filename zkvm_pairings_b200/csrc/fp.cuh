// Fp: BLS12-381 base field on sm_100a -- 14 signed limbs of 28 bits, Montgomery form R = 2^392.
//
// Replaces the reference's host Fp arithmetic (BigUint mul/add followed by "% p",
// /root/reference/src/fp.rs:351-368, :415-434) and its zkVM precompile calls
// (bls12381_sys_bigint / syscall_bls12381_fp_mulmod, src/fp.rs:126,376,443).  Values cross the
// boundary as canonical little-endian limbs (src/fp.rs:24) and are converted at load/store.
//
// Why 14 x 28 bits and not 12 x 32 (the v0 design): on B200 a plain IMAD.WIDE issues at the full
// fmaheavy rate (measured 18.2 T/s) but the carry-chained IMAD.WIDE.U32.X that saturated 32-bit
// limbs need issues at HALF of it (9.1 T/s; profiles/r1a_v0_saturated_cios_ncu_summary.txt).
// With 28-bit limbs every 32x32->64 product has 8 spare bits, so whole columns of partial products
// accumulate in 64-bit registers with NO carries: a Montgomery product is 14*14 (a*b) + 14*14
// (m*p) + 14 (m = t*n0') = 420 full-rate IMADs with 14-way instruction-level parallelism, against
// 300 half-rate ones before.  Additions and subtractions become 14 independent 32-bit adds (no
// carry chain, no conditional subtraction): limbs are SIGNED and values are kept lazily reduced.
//
// Representation invariants (checked statically by the bound tracker, see ZKP_TRACK_BOUNDS):
//   * value  v = sum l[i] * 2^(28 i), congruent to (x * 2^392) mod p, |v| < 2^11 * p;
//   * "normalized": l[0..12] in [0, 2^28), l[13] small and signed (it carries the sign of v);
//   * add/sub/neg are limb-wise and only grow the limb bound; fp_wnorm() brings l[0..12] back to
//     [-16, 2^28 + 16] in one parallel round;
//   * fp_mul needs  14 * max|a.l| * max|b.l| + 2^60 < 2^63  and returns a normalized value in
//     (ab/R, ab/R + p), i.e. within (-0.1p, 1.1p) for operands below 16p.
//
// The header is plain C++ (no PTX), so the identical code is exercised on the CPU by the dev
// simulation tests/host_sim/sim.cpp (never linked into libzkpair.so).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(ZKP_HOST_SIM)
#define ZKP_DEVICE_BUILD 1
#define ZKP_HD __device__ __forceinline__
#define ZKP_MEMBER __device__ __forceinline__
#define ZKP_HOSTDEV __host__ __device__ inline
#define ZKP_NOINLINE __device__ __noinline__
#define ZKP_CONST __device__ __constant__ const
#else
#define ZKP_HD static inline
#define ZKP_MEMBER inline
#define ZKP_HOSTDEV static inline
#define ZKP_NOINLINE static __attribute__((noinline))
#define ZKP_CONST static const
#endif

#include "consts.cuh"

#ifdef ZKP_TRACK_BOUNDS
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#endif

namespace zkp {

#define ZKP_NL 14
#define ZKP_M28 0x0fffffff

struct Fp {
    int32_t l[ZKP_NL];
#ifdef ZKP_TRACK_BOUNDS
    // worst-case bounds carried along in the CPU dev simulation only: |l[0..12]| <= lb,
    // |l[13]| <= tb, |value| <= vb * p.  They depend on the formula DAG, not on the data.
    double lb, tb, vb;
#endif
};

#ifdef ZKP_TRACK_BOUNDS
#define ZKP_TOP_PER_P 106514.0          /* ceil(p / 2^364) */
#define ZKP_R_OVER_P 2520.0             /* floor(2^392 / p) */
static double g_max_lb = 0, g_max_tb = 0, g_max_vb = 0, g_max_col = 0;
static inline void zkp_bound_fail(const char *what, double v) {
    fprintf(stderr, "ZKP bound violation: %s (%.4g = 2^%.2f)\n", what, v, __builtin_log2(v));
    void *bt[32];
    int n = backtrace(bt, 32);
    backtrace_symbols_fd(bt, n, 2);
    abort();
}
static inline void zkp_set_bounds(Fp &r, double lb, double tb, double vb) {
    r.lb = lb; r.tb = tb; r.vb = vb;
    if (lb > g_max_lb) g_max_lb = lb;
    if (tb > g_max_tb) g_max_tb = tb;
    if (vb > g_max_vb) g_max_vb = vb;
    if (lb >= 2147483000.0) zkp_bound_fail("limb magnitude reaches 2^31", lb);
    if (tb >= 2147483000.0) zkp_bound_fail("top limb magnitude reaches 2^31", tb);
}
#define ZKP_SETB(r, lb, tb, vb) zkp_set_bounds(r, lb, tb, vb)
#else
#define ZKP_SETB(r, lb, tb, vb)
#endif

// ------------------------------------------------------------------ lane pairing
//
// Two adjacent lanes (2k, 2k+1) of a warp cooperate on one pairing: every Fp2 value is split, the
// even lane holds c0 and the odd lane c1 (tower.cuh).  The only communication is a 14-word
// shfl.xor with the partner, synchronised on the pair's own two-lane mask so that pairs may
// diverge from each other (point generation, infinity handling) without deadlock.
#ifdef ZKP_DEVICE_BUILD
ZKP_HD int lane_par() { return (int)(threadIdx.x & 1u); }
ZKP_HD unsigned pair_mask() { return 3u << (threadIdx.x & 30u); }
ZKP_HD int32_t word_xchg(int32_t v) { return __shfl_xor_sync(pair_mask(), v, 1); }
#else
// CPU dev simulation: the two lanes are two host threads in lock-step (tests/host_sim/sim.cpp)
extern thread_local int zkp_sim_par;
int32_t zkp_sim_word_xchg(int32_t v);
void zkp_sim_xchg(void *buf, unsigned long bytes);
ZKP_HD int lane_par() { return zkp_sim_par; }
ZKP_HD int32_t word_xchg(int32_t v) { return zkp_sim_word_xchg(v); }
#endif
ZKP_HD bool lane_or(bool x) { return (x | (word_xchg(x ? 1 : 0) != 0)); }
ZKP_HD bool lane_and(bool x) { return (x & (word_xchg(x ? 1 : 0) != 0)); }

// ------------------------------------------------------------------ constants / trivial ops
ZKP_HD Fp fp_const(const int32_t *k) {   // a normalized constant (Montgomery form), value < p
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = k[i];
    ZKP_SETB(r, 268435456.0, ZKP_TOP_PER_P, 1.0);
    return r;
}
ZKP_HD Fp fp_zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = 0;
    ZKP_SETB(r, 0.0, 0.0, 0.0);
    return r;
}
ZKP_HD Fp fp_one() { return fp_const(ZKP_ONE); }   // Montgomery one; canonical one is [1,0,..] (src/fp.rs:154-156)

// (a + b), lazily: no carry, no reduction   -- value-equal mod p to src/fp.rs:351-368
ZKP_HD Fp fp_add(const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = a.l[i] + b.l[i];
    ZKP_SETB(r, a.lb + b.lb, a.tb + b.tb, a.vb + b.vb);
    return r;
}
// (a - b), lazily (limbs are signed)        -- src/fp.rs:407-411
ZKP_HD Fp fp_sub(const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = a.l[i] - b.l[i];
    ZKP_SETB(r, a.lb + b.lb, a.tb + b.tb, a.vb + b.vb);
    return r;
}
// -a                                         -- src/fp.rs:381-405
ZKP_HD Fp fp_neg(const Fp &a) {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = -a.l[i];
    ZKP_SETB(r, a.lb, a.tb, a.vb);
    return r;
}
ZKP_HD Fp fp_dbl(const Fp &a) { return fp_add(a, a); }

// the partner lane's copy of a value
ZKP_HD Fp fp_xchg(const Fp &a) {
#ifdef ZKP_DEVICE_BUILD
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = word_xchg(a.l[i]);
    return r;
#else
    Fp r = a;                      // bounds travel with the value in the tracker build
    zkp_sim_xchg(&r, sizeof(Fp));
    return r;
#endif
}
// c ? a : b, limb-wise (c is lane-uniform per value, not per limb)
ZKP_HD Fp fp_select(bool c, const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = c ? a.l[i] : b.l[i];
#ifdef ZKP_TRACK_BOUNDS
    ZKP_SETB(r, a.lb > b.lb ? a.lb : b.lb, a.tb > b.tb ? a.tb : b.tb, a.vb > b.vb ? a.vb : b.vb);
#endif
    return r;
}

// Weak normalization: one parallel carry round.  l[0..12] end in [-16, 2^28 + 16]; value unchanged.
ZKP_HD Fp fp_wnorm(const Fp &a) {
    Fp r;
    r.l[0] = a.l[0] & ZKP_M28;
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) r.l[i] = (a.l[i] & ZKP_M28) + (a.l[i - 1] >> 28);
    r.l[ZKP_NL - 1] = a.l[ZKP_NL - 1] + (a.l[ZKP_NL - 2] >> 28);
    ZKP_SETB(r, 268435456.0 + a.lb / 268435456.0 + 1.0, a.tb + a.lb / 268435456.0 + 1.0, a.vb);
    return r;
}
// Full normalization: serial carry propagation.  l[0..12] end in [0, 2^28); l[13] takes the rest.
ZKP_HD Fp fp_norm(const Fp &a) {
    Fp r;
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL - 1; i++) {
        int32_t t = a.l[i] + c;
        r.l[i] = t & ZKP_M28;
        c = t >> 28;
    }
    r.l[ZKP_NL - 1] = a.l[ZKP_NL - 1] + c;
    ZKP_SETB(r, 268435455.0, a.tb + a.lb / 268435456.0 + 1.0, a.vb);
    return r;
}

// Value reduction for the rare data paths that carry a value through additions only (the z terms
// of the cyclotomic squaring): subtracts q*p with q = round(v/p) estimated from the top limb, so the
// result lies in (-0.51p, 0.51p).  Any limb bound below 2^31 and |v| < 1000p are accepted: the
// multiply-subtract runs in 64 bits (14 IMAD.WIDE) and one carry round brings the limbs back to
// [-1026, 2^28 + 1026].
#define ZKP_VREDUCE_K 10322781ll   /* round(2^404 / p) = 2^40 / (p / 2^364) */
ZKP_HD Fp fp_vreduce(const Fp &a) {
#ifdef ZKP_TRACK_BOUNDS
    if (a.vb > 1000.0) zkp_bound_fail("fp_vreduce input value bound", a.vb);
#endif
    // the top limb only sees v / 2^364 after the lower limbs' carries are folded in
    int32_t top = a.l[ZKP_NL - 1] + (a.l[ZKP_NL - 2] >> 28);
    int32_t q = (int32_t)(((int64_t)top * ZKP_VREDUCE_K + (1ll << 39)) >> 40);
    Fp r;
    int64_t x = (int64_t)a.l[0] - (int64_t)q * ZKP_P[0];
    r.l[0] = (int32_t)((uint32_t)x & ZKP_M28);
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) {
        int32_t c = (int32_t)(x >> 28);
        x = (int64_t)a.l[i] - (int64_t)q * ZKP_P[i];
        r.l[i] = (int32_t)((uint32_t)x & ZKP_M28) + c;
    }
    r.l[ZKP_NL - 1] = a.l[ZKP_NL - 1] - q * ZKP_P[ZKP_NL - 1] + (int32_t)(x >> 28);
    ZKP_SETB(r, 268435456.0 + 1026.0, 0.51 * ZKP_TOP_PER_P + 1030.0, 0.52);
    return r;
}

// ------------------------------------------------------------------ Montgomery product
//
// a*b/2^392 mod p as a normalized value in (ab/R, ab/R + p).  Separated operand scanning over 27
// 64-bit column accumulators: 196 IMAD.WIDE for a*b, then per reduction step one IMAD (m) and 14
// IMAD.WIDE.U32 (m*p); no carries anywhere, the columns are resolved by two shifts each.
// m = (t * n0') mod 2^28 as a plain 32-bit value.  The opaque move keeps the compiler from
// re-deriving the multiplier as a masked 64-bit quantity (it then emits 64-bit multiplies whose
// zero high halves survive as an extra add after every IMAD.WIDE).
ZKP_HD int32_t mont_m(uint32_t t_lo) {
    int32_t m = (int32_t)((t_lo * ZKP_N0INV) & ZKP_M28);
#ifdef ZKP_DEVICE_BUILD
    asm("" : "+r"(m));
#endif
    return m;
}
// reduction half shared by the one- and two-product forms: col[0..26] -> normalized limbs
ZKP_HD Fp mont_reduce(int64_t *col) {
    int64_t carry = 0;
#pragma unroll
    for (int k = 0; k < ZKP_NL; k++) {
        int64_t t = col[k] + carry;
        int32_t m = mont_m((uint32_t)t);
        t += (int64_t)m * (int64_t)ZKP_P[0];
        carry = t >> 28;   // exact: the low 28 bits of t are zero now
#pragma unroll
        for (int j = 1; j < ZKP_NL; j++) col[k + j] += (int64_t)m * (int64_t)ZKP_P[j];
    }
    Fp r;
#pragma unroll
    for (int k = ZKP_NL; k < 2 * ZKP_NL - 1; k++) {
        int64_t t = col[k] + carry;
        r.l[k - ZKP_NL] = (int32_t)((uint32_t)t & ZKP_M28);
        carry = t >> 28;
    }
    r.l[ZKP_NL - 1] = (int32_t)carry;
    return r;
}
ZKP_HD Fp mont_mul(const Fp &a, const Fp &b) {
#ifdef ZKP_TRACK_BOUNDS
    {
        double A = a.lb > a.tb ? a.lb : a.tb, B = b.lb > b.tb ? b.lb : b.tb;
        double col = 14.0 * A * B + 14.0 * 72057594037927936.0 + 1099511627776.0;
        if (col > g_max_col) g_max_col = col;
        if (col >= 9.2e18) zkp_bound_fail("column accumulator reaches 2^63 in mont_mul", col);
        if (a.vb * b.vb / ZKP_R_OVER_P + 1.0 > 2000.0) zkp_bound_fail("product value bound", a.vb * b.vb);
    }
#endif
    int64_t col[2 * ZKP_NL - 1];
#pragma unroll
    for (int j = 0; j < ZKP_NL; j++) col[j] = (int64_t)a.l[0] * (int64_t)b.l[j];
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) {
#pragma unroll
        for (int j = 0; j < ZKP_NL - 1; j++) col[i + j] += (int64_t)a.l[i] * (int64_t)b.l[j];
        col[i + ZKP_NL - 1] = (int64_t)a.l[i] * (int64_t)b.l[ZKP_NL - 1];
    }
    Fp r = mont_reduce(col);
#ifdef ZKP_TRACK_BOUNDS
    {
        double vb = a.vb * b.vb / ZKP_R_OVER_P + 1.0;
        ZKP_SETB(r, 268435455.0, vb * ZKP_TOP_PER_P + 2.0, vb);
    }
#endif
    return r;
}
// (u*v + w*z)/2^392 mod p with ONE reduction (lazy "sum of products"): 392 + 210 IMADs.  This is
// one lane's half of an Fp2 product.  Needs 14*(|u||v| + |w||z|) + 2^60 < 2^63.
ZKP_HD Fp mont_mul2(const Fp &u, const Fp &v, const Fp &w, const Fp &z) {
#ifdef ZKP_TRACK_BOUNDS
    {
        double U = u.lb > u.tb ? u.lb : u.tb, V = v.lb > v.tb ? v.lb : v.tb;
        double W = w.lb > w.tb ? w.lb : w.tb, Z = z.lb > z.tb ? z.lb : z.tb;
        double col = 14.0 * (U * V + W * Z) + 14.0 * 72057594037927936.0 + 1099511627776.0;
        if (col > g_max_col) g_max_col = col;
        if (col >= 9.2e18) zkp_bound_fail("column accumulator reaches 2^63 in mont_mul2", col);
        if ((u.vb * v.vb + w.vb * z.vb) / ZKP_R_OVER_P + 1.0 > 2000.0) zkp_bound_fail("product value bound", u.vb * v.vb + w.vb * z.vb);
    }
#endif
    int64_t col[2 * ZKP_NL - 1];
#pragma unroll
    for (int j = 0; j < ZKP_NL; j++) col[j] = (int64_t)u.l[0] * (int64_t)v.l[j];
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) {
#pragma unroll
        for (int j = 0; j < ZKP_NL - 1; j++) col[i + j] += (int64_t)u.l[i] * (int64_t)v.l[j];
        col[i + ZKP_NL - 1] = (int64_t)u.l[i] * (int64_t)v.l[ZKP_NL - 1];
    }
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
#pragma unroll
        for (int j = 0; j < ZKP_NL; j++) col[i + j] += (int64_t)w.l[i] * (int64_t)z.l[j];
    }
    Fp r = mont_reduce(col);
#ifdef ZKP_TRACK_BOUNDS
    {
        double vb = (u.vb * v.vb + w.vb * z.vb) / ZKP_R_OVER_P + 1.0;
        ZKP_SETB(r, 268435455.0, vb * ZKP_TOP_PER_P + 2.0, vb);
    }
#endif
    return r;
}
// Value-equivalent (after conversion) to src/fp.rs:413-434 / :452-455.
ZKP_HD Fp fp_mul(const Fp &a, const Fp &b) { return mont_mul(a, b); }

// ------------------------------------------------------------------ boundary conversions
//
// Canonical form = 12 saturated 32-bit words (= the six u64 limbs of src/fp.rs:24), value in [0,p).

// true when w < p
ZKP_HD bool words_lt_p(const uint32_t *w) {
    int64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        int64_t d = (int64_t)w[i] - (int64_t)ZKP_P32[i] + borrow;
        borrow = d >> 32;   // 0 or -1
    }
    return borrow != 0;
}
// canonical words -> Montgomery Fp; sets bad when w >= p (such inputs are rejected at the boundary
// because the reference's neg is undefined there, src/fp.rs:383-405)
ZKP_HD Fp fp_from_words(const uint32_t *w, bool &bad) {
    bad = bad | !words_lt_p(w);
    Fp a;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        const int bit = 28 * i, idx = bit >> 5, sh = bit & 31;
        uint32_t lo = w[idx] >> sh;
        uint32_t hi = (sh > 4 && idx + 1 < 12) ? (w[idx + 1] << (32 - sh)) : 0u;
        a.l[i] = (int32_t)((lo | hi) & ZKP_M28);
    }
    ZKP_SETB(a, 268435455.0, 16777216.0, 10.0);   // any 384-bit input (even a rejected one) stays in range
    return mont_mul(a, fp_const(ZKP_R2));
}
// Montgomery Fp (any lazily reduced value the tracker allows) -> canonical words in [0,p)
ZKP_HD void fp_to_words(uint32_t *w, const Fp &m) {
    Fp one = fp_zero();
    one.l[0] = 1;
    ZKP_SETB(one, 1.0, 0.0, 1.0);
    Fp t = mont_mul(m, one);                 // value in (-0.8p, 1.8p) for |m| < 2000p
    t = fp_norm(fp_add(t, fp_const(ZKP_P))); // + p: strictly positive, fully normalized
    uint64_t acc = 0;
    int have = 0, wi = 0;
    uint32_t v[13];
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        acc |= (uint64_t)(uint32_t)t.l[i] << have;
        have += 28;
        if (have >= 32) {
            v[wi++] = (uint32_t)acc;
            acc >>= 32;
            have -= 32;
        }
    }
    v[wi] = (uint32_t)acc;   // wi == 12 here; bits 384.. (zero: value < 3p < 2^384)
    // value < 2.8p: two conditional subtractions of p
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
        uint32_t d[12];
        int64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            int64_t x = (int64_t)v[i] - (int64_t)ZKP_P32[i] + borrow;
            d[i] = (uint32_t)x;
            borrow = x >> 32;
        }
        bool ge = borrow == 0;
#pragma unroll
        for (int i = 0; i < 12; i++) v[i] = ge ? d[i] : v[i];
    }
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = v[i];
}
// Comparisons go through the canonical form (rare: flags and degenerate cases of the group law).
ZKP_HD bool fp_is_zero(const Fp &a) {
    uint32_t w[12];
    fp_to_words(w, a);
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= w[i];
    return t == 0;
}

}  // namespace zkp
