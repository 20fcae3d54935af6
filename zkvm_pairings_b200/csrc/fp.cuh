// Fp: BLS12-381 base field on sm_100a -- 12 x 32-bit limbs, Montgomery form (R = 2^384).
//
// Replaces the reference's host Fp arithmetic (BigUint mul/add followed by "% p",
// /root/reference/src/fp.rs:351-368, :415-434) and its zkVM precompile calls
// (bls12381_sys_bigint / syscall_bls12381_fp_mulmod, src/fp.rs:126,376,443).  Values cross the
// boundary as canonical little-endian limbs (src/fp.rs:24); inside the kernels they are in
// Montgomery form.
//
// The multiplier is a CIOS Montgomery product written as carry chains of 32x32->64 wide
// multiply-accumulates (PTX mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into one
// IMAD.WIDE.U32 with carry-in/out predicates).  Partial products are kept in two interleaved
// accumulators ("even" = 64-bit aligned columns, "odd" = columns offset by 32 bits) so every
// wide MAC lands on a register pair and the two chains of a row are independent (ILP 2).
// 12 rows x (12 a*b + 12 m*p) + 12 (m = t0 * n0') = 300 wide MACs per product.
//
// The same header compiles as plain C++ (ZKP_HOST_SIM) with the carry flag emulated, so the
// limb-level algorithm is exercised by the CPU test-suite.  That build is a DEV SIMULATION of the
// device code for tests only; it is never linked into libzkpair.so.
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

namespace zkp {

struct Fp {
    uint32_t l[12];
};

// ------------------------------------------------------------------ carry-chain primitives
#ifdef ZKP_DEVICE_BUILD
ZKP_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation of the PTX condition-code register (one flag per thread)
static thread_local uint32_t ZKP_CF = 0;
ZKP_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; ZKP_CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKP_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + ZKP_CF; ZKP_CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKP_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + ZKP_CF; }
ZKP_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; ZKP_CF = (uint32_t)(d >> 63); return (uint32_t)d; }
ZKP_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - ZKP_CF; ZKP_CF = (uint32_t)(d >> 63); return (uint32_t)d; }
ZKP_HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - ZKP_CF; }
ZKP_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)(a * b) + c; ZKP_CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKP_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)(a * b) + c + ZKP_CF; ZKP_CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKP_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (((uint64_t)a * b) >> 32) + c + ZKP_CF; ZKP_CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKP_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c + ZKP_CF; }
#endif

// ------------------------------------------------------------------ basic ops

ZKP_HD Fp fp_zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
    return r;
}
ZKP_HD Fp fp_one() {   // Montgomery form of 1; canonical one is [1,0,..] (src/fp.rs:154-156)
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = ZKP_ONE[i];
    return r;
}
ZKP_HD bool fp_is_zero(const Fp &a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i];
    return t == 0;
}
ZKP_HD bool fp_eq(const Fp &a, const Fp &b) {   // raw-limb equality, like src/fp.rs:53-58
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// r = a - k, returns borrow mask (0xffffffff when a < k); k = 12 constant limbs
ZKP_HD uint32_t sub_limbs(Fp &r, const Fp &a, const uint32_t *k) {
    r.l[0] = sub_cc(a.l[0], k[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) r.l[i] = subc_cc(a.l[i], k[i]);
    return subc(0, 0);
}
// a in [0, 2p) -> [0, p)
ZKP_HD Fp fp_reduce_once(const Fp &a) {
    Fp t;
    uint32_t borrow = sub_limbs(t, a, ZKP_P);
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
    return r;
}
// a + b without reduction (caller guarantees the sum stays below 2^384)
ZKP_HD Fp fp_add_nr(const Fp &a, const Fp &b) {
    Fp r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[11] = addc(a.l[11], b.l[11]);
    return r;
}
// (a + b) mod p for a, b in [0,p)   -- src/fp.rs:351-368
ZKP_HD Fp fp_add(const Fp &a, const Fp &b) { return fp_reduce_once(fp_add_nr(a, b)); }
// (a - b) mod p for a, b in [0,p)   -- src/fp.rs:407-411
ZKP_HD Fp fp_sub(const Fp &a, const Fp &b) {
    Fp d;
    d.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) d.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t mask = subc(0, 0);
    Fp r;
    r.l[0] = add_cc(d.l[0], ZKP_P[0] & mask);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = addc_cc(d.l[i], ZKP_P[i] & mask);
    r.l[11] = addc(d.l[11], ZKP_P[11] & mask);
    return r;
}
// a - b + p, for a in [0,2p), b in [0,p]: result in (0, 3p) -- no conditional; feeds a multiplier
ZKP_HD Fp fp_sub_nr(const Fp &a, const Fp &b) {
    Fp t;
    t.l[0] = add_cc(a.l[0], ZKP_P[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) t.l[i] = addc_cc(a.l[i], ZKP_P[i]);
    t.l[11] = addc(a.l[11], ZKP_P[11]);
    Fp r;
    r.l[0] = sub_cc(t.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = subc_cc(t.l[i], b.l[i]);
    r.l[11] = subc(t.l[11], b.l[11]);
    return r;
}
// -a mod p  -- src/fp.rs:381-405 (zero stays zero)
ZKP_HD Fp fp_neg(const Fp &a) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) nz |= a.l[i];
    uint32_t mask = nz ? 0xffffffffu : 0u;
    Fp r;
    r.l[0] = sub_cc(ZKP_P[0] & mask, a.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = subc_cc(ZKP_P[i] & mask, a.l[i]);
    r.l[11] = subc(ZKP_P[11] & mask, a.l[11]);
    return r;
}
ZKP_HD Fp fp_dbl(const Fp &a) { return fp_add(a, a); }

// ------------------------------------------------------------------ Montgomery product
//
// Returns a*b/R mod p as a value in [0, 2p) whenever a*b < p*R; `a` (the operand multiplied
// through every row) must be < 6p, `b` may be any 384-bit value.  See the file header for the
// even/odd accumulator layout.

// x[j..j+1] += k[j]*m for j = 0,2,..,10 (one carry chain, 6 wide MACs); leaves carry-out in CF
template <int OFF>
ZKP_HD void chain_mad(uint32_t *x, const uint32_t *k, uint32_t m) {
    x[0] = mad_lo_cc(k[OFF], m, x[0]);
    x[1] = madc_hi_cc(k[OFF], m, x[1]);
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        x[j] = madc_lo_cc(k[j + OFF], m, x[j]);
        x[j + 1] = madc_hi_cc(k[j + OFF], m, x[j + 1]);
    }
}
// y >>= 64 bits; y[j..j+1] += a[j+1]*m for j = 0,2,..,10, consuming the incoming carry
ZKP_HD void chain_mad_rshift(uint32_t *y, const uint32_t *a, uint32_t m) {
#pragma unroll
    for (int j = 0; j < 10; j += 2) {
        y[j] = madc_lo_cc(a[j + 1], m, y[j + 2]);
        y[j + 1] = madc_hi_cc(a[j + 1], m, y[j + 3]);
    }
    y[10] = madc_lo_cc(a[11], m, 0);
    y[11] = madc_hi(a[11], m, 0);
}
// reduction half of a row: m = x0 * n0'; y += p_odd*m; x += p_even*m; carry of x into y[11]
ZKP_HD void row_reduce(uint32_t *x, uint32_t *y) {
    uint32_t m = x[0] * ZKP_N0INV;
    chain_mad<1>(y, ZKP_P, m);
    chain_mad<0>(x, ZKP_P, m);
    y[11] = addc(y[11], 0);
}
// one full CIOS row (not the first): x = array in the even role, y = array in the odd role
ZKP_HD void row_mul(uint32_t *x, uint32_t *y, const uint32_t *a, uint32_t bi) {
    x[0] = add_cc(x[0], y[1]);
    chain_mad_rshift(y, a, bi);
    chain_mad<0>(x, a, bi);
    y[11] = addc(y[11], 0);
    row_reduce(x, y);
}

ZKP_HD Fp mont_mul_raw(const Fp &a, const Fp &b) {
    uint32_t ev[12], od[12];
    // row 0: disjoint 64-bit products, no carries
#pragma unroll
    for (int j = 0; j < 12; j += 2) {
        uint64_t e = (uint64_t)a.l[j] * b.l[0];
        uint64_t o = (uint64_t)a.l[j + 1] * b.l[0];
        ev[j] = (uint32_t)e;
        ev[j + 1] = (uint32_t)(e >> 32);
        od[j] = (uint32_t)o;
        od[j + 1] = (uint32_t)(o >> 32);
    }
    row_reduce(ev, od);
#pragma unroll
    for (int i = 1; i < 12; i += 2) {
        row_mul(od, ev, a.l, b.l[i]);
        if (i + 1 < 12) row_mul(ev, od, a.l, b.l[i + 1]);
    }
    // after row 11 the even role is `od` (od[0] == 0): T = ev + (od >> 32)
    Fp r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) r.l[k] = addc_cc(ev[k], od[k + 1]);
    r.l[11] = addc(ev[11], 0);
    return r;
}

// Montgomery product, fully reduced to [0,p).  Inputs: a < 6p, a*b < p*R.
// Value-equivalent (after conversion) to src/fp.rs:413-434.
ZKP_HD Fp fp_mul(const Fp &a, const Fp &b) { return fp_reduce_once(mont_mul_raw(a, b)); }
ZKP_HD Fp fp_sqr(const Fp &a) { return fp_mul(a, a); }   // src/fp.rs:452-455

// canonical [0,p) limbs -> Montgomery form, and back
ZKP_HD Fp fp_to_mont(const Fp &a) {
    Fp r2;
#pragma unroll
    for (int i = 0; i < 12; i++) r2.l[i] = ZKP_R2[i];
    return fp_mul(a, r2);
}
ZKP_HD Fp fp_from_mont(const Fp &a) {
    Fp one = fp_zero();
    one.l[0] = 1;
    return fp_mul(a, one);
}
// true when the 12 limbs encode a value < p (boundary check; inputs >= p are rejected because the
// reference's neg is undefined there, src/fp.rs:383-405)
ZKP_HD bool fp_is_canonical(const Fp &a) {
    Fp t;
    return sub_limbs(t, a, ZKP_P) != 0;
}

}  // namespace zkp
