// Fp: BLS12-381 base field on sm_100a -- 12 saturated 32-bit limbs, Montgomery form R = 2^384.
//
// Replaces the reference's host Fp arithmetic (BigUint mul/add followed by "% p",
// /root/reference/src/fp.rs:351-368, :415-434) and its zkVM precompile calls
// (bls12381_sys_bigint / syscall_bls12381_fp_mulmod, src/fp.rs:126,376,443).  Values cross the
// boundary as canonical little-endian limbs (src/fp.rs:24) and are converted at load/store.
//
// Cost model (measured on B200, tools/imad_probe.cu): every 32x32->64 multiply-accumulate
// (IMAD.WIDE.U32, with or without carry in/out) occupies the fmaheavy pipe for 4 cycles per warp
// and is THE scarce resource; IADD3/LOP3/SEL run on the separate ALU pipe.  So the design
// minimises wide MACs and lets additions cost what they cost:
//   * saturated 32-bit limbs: 12 x 12 = 144 MACs per product, 12 x 13 = 156 per Montgomery
//     reduction (the 14 x 28-bit carry-free layout tried in between needs 196 / 210);
//   * rows are carry chains of mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into one
//     IMAD.WIDE.U32(.X) each; partial products live in two interleaved accumulators ("even" =
//     64-bit aligned columns, "odd" = columns offset by 32 bits), so every MAC lands on a
//     register pair and the even/odd chains of a row are independent;
//   * "two products, one reduction": mont_mul2 computes (u*v + w*z)/R with 288 + 156 MACs -- one
//     lane's half of an Fp2 product (tower.cuh) -- instead of 2 x 300.
//
// Representation invariant: every Fp that leaves a function of this file is a value in [0, 2p]
// congruent to x * 2^384 mod p ("2p-redundant"): add/sub correct by +-2p with one conditional
// step and never produce canonical values; only the boundary (fp_to_words) reduces to [0, p).
// Montgomery bound: for T = u*v (+ w*z) < p * 2^384 the result (T + m*p)/R lies in [0, 2p).
// With all operands <= 2p, T <= 8 p^2 < 0.82 * p * 2^384 (p < 2^381: three spare bits).  The
// squaring passes one lazily added operand (<= 4p) together with a corrected one (<= 2p): same
// bound.  The CPU dev simulation (tests/host_sim/sim.cpp) asserts these operand bounds on every
// call (ZKP_SIM_ASSERT); they depend on the call sites, not on the data.
//
// The header is plain C++ plus ten PTX carry primitives that have a host emulation, so the
// identical code is exercised on the CPU by the dev simulation (never linked into libzkpair.so).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(ZKP_HOST_SIM)
#define ZKP_DEVICE_BUILD 1
#define ZKP_HD __device__ __forceinline__
#define ZKP_MEMBER __device__ __forceinline__
#define ZKP_HOSTDEV __host__ __device__ inline
#define ZKP_NOINLINE __device__ __noinline__
#define ZKP_CONST __device__ __constant__ const
#define ZKP_SIM_ASSERT(cond, what)
#else
#include <cstdio>
#include <cstdlib>
#define ZKP_HD static inline
#define ZKP_MEMBER inline
#define ZKP_HOSTDEV static inline
#define ZKP_NOINLINE static __attribute__((noinline))
#define ZKP_CONST static const
#define ZKP_SIM_ASSERT(cond, what)                                                       \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            fprintf(stderr, "ZKP bound violation: %s (%s:%d)\n", what, __FILE__, __LINE__); \
            abort();                                                                     \
        }                                                                                \
    } while (0)
#endif

#include "consts.cuh"

// Block-wide rendezvous of the converged kernels at loop boundaries: the warps of a block then run the same
// stretch of straight-line code at about the same time and share its instruction-cache lines (the Miller
// loop's hot code is ~85 KB against a 32 KB L1.5 I-cache).  level = how fine-grained the point is.
#if defined(ZKP_DEVICE_BUILD) && defined(ZKP_CONVERGED) && defined(ZKP_LOOP_SYNC)
#define ZKP_CODE_SYNC(level) do { if ((level) <= ZKP_LOOP_SYNC) __syncthreads(); } while (0)
#else
#define ZKP_CODE_SYNC(level) do { } while (0)
#endif

namespace zkp {

#define ZKP_NL 12

struct alignas(16) Fp {
    uint32_t l[ZKP_NL];
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
// host emulation of the PTX condition-code register (one flag per simulated lane = host thread)
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

// ------------------------------------------------------------------ lane pairing
//
// Two adjacent lanes (2k, 2k+1) of a warp cooperate on one pairing: every Fp2 value is split, the
// even lane holds c0 and the odd lane c1 (tower.cuh).  The only communication is a 12-word
// shfl.xor with the partner, synchronised on the pair's own two-lane mask so that pairs may
// diverge from each other (point generation, infinity handling) without deadlock.
#ifdef ZKP_DEVICE_BUILD
ZKP_HD int lane_par() { return (int)(threadIdx.x & 1u); }
#ifdef ZKP_CONVERGED
// translation units whose kernels keep all 32 lanes on one path (pairing_kernel.cu): plain SHFL
ZKP_HD unsigned pair_mask() { return 0xffffffffu; }
#else
ZKP_HD unsigned pair_mask() { return 3u << (threadIdx.x & 30u); }
#endif
ZKP_HD uint32_t word_xchg(uint32_t v) { return __shfl_xor_sync(pair_mask(), v, 1); }
// the even (PAR = 0) / odd (PAR = 1) lane's word, in both lanes of the pair
template <int PAR>
ZKP_HD uint32_t word_bcast(uint32_t v) { return __shfl_sync(pair_mask(), v, (int)((threadIdx.x & 30u) | PAR)); }
#else
// CPU dev simulation: the two lanes are two host threads in lock-step (tests/host_sim/sim.cpp)
extern thread_local int zkp_sim_par;
extern thread_local unsigned long long zkp_sim_macs;   // wide MACs issued by this simulated lane (work accounting)
uint32_t zkp_sim_word_xchg(uint32_t v);
void zkp_sim_xchg(void *buf, unsigned long bytes);
ZKP_HD int lane_par() { return zkp_sim_par; }
ZKP_HD uint32_t word_xchg(uint32_t v) { return zkp_sim_word_xchg(v); }
template <int PAR>
ZKP_HD uint32_t word_bcast(uint32_t v) { uint32_t o = zkp_sim_word_xchg(v); return zkp_sim_par == PAR ? v : o; }
#endif
ZKP_HD bool lane_or(bool x) { return (x | (word_xchg(x ? 1u : 0u) != 0)); }
// true when x holds in any lane that shares this lane's control flow: the whole block (or warp) in the
// converged kernels, the lane pair elsewhere -- i.e. a predicate every such lane may branch on together
ZKP_HD bool group_any(bool x) {
#if defined(ZKP_DEVICE_BUILD) && defined(ZKP_CONVERGED) && defined(ZKP_LOOP_SYNC)
    return __syncthreads_or(x) != 0;   // block-uniform: the guarded code may contain rendezvous points
#elif defined(ZKP_DEVICE_BUILD) && defined(ZKP_CONVERGED)
    return __any_sync(0xffffffffu, x) != 0;
#else
    return lane_or(x);
#endif
}
ZKP_HD bool lane_and(bool x) { return (x & (word_xchg(x ? 1u : 0u) != 0)); }

// ------------------------------------------------------------------ constants / trivial ops
ZKP_HD Fp fp_const(const uint32_t *k) {   // a constant in Montgomery form, value < p
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = k[i];
    return r;
}
ZKP_HD Fp fp_zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = 0;
    return r;
}
ZKP_HD Fp fp_one() { return fp_const(ZKP_ONE); }   // Montgomery one; canonical one is [1,0,..] (src/fp.rs:154-156)

// a >= k ?  (k = 12 constant words)
ZKP_HD bool fp_geq_const(const Fp &a, const uint32_t *k) {
    sub_cc(a.l[0], k[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) subc_cc(a.l[i], k[i]);
    return subc(0, 0) == 0;
}
#ifndef ZKP_DEVICE_BUILD
ZKP_HD bool fp_leq_2p(const Fp &a) { return !fp_geq_const(a, ZKP_2P1); }   // dev-simulation bound checks
ZKP_HD bool fp_leq_4p(const Fp &a) { return !fp_geq_const(a, ZKP_4P1); }
#endif

// plain 384-bit sum, no correction: the caller guarantees a + b < 2^384 and a consumer that
// accepts the larger value (one operand of a single-product mont_mul)
ZKP_HD Fp fp_add_lazy(const Fp &a, const Fp &b) {
    Fp r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[ZKP_NL - 1] = addc(a.l[ZKP_NL - 1], b.l[ZKP_NL - 1]);
    return r;
}
// a in [0, 4p] -> [0, 2p]: subtract 2p when that does not go negative
ZKP_HD Fp fp_correct(const Fp &s) {
    Fp t;
    t.l[0] = sub_cc(s.l[0], ZKP_2P[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) t.l[i] = subc_cc(s.l[i], ZKP_2P[i]);
    bool neg = subc(0, 0) != 0;
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = neg ? s.l[i] : t.l[i];
    return r;
}
// (a + b) mod p, 2p-redundant   -- value-equal mod p to src/fp.rs:351-368
ZKP_HD Fp fp_add(const Fp &a, const Fp &b) { return fp_correct(fp_add_lazy(a, b)); }
// (a - b) mod p, 2p-redundant   -- src/fp.rs:407-411
ZKP_HD Fp fp_sub(const Fp &a, const Fp &b) {
    Fp d;
    d.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) d.l[i] = subc_cc(a.l[i], b.l[i]);
#if defined(ZKP_SUB_NOBRANCH)
    // branch-free: d + 2p is formed unconditionally, one limb behind the subtraction chain (ptxas overlaps the two
    // carry chains), and selected by the borrow -- no BSSY / BRA / BSYNC around a second serial chain
    bool neg = subc(0, 0) != 0;
    Fp e;
    e.l[0] = add_cc(d.l[0], ZKP_2P[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) e.l[i] = addc_cc(d.l[i], ZKP_2P[i]);
    e.l[ZKP_NL - 1] = addc(d.l[ZKP_NL - 1], ZKP_2P[ZKP_NL - 1]);
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = neg ? e.l[i] : d.l[i];
    return r;
#else
    if (subc(0, 0) != 0) {   // a < b: add 2p back (a short predicated carry chain, no selects)
        d.l[0] = add_cc(d.l[0], ZKP_2P[0]);
#pragma unroll
        for (int i = 1; i < ZKP_NL - 1; i++) d.l[i] = addc_cc(d.l[i], ZKP_2P[i]);
        d.l[ZKP_NL - 1] = addc(d.l[ZKP_NL - 1], ZKP_2P[ZKP_NL - 1]);
    }
    return d;
#endif
}
// -a = 2p - a, in [0, 2p]                    -- src/fp.rs:381-405
ZKP_HD Fp fp_neg(const Fp &a) {
    Fp r;
    r.l[0] = sub_cc(ZKP_2P[0], a.l[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) r.l[i] = subc_cc(ZKP_2P[i], a.l[i]);
    r.l[ZKP_NL - 1] = subc(ZKP_2P[ZKP_NL - 1], a.l[ZKP_NL - 1]);
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
    Fp r = a;
    zkp_sim_xchg(&r, sizeof(Fp));
    return r;
#endif
}
// the value held by the pair's even (PAR = 0) or odd (PAR = 1) lane, delivered to both lanes
template <int PAR>
ZKP_HD Fp fp_bcast(const Fp &a) {
#ifdef ZKP_DEVICE_BUILD
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = word_bcast<PAR>(a.l[i]);
    return r;
#else
    Fp o = a;
    zkp_sim_xchg(&o, sizeof(Fp));
    return zkp_sim_par == PAR ? a : o;
#endif
}
// c ? a : b, limb-wise (c is lane-uniform per value, not per limb)
ZKP_HD Fp fp_select(bool c, const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}

// ------------------------------------------------------------------ Montgomery product
//
// CIOS over two interleaved accumulators.  X is the array in the "even" role (words 0..11 of the
// running sum), Y the one in the "odd" role (words 1..12).  After every row the sum is divisible
// by 2^32; instead of shifting registers the arrays swap roles (the old odd array is the new even
// one, the old even array moves down 64 bits inside the first chain of the next row).

// x[j..j+1] += k[j+OFF]*m for j = 0,2,..,10 (one carry chain, 6 wide MACs); leaves carry-out in CF
template <int OFF>
ZKP_HD void chain_mad(uint32_t *x, const uint32_t *k, uint32_t m) {
    x[0] = mad_lo_cc(k[OFF], m, x[0]);
    x[1] = madc_hi_cc(k[OFF], m, x[1]);
#pragma unroll
    for (int j = 2; j < ZKP_NL; j += 2) {
        x[j] = madc_lo_cc(k[j + OFF], m, x[j]);
        x[j + 1] = madc_hi_cc(k[j + OFF], m, x[j + 1]);
    }
}
// y >>= 64 bits; y[j..j+1] += a[j+1]*m for j = 0,2,..,10, consuming the incoming carry
ZKP_HD void chain_mad_rshift(uint32_t *y, const uint32_t *a, uint32_t m) {
#pragma unroll
    for (int j = 0; j < ZKP_NL - 2; j += 2) {
        y[j] = madc_lo_cc(a[j + 1], m, y[j + 2]);
        y[j + 1] = madc_hi_cc(a[j + 1], m, y[j + 3]);
    }
    y[ZKP_NL - 2] = madc_lo_cc(a[ZKP_NL - 1], m, 0);
    y[ZKP_NL - 1] = madc_hi(a[ZKP_NL - 1], m, 0);
}
// x/y += k*m over both chains; the carry out of the even chain lands in the top odd word.  The
// carry out of the odd chain is zero by the operand bounds (top words: u,w <= 0x68044800,
// p = 0x1a0111ea; their sum stays below 2^32).
ZKP_HD void row_mad(uint32_t *x, uint32_t *y, const uint32_t *k, uint32_t m) {
    chain_mad<1>(y, k, m);
    chain_mad<0>(x, k, m);
    y[ZKP_NL - 1] = addc(y[ZKP_NL - 1], 0);
}
// reduction half of a row: m = x0 * n0'; x/y += p*m
ZKP_HD void row_reduce(uint32_t *x, uint32_t *y) {
    uint32_t m = x[0] * ZKP_N0INV;
    row_mad(x, y, ZKP_P, m);
}
// first row: disjoint 64-bit products, no carries
ZKP_HD void row_first(uint32_t *x, uint32_t *y, const uint32_t *a, uint32_t b0) {
#pragma unroll
    for (int j = 0; j < ZKP_NL; j += 2) {
        uint64_t e = (uint64_t)a[j] * b0;
        uint64_t o = (uint64_t)a[j + 1] * b0;
        x[j] = (uint32_t)e;
        x[j + 1] = (uint32_t)(e >> 32);
        y[j] = (uint32_t)o;
        y[j + 1] = (uint32_t)(o >> 32);
    }
}
// later rows: x = array entering the even role, y = array entering the odd role
ZKP_HD void row_next(uint32_t *x, uint32_t *y, const uint32_t *a, uint32_t bi) {
    x[0] = add_cc(x[0], y[1]);
    chain_mad_rshift(y, a, bi);
    chain_mad<0>(x, a, bi);
    y[ZKP_NL - 1] = addc(y[ZKP_NL - 1], 0);
}
// after row 11 the even role is `od` (od[0] == 0): T = ev + (od >> 32)
ZKP_HD Fp mont_finish(const uint32_t *ev, const uint32_t *od) {
    Fp r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < ZKP_NL - 1; k++) r.l[k] = addc_cc(ev[k], od[k + 1]);
    r.l[ZKP_NL - 1] = addc(ev[ZKP_NL - 1], 0);
    return r;
}

// a*b/2^384 mod p in [0, 2p).  Needs a*b < p * 2^384: both <= 2p, or one <= 4p and the other <= 2p.
// 300 wide MACs.  `b` may live in constant memory (its words are only used as scalars).
ZKP_HD Fp mont_mul(const Fp &a, const uint32_t *b) {
    uint32_t ev[ZKP_NL], od[ZKP_NL];
    row_first(ev, od, a.l, b[0]);
    row_reduce(ev, od);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i += 2) {
        row_next(od, ev, a.l, b[i]);
        row_reduce(od, ev);
        if (i + 1 < ZKP_NL) {
            row_next(ev, od, a.l, b[i + 1]);
            row_reduce(ev, od);
        }
    }
    return mont_finish(ev, od);
}
ZKP_HD Fp mont_mul(const Fp &a, const Fp &b) {
#ifndef ZKP_DEVICE_BUILD
    zkp_sim_macs += 300;
    ZKP_SIM_ASSERT((fp_leq_2p(a) && fp_leq_4p(b)) || (fp_leq_4p(a) && fp_leq_2p(b)), "mont_mul operand bound");
#endif
    return mont_mul(a, b.l);
}
// (u*v + w*z)/2^384 mod p in [0, 2p) with ONE reduction (lazy "sum of products"): 288 + 156 wide
// MACs.  This is one lane's half of an Fp2 product.  All four operands <= 2p.
ZKP_HD Fp mont_mul2(const Fp &u, const Fp &v, const Fp &w, const Fp &z) {
#ifndef ZKP_DEVICE_BUILD
    zkp_sim_macs += 444;
    ZKP_SIM_ASSERT(fp_leq_2p(u) && fp_leq_2p(v) && fp_leq_2p(w) && fp_leq_2p(z), "mont_mul2 operand bound");
#endif
    uint32_t ev[ZKP_NL], od[ZKP_NL];
    row_first(ev, od, u.l, v.l[0]);
    row_mad(ev, od, w.l, z.l[0]);
    row_reduce(ev, od);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i += 2) {
        row_next(od, ev, u.l, v.l[i]);
        row_mad(od, ev, w.l, z.l[i]);
        row_reduce(od, ev);
        if (i + 1 < ZKP_NL) {
            row_next(ev, od, u.l, v.l[i + 1]);
            row_mad(ev, od, w.l, z.l[i + 1]);
            row_reduce(ev, od);
        }
    }
    return mont_finish(ev, od);
}
// Value-equivalent (after conversion) to src/fp.rs:413-434 / :452-455.
ZKP_HD Fp fp_mul(const Fp &a, const Fp &b) { return mont_mul(a, b); }

// ------------------------------------------------------------------ unreduced ("wide") values: lazy reduction
//
// A Montgomery reduction is 156 of the 300 wide MACs of a product.  Where several products are only ever added to or
// subtracted from each other before anybody looks at them (the Karatsuba recombination of an Fp6 product, an Fp4
// square), the products are kept UNREDUCED -- 24 words, at most 8 p^2 each for 2p-redundant operands -- recombined by
// plain 768-bit additions and subtractions modulo 2^768 (no correction steps), and only the results are reduced:
// 3 reductions instead of 6 per lane for an Fp6 product (tower.cuh).  Intermediate values may wrap around modulo
// 2^768; before the reduction a constant multiple of p^2 (ZKP_P2X8 etc., = 0 mod p) is added that makes the TRUE value
// non-negative, and the call sites keep it below 2^768 = 96.9 p^2 (bounds stated at each call site, interval
// arithmetic over the operand bound 2p).
struct alignas(16) FpW {
    uint32_t l[2 * ZKP_NL];
};
// One row of a plain product, E/O += a * b * 2^(32 i), on split accumulators like the CIOS rows above: E holds the
// 64-bit columns at even words (E[k] = word k), O the columns at odd words (O[k] = word k + 1).  The carry out of
// a chain lands in a word no earlier row has used for data (it holds at most the carries of the neighbouring rows).
ZKP_HD void wide_row(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t b, int i) {
    // The chain over the odd limbs ends with a[11] * b < 2^63 (a <= 4p: a[11] <= 0x68044800) on top of a column that
    // so far holds only a few carries: no carry out.  The chain over the even limbs ends with a[10] * b: its carry is
    // absorbed by the next word up, which is either untouched or holds the carries of the neighbouring rows.
    if ((i & 1) == 0) {
        chain_mad<1>(O + i, a, b);
        chain_mad<0>(E + i, a, b);
        E[i + ZKP_NL] = addc(E[i + ZKP_NL], 0);
    } else {
        chain_mad<1>(E + i + 1, a, b);
        chain_mad<0>(O + i - 1, a, b);
        O[i + ZKP_NL - 1] = addc(O[i + ZKP_NL - 1], 0);
    }
}
ZKP_HD FpW wide_merge(const uint32_t *E, const uint32_t *O) {
    FpW t;
    t.l[0] = E[0];
    t.l[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * ZKP_NL - 1; k++) t.l[k] = addc_cc(E[k], O[k - 1]);
    t.l[2 * ZKP_NL - 1] = addc(E[2 * ZKP_NL - 1], O[2 * ZKP_NL - 2]);
    return t;
}
// u * v + w * z, unreduced: 288 wide MACs.  Operands <= 2p each (u, w may be <= 4p): <= 16 p^2.
ZKP_HD FpW mul_wide2(const Fp &u, const Fp &v, const Fp &w, const Fp &z) {
#ifndef ZKP_DEVICE_BUILD
    zkp_sim_macs += 288;
    ZKP_SIM_ASSERT(fp_leq_4p(u) && fp_leq_2p(v) && fp_leq_4p(w) && fp_leq_2p(z), "mul_wide2 operand bound");
#endif
    uint32_t E[2 * ZKP_NL], O[2 * ZKP_NL];
#pragma unroll
    for (int k = 0; k < 2 * ZKP_NL; k++) E[k] = O[k] = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        wide_row(E, O, u.l, v.l[i], i);
        wide_row(E, O, w.l, z.l[i], i);
    }
    return wide_merge(E, O);
}
// u * v, unreduced: 144 wide MACs.  u <= 4p, v <= 2p.
ZKP_HD FpW mul_wide(const Fp &u, const Fp &v) {
#ifndef ZKP_DEVICE_BUILD
    zkp_sim_macs += 144;
    ZKP_SIM_ASSERT(fp_leq_4p(u) && fp_leq_2p(v), "mul_wide operand bound");
#endif
    uint32_t E[2 * ZKP_NL], O[2 * ZKP_NL];
#pragma unroll
    for (int k = 0; k < 2 * ZKP_NL; k++) E[k] = O[k] = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) wide_row(E, O, u.l, v.l[i], i);
    return wide_merge(E, O);
}
// 768-bit arithmetic modulo 2^768 (wrap-around is harmless as long as the value that is finally reduced is in range)
ZKP_HD FpW fpw_add(const FpW &a, const FpW &b) {
    FpW r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 2 * ZKP_NL - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[2 * ZKP_NL - 1] = addc(a.l[2 * ZKP_NL - 1], b.l[2 * ZKP_NL - 1]);
    return r;
}
ZKP_HD FpW fpw_sub(const FpW &a, const FpW &b) {
    FpW r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 2 * ZKP_NL - 1; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    r.l[2 * ZKP_NL - 1] = subc(a.l[2 * ZKP_NL - 1], b.l[2 * ZKP_NL - 1]);
    return r;
}
ZKP_HD FpW fpw_add_const(const FpW &a, const uint32_t *k) {   // k = 24 constant words
    FpW r;
    r.l[0] = add_cc(a.l[0], k[0]);
#pragma unroll
    for (int i = 1; i < 2 * ZKP_NL - 1; i++) r.l[i] = addc_cc(a.l[i], k[i]);
    r.l[2 * ZKP_NL - 1] = addc(a.l[2 * ZKP_NL - 1], k[2 * ZKP_NL - 1]);
    return r;
}
ZKP_HD FpW fpw_xchg(const FpW &a) {
#ifdef ZKP_DEVICE_BUILD
    FpW r;
#pragma unroll
    for (int i = 0; i < 2 * ZKP_NL; i++) r.l[i] = word_xchg(a.l[i]);
    return r;
#else
    FpW r = a;
    zkp_sim_xchg(&r, sizeof(FpW));
    return r;
#endif
}
// One row of a reduction WITHOUT a product row in front of it: x enters the even role, y the odd role (and moves
// down 64 bits inside its chain, exactly as in row_next); m clears the low word.
ZKP_HD void row_shift_reduce(uint32_t *x, uint32_t *y) {
    x[0] = add_cc(x[0], y[1]);
    uint32_t m = x[0] * ZKP_N0INV;
    chain_mad_rshift(y, ZKP_P, m);
    chain_mad<0>(x, ZKP_P, m);
    y[ZKP_NL - 1] = addc(y[ZKP_NL - 1], 0);
}
// T / 2^384 mod p for an unreduced T < 2^768: 156 wide MACs.  T = Tlo + Thi 2^384: the low half is reduced
// ((Tlo + m p) / 2^384 <= p), the high half added on top: result < T / 2^384 + p + 1 -- the CALLER corrects
// it back to [0, 2p] according to its bound on T (fp_correct / fp_correct8).
ZKP_HD Fp mont_redc(const FpW &t) {
#ifndef ZKP_DEVICE_BUILD
    zkp_sim_macs += 156;
#endif
    uint32_t ev[ZKP_NL], od[ZKP_NL];
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        ev[i] = t.l[i];
        od[i] = 0;
    }
    row_reduce(ev, od);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i += 2) {
        row_shift_reduce(od, ev);
        if (i + 1 < ZKP_NL) row_shift_reduce(ev, od);
    }
    Fp r = mont_finish(ev, od);
    r.l[0] = add_cc(r.l[0], t.l[ZKP_NL]);
#pragma unroll
    for (int i = 1; i < ZKP_NL - 1; i++) r.l[i] = addc_cc(r.l[i], t.l[ZKP_NL + i]);
    r.l[ZKP_NL - 1] = addc(r.l[ZKP_NL - 1], t.l[2 * ZKP_NL - 1]);
    return r;
}
// a in [0, 8p) -> [0, 2p]: subtract 4p, then 2p, each when that does not go negative
ZKP_HD Fp fp_correct8(const Fp &s) {
    Fp t;
    t.l[0] = sub_cc(s.l[0], ZKP_4P[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) t.l[i] = subc_cc(s.l[i], ZKP_4P[i]);
    bool neg = subc(0, 0) != 0;
    Fp r;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) r.l[i] = neg ? s.l[i] : t.l[i];
    return fp_correct(r);
}

// ------------------------------------------------------------------ boundary conversions
//
// Canonical form = 12 saturated 32-bit words (= the six u64 limbs of src/fp.rs:24), value in [0,p).

// canonical words -> Montgomery Fp; sets bad when w >= p (such inputs are rejected at the boundary
// because the reference's neg is undefined there, src/fp.rs:383-405)
ZKP_HD Fp fp_from_words(const uint32_t *w, bool &bad) {
    Fp a;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) a.l[i] = w[i];
    bad = bad | fp_geq_const(a, ZKP_P);
    return mont_mul(a, ZKP_R2);   // any 384-bit a: a * R2 < 2^384 * p, result in [0, 2p)
}
// Montgomery Fp (2p-redundant) -> canonical words in [0,p)
ZKP_HD void fp_to_words(uint32_t *w, const Fp &m) {
    uint32_t one[ZKP_NL];
    one[0] = 1;
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) one[i] = 0;
    Fp t = mont_mul(m, one);   // (m + k*p)/R <= p
    Fp d;
    d.l[0] = sub_cc(t.l[0], ZKP_P[0]);
#pragma unroll
    for (int i = 1; i < ZKP_NL; i++) d.l[i] = subc_cc(t.l[i], ZKP_P[i]);
    bool lt = subc(0, 0) != 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) w[i] = lt ? t.l[i] : d.l[i];
}
// value == 0 mod p for a 2p-redundant value: the only representatives are 0, p and 2p
ZKP_HD bool fp_is_zero(const Fp &a) {
    uint32_t z = 0, zp = 0, z2 = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        z |= a.l[i];
        zp |= a.l[i] ^ ZKP_P[i];
        z2 |= a.l[i] ^ ZKP_2P[i];
    }
    return (z == 0) | (zp == 0) | (z2 == 0);
}

}  // namespace zkp
