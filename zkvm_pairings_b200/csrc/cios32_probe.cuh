// Carry-chained 32-bit wide-MAC primitives (PTX mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X).
//
// This is the inner step of the v0 design (12 x 32-bit saturated limbs, CIOS rows as carry chains,
// what the north star literally asks for).  Measured on B200 the carry form issues at HALF the rate
// of a plain IMAD.WIDE.U32 (9.1 vs 18.2 T MAC/s, profiles/r1a_v0_saturated_cios_ncu_summary.txt), so
// the product arithmetic moved to carry-free 28-bit limbs (fp.cuh).  Only the roofline probe
// zkp_imad_peak(kind = 2) still uses this file, to keep that measurement reproducible.
#pragma once
#include <stdint.h>
namespace zkp {
// ------------------------------------------------------------------ carry-chain primitives
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// x[j..j+1] += k[j]*m for j = 0,2,..,10 (one carry chain, 6 wide MACs); leaves carry-out in CF
template <int OFF>
__device__ __forceinline__ void chain_mad(uint32_t *x, const uint32_t *k, uint32_t m) {
    x[0] = mad_lo_cc(k[OFF], m, x[0]);
    x[1] = madc_hi_cc(k[OFF], m, x[1]);
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        x[j] = madc_lo_cc(k[j + OFF], m, x[j]);
        x[j + 1] = madc_hi_cc(k[j + OFF], m, x[j + 1]);
    }
}
}  // namespace zkp
