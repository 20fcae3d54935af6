// k_pairing: the hot kernel (Miller loop and/or final exponentiation), compiled on its own with
// ZKP_CONVERGED: its control flow is identical in all 32 lanes of a warp (tail lanes recompute the
// last element, points at infinity are handled by selects), so the lane-pair exchanges are plain
// full-mask SHFLs instead of the match/vote-guarded pair-masked ones the divergent kernels need.
#include <cuda_runtime.h>

#include <atomic>

#define ZKP_CONVERGED 1
#define zkp zkp_conv   // this unit's own copy of the device functions (kernels.cu holds the pair-masked one)
// Block-wide rendezvous points (ZKP_CODE_SYNC, fp.cuh) at the entry of every Fp6-level body, and with them FOUR
// blocks per SM at 128 registers: once the four warps of a block share their instruction fetches, occupancy pays
// where it used to lose to instruction-fetch stalls.  Miller loop only, 2^20 (profiles/r1o_occupancy_variants.txt):
//   2 blocks/SM (244 registers): no points 311.7 ms, once per iteration 305.6, Fp6-level 310.4
//   3 blocks/SM (168): once per iteration 287.5, Fp6-level 289.1
//   4 blocks/SM (128): once per iteration 287.1, Fp6-level 280.2, Fp2-level 286.3;  5 blocks 307.5, 6 blocks 315.9
// With the lazy forms below (hot code ~100 KB) a rendezvous at every Fp2-level body (5) measures slightly better in steady state --
// three interleaved repetitions at 2^20 (profiles/r2z_sync_order_variants.txt): pairing 518.3..518.4 -> 517.3..517.5 ms, together with
// ZKP_LAZY_ORDER 516.6..516.8 ms (-0.3 %); Miller loop alone 272.9..274.8 -> 270.5..270.7 ms; no difference at 2^16.  The counters of a
// single cold launch under ncu do not show why (instruction-fetch stalls 0.43 -> 0.50 per issue, profiles/r2aa_ncu_summary.txt).
#ifndef ZKP_MILLER_SYNC
#define ZKP_MILLER_SYNC 5
#endif
#define ZKP_LOOP_SYNC ZKP_MILLER_SYNC
// Lazy reduction (tower.cuh ZKP_LAZY, fp.cuh FpW) in THIS unit only: ZKP_MILLER_LAZY = 3 (shipped) recombines UNREDUCED Fp2
// products in the Fp6 products and the sparse line products of the Miller loop -- 3 reductions instead of 6 / 5 per lane,
// -9.8 % wide MACs per Miller loop, bit-identical results.  Measured at 2^20, three interleaved repetitions
// (profiles/r2l_lazy_reduction.txt, call r2t): pairing 520.5 -> 516.4 ms (-0.8 %), 2^16 pairings 34.8 -> 34.4 ms, 4-pair checks
// 265.2..268.0 -> 264.2 ms, with prepared tables 220.1..221.8 -> 214.8 ms (-2.7 %).  The final-exponentiation unit keeps the
// reduced forms (the lazy Fp4 squares measured +2 % there: larger hot loop, more spills).
#ifndef ZKP_MILLER_LAZY
#define ZKP_MILLER_LAZY 3
#endif
#ifndef ZKP_LAZY
#define ZKP_LAZY ZKP_MILLER_LAZY
#endif
#include "../../include/zkpair.h"
#include "fe_scratch.cuh"

// fe_kernel.cu
cudaError_t zkp_launch_fe_stages(void *scratch, size_t n, uint64_t *out, uint8_t *is_one, cudaStream_t st, const ZkpFeAux *aux,
                                 int *launches, int forked);
size_t zkp_fe_split_point(size_t n, const ZkpFeAux *aux);

#ifndef ZKP_TPB
#define ZKP_TPB 128           // threads per block
#endif
#ifndef ZKP_MILLER_SPLIT
#define ZKP_MILLER_SPLIT 0    // 1: the two-stream halves of the final exponentiation start at the Miller kernel (see zkp_launch_k_pairing)
#endif
#ifndef ZKP_MIN_BLOCKS
#define ZKP_MIN_BLOCKS 4      // resident blocks per SM the register allocator must allow (Miller kernel)
#endif

using namespace zkp;

// ZKP_SMEM_STATE -- which per-thread state lives in SHARED memory (a per-thread slice at a stride of 76 or 108 words,
// both 12 mod 32, so that the 128-bit accesses of a quarter-warp fall into disjoint bank groups):
//   0  nothing: f, R and every temporary in thread-local memory (round 1)
//   1  the Miller accumulator f (304 B / thread)
//   2  f and, for single pairings, the G2 accumulator R (432 B)
//   3  f and the ONE Fp6 temporary of the in-place Fp12 operations (432 B; tower.cuh ZKP_INPLACE12) -- shipped
// Measured, Miller loop at 2^20 (profiles/r2b_smem_state_variants.txt, r2d_miller_ab.txt, r2f_ncu_summary.txt):
//   280.5 ms (0) -> 276.6 (1) / 276.9 (2) -> 273.7 (3); 3 blocks per SM at 168 registers 286.2, 5 blocks with state 1 300.2.
//   DRAM traffic of the kernel per 2^16 pairings 15.76 GB (0) -> 6.26 (2) -> 0.92 (3); local loads 195.1M -> 146.5M -> 77.8M,
//   local stores 100.1M -> 84.0M -> 42.4M; long-scoreboard stall 0.97 -> 0.86 -> 0.45 per issue; fmaheavy 79.9 -> 80.1 -> 82.2 %.
#ifndef ZKP_SMEM_STATE
#define ZKP_SMEM_STATE 3
#endif
#if ZKP_SMEM_STATE >= 3 && !ZKP_INPLACE12
#error "ZKP_SMEM_STATE=3 keeps the in-place temporary in shared memory: needs ZKP_INPLACE12=1"
#endif
#define ZKP_SMEM_STRIDE(K) ((ZKP_SMEM_STATE >= 3 || ((K) == 1 && ZKP_SMEM_STATE == 2)) ? 432 : 304)

// mode: bit0 Miller loop, bit1 first half of the final exponentiation.  One lane pair per check of
// k (<= K) pairs.  Without bit1 the Miller output is stored canonically to `out`.
template <int K>
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_pairing(int mode, const uint64_t *__restrict__ g1, const uint8_t *__restrict__ g1inf,
          const uint64_t *__restrict__ g2, const uint8_t *__restrict__ g2inf, int k,
          const uint64_t *__restrict__ in12, uint64_t *__restrict__ out, uint8_t *__restrict__ is_one,
          uint32_t *err, size_t i0, size_t n, FeScratch fs, const Fp *__restrict__ tab, const uint8_t *__restrict__ tabinf, int kf) {
    // this launch covers the checks [i0, n) of the batch
    size_t i = i0 + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1);
    bool live = i < n;
    if (!live) i = n - 1;   // stay converged: redo the last element, store nothing
    size_t e = i * (size_t)k, e2 = i * (size_t)(k - kf);   // the last kf pairs of a check use prepared G2 tables
    bool bad = false;
#if ZKP_SMEM_STATE
    // this thread's slice of shared memory: f first, then R (state 2) or the in-place temporary (state 3)
    extern __shared__ uint4 zkp_smem[];
    char *slice = reinterpret_cast<char *>(zkp_smem) + (size_t)threadIdx.x * ZKP_SMEM_STRIDE(K);
    Fp12 &f = *reinterpret_cast<Fp12 *>(slice);
    G2P *rs_ext = (K == 1 && ZKP_SMEM_STATE == 2) ? reinterpret_cast<G2P *>(slice + sizeof(Fp12)) : nullptr;
    Fp6 *tmp_ext = ZKP_SMEM_STATE >= 3 ? reinterpret_cast<Fp6 *>(slice + sizeof(Fp12)) : nullptr;
#else
    Fp12 f;
    G2P *rs_ext = nullptr;
    Fp6 *tmp_ext = nullptr;
#endif
    pairing_front<K>(f, bad, mode, g1 ? g1 + 12 * e : nullptr, g1inf ? g1inf + e : nullptr, g2 ? g2 + 24 * e2 : nullptr,
                     g2inf ? g2inf + e2 : nullptr, k, in12 ? in12 + 72 * i : nullptr, tab, tabinf, kf, rs_ext, tmp_ext);
    if (mode & ZKP_DO_FINAL_EXP) {
        FeState s;
        Fp nrm = fe_prepare(s, f);
        if (live) {
            size_t lane = 2 * i + lane_par();
            park_fp12(fs, lane, ZKP_SLOT_F, f);
            fs.lanes[(size_t)(ZKP_SLOT_FES + 0) * fs.n2 + lane] = s.c.c0.c;
            fs.lanes[(size_t)(ZKP_SLOT_FES + 1) * fs.n2 + lane] = s.c.c1.c;
            fs.lanes[(size_t)(ZKP_SLOT_FES + 2) * fs.n2 + lane] = s.c.c2.c;
            fs.lanes[(size_t)(ZKP_SLOT_FES + 3) * fs.n2 + lane] = s.t.c;
            if (lane_par() == 0) fs.norm[i] = nrm;
        }
    } else {
        bool one = store_fp12(out + 72 * i, f, live);
        if (is_one && live && lane_par() == 0) is_one[i] = one ? 1 : 0;
    }
    if (lane_or(bad) && err && live && lane_par() == 0) atomicOr(err, 1u);
}

extern "C" void zkp_miller_geometry(int *tpb, int *blocks, int *sync) { *tpb = ZKP_TPB; *blocks = ZKP_MIN_BLOCKS; *sync = ZKP_MILLER_SYNC; }

static int pair_capacity(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

// `scratch`: zkp_fe_scratch_bytes(n) device bytes when mode has bit1 set (else unused)
cudaError_t zkp_launch_k_pairing(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                 size_t n, int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one, uint32_t *err,
                                 void *scratch, const void *tab, const uint8_t *tabinf, int kf, cudaStream_t st, const ZkpFeAux *aux, int *launches) {
    if (n == 0) return cudaSuccess;
    FeScratch fs;
    fs.lanes = (Fp *)scratch;
    fs.norm = fs.lanes ? fs.lanes + 2 * n * ZKP_FE_LANE_FP : nullptr;
    fs.n2 = 2 * n;
    dim3 b(ZKP_TPB);
#if ZKP_SMEM_STATE
    // function attributes are per DEVICE: opt in to > 48 KB of dynamic shared memory once on each device that launches
    static std::atomic<unsigned long long> attr_done{0};
    int cur = 0;
    cudaGetDevice(&cur);
    if (!((attr_done.load() >> (cur & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_pairing<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZKP_SMEM_STRIDE(1) * ZKP_TPB);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pairing<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZKP_SMEM_STRIDE(2) * ZKP_TPB);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pairing<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZKP_SMEM_STRIDE(4) * ZKP_TPB);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pairing<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZKP_SMEM_STRIDE(8) * ZKP_TPB);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(k_pairing<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_pairing<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_pairing<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_pairing<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_done.fetch_or(1ull << (cur & 63));
    }
#define ZKP_SM(K) (size_t)(ZKP_SMEM_STRIDE(K) * ZKP_TPB)
#else
#define ZKP_SM(K) 0
#endif
    // ZKP_MILLER_SPLIT: a batch whose final exponentiation runs as two halves on two streams (fe_kernel.cu) splits already HERE,
    // so that the tail of the first half's Miller kernel is covered by the second half's blocks and the second half's tail by the
    // first stage kernels of the first half.
    size_t na = n;
    int forked = 0;
#if ZKP_MILLER_SPLIT
    if (mode & ZKP_DO_FINAL_EXP) na = zkp_fe_split_point(n, aux);
    if (na < n) {
        cudaError_t e = cudaEventRecord(aux->fork, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(aux->s2, aux->fork, 0);
        if (e != cudaSuccess) return e;
        forked = 1;
    }
#endif
    *launches = 0;
    for (int half = 0; half < (na < n ? 2 : 1); half++) {
        size_t lo = half ? na : 0, hi = half ? n : na;
        cudaStream_t sh = half ? aux->s2 : st;
        dim3 g((unsigned)((2 * (hi - lo) + ZKP_TPB - 1) / ZKP_TPB));
        switch (pair_capacity(k)) {
            case 1: k_pairing<1><<<g, b, ZKP_SM(1), sh>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, lo, hi, fs, (const Fp *)tab, tabinf, kf); break;
            case 2: k_pairing<2><<<g, b, ZKP_SM(2), sh>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, lo, hi, fs, (const Fp *)tab, tabinf, kf); break;
            case 4: k_pairing<4><<<g, b, ZKP_SM(4), sh>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, lo, hi, fs, (const Fp *)tab, tabinf, kf); break;
            default: k_pairing<8><<<g, b, ZKP_SM(8), sh>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, lo, hi, fs, (const Fp *)tab, tabinf, kf); break;
        }
        *launches += 1;
    }
    if (mode & ZKP_DO_FINAL_EXP) {
        int nl = 0;
        cudaError_t rc = zkp_launch_fe_stages(scratch, n, out, is_one, st, aux, &nl, forked);
        *launches += nl;
        if (rc != cudaSuccess) return rc;
    }
    return cudaGetLastError();
}
