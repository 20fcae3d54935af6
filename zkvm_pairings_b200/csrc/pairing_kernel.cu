// k_pairing: the hot kernel (Miller loop and/or final exponentiation), compiled on its own with
// ZKP_CONVERGED: its control flow is identical in all 32 lanes of a warp (tail lanes recompute the
// last element, points at infinity are handled by selects), so the lane-pair exchanges are plain
// full-mask SHFLs instead of the match/vote-guarded pair-masked ones the divergent kernels need.
#include <cuda_runtime.h>

#define ZKP_CONVERGED 1
#define zkp zkp_conv   // this unit's own copy of the device functions (kernels.cu holds the pair-masked one)
#include "../../include/zkpair.h"
#include "ops.cuh"

#ifndef ZKP_TPB
#define ZKP_TPB 128           // threads per block
#endif
#ifndef ZKP_MIN_BLOCKS
#define ZKP_MIN_BLOCKS 2      // resident blocks per SM the register allocator must allow (Miller kernel)
#endif
#ifndef ZKP_MIN_BLOCKS_FE
#define ZKP_MIN_BLOCKS_FE 3   // same for the final-exponentiation kernel (measured: 93.9 ms vs 96.4 at 2, 2^18)
#endif

using namespace zkp;

// ------------------------------------------------------------------ final-exponentiation scratch
//
// The final exponentiation is three launches: k_pairing (Miller loop and/or load, then fe_prepare)
// -> k_fe_batch_inv -> k_fe_finish.  Between them each lane parks its half of f (6 Fp), of the
// inverse cofactors (3 Fp) and of t (1 Fp) in `lanes`, and the pair's norm in `norm` -- internal
// Montgomery limbs, never seen by the caller.
#define ZKP_FE_LANE_FP 10
struct FeScratch {
    Fp *lanes;   // [2 * n][ZKP_FE_LANE_FP]
    Fp *norm;    // [n], replaced by its inverse in place
};
extern "C" size_t zkp_fe_scratch_bytes(size_t n) { return n * (2 * ZKP_FE_LANE_FP + 1) * sizeof(Fp); }

// mode: bit0 Miller loop, bit1 first half of the final exponentiation.  One lane pair per check of
// k (<= K) pairs.  Without bit1 the Miller output is stored canonically to `out`.
template <int K>
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_pairing(int mode, const uint64_t *__restrict__ g1, const uint8_t *__restrict__ g1inf,
          const uint64_t *__restrict__ g2, const uint8_t *__restrict__ g2inf, int k,
          const uint64_t *__restrict__ in12, uint64_t *__restrict__ out, uint8_t *__restrict__ is_one,
          uint32_t *err, size_t n, FeScratch fs, const Fp *__restrict__ tab, const uint8_t *__restrict__ tabinf, int kf) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    bool live = i < n;
    if (!live) i = n - 1;   // stay converged: redo the last element, store nothing
    size_t e = i * (size_t)k, e2 = i * (size_t)(k - kf);   // the last kf pairs of a check use prepared G2 tables
    bool bad = false;
    Fp12 f;
    pairing_front<K>(f, bad, mode, g1 ? g1 + 12 * e : nullptr, g1inf ? g1inf + e : nullptr, g2 ? g2 + 24 * e2 : nullptr,
                     g2inf ? g2inf + e2 : nullptr, k, in12 ? in12 + 72 * i : nullptr, tab, tabinf, kf);
    if (mode & ZKP_DO_FINAL_EXP) {
        FeState s;
        Fp nrm = fe_prepare(s, f);
        if (live) {
            Fp *L = fs.lanes + (2 * i + lane_par()) * ZKP_FE_LANE_FP;
            const Fp2 *c = &f.c0.c0;
#pragma unroll
            for (int j = 0; j < 6; j++) L[j] = c[j].c;
            L[6] = s.c.c0.c; L[7] = s.c.c1.c; L[8] = s.c.c2.c; L[9] = s.t.c;
            if (lane_par() == 0) fs.norm[i] = nrm;
        }
    } else {
        bool one = store_fp12(out + 72 * i, f, live);
        if (is_one && live && lane_par() == 0) is_one[i] = one ? 1 : 0;
    }
    if (lane_or(bad) && err && live && lane_par() == 0) atomicOr(err, 1u);
}

// norm[i] <- 1 / norm[i]: every thread inverts a run of ZKP_INV_RUN norms with one Fermat ladder
#ifndef ZKP_INV_RUN
#define ZKP_INV_RUN 16
#endif
__global__ void __launch_bounds__(128) k_fe_batch_inv(Fp *norm, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * ZKP_INV_RUN;
    if (lo >= n) return;
    int cnt = (int)(n - lo < ZKP_INV_RUN ? n - lo : ZKP_INV_RUN);
    Fp pre[ZKP_INV_RUN];
    fp_batch_inv(norm + lo, pre, cnt);
}

// second half of the final exponentiation
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS_FE)
k_fe_finish(FeScratch fs, uint64_t *__restrict__ out, uint8_t *__restrict__ is_one, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    bool live = i < n;
    if (!live) i = n - 1;
    const Fp *L = fs.lanes + (2 * i + lane_par()) * ZKP_FE_LANE_FP;
    Fp12 f;
    FeState s;
    Fp2 *c = &f.c0.c0;
#pragma unroll
    for (int j = 0; j < 6; j++) c[j].c = L[j];
    s.c.c0.c = L[6]; s.c.c1.c = L[7]; s.c.c2.c = L[8]; s.t.c = L[9];
    Fp ninv = fs.norm[i];
    fe_finish(f, f, s, ninv);
    bool one = store_fp12(out + 72 * i, f, live);
    if (is_one && live && lane_par() == 0) is_one[i] = one ? 1 : 0;
}

static int pair_capacity(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

// `scratch`: zkp_fe_scratch_bytes(n) device bytes when mode has bit1 set (else unused)
cudaError_t zkp_launch_k_pairing(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                 size_t n, int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one, uint32_t *err,
                                 void *scratch, const void *tab, const uint8_t *tabinf, int kf, cudaStream_t st, int *launches) {
    if (n == 0) return cudaSuccess;
    FeScratch fs;
    fs.lanes = (Fp *)scratch;
    fs.norm = fs.lanes ? fs.lanes + 2 * n * ZKP_FE_LANE_FP : nullptr;
    dim3 g((unsigned)((2 * n + ZKP_TPB - 1) / ZKP_TPB)), b(ZKP_TPB);
    switch (pair_capacity(k)) {
        case 1: k_pairing<1><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        case 2: k_pairing<2><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        case 4: k_pairing<4><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        default: k_pairing<8><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
    }
    *launches = 1;
    if (mode & ZKP_DO_FINAL_EXP) {
        size_t threads = (n + ZKP_INV_RUN - 1) / ZKP_INV_RUN;
        k_fe_batch_inv<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(fs.norm, n);
        k_fe_finish<<<g, b, 0, st>>>(fs, out, is_one, n);
        *launches = 3;
    }
    return cudaGetLastError();
}
