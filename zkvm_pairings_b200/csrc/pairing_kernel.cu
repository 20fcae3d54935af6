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
#ifndef ZKP_FE_SPLIT_MIN
#define ZKP_FE_SPLIT_MIN ((size_t)1 << 15)   // checks; smaller batches run their final exponentiation as one piece
#endif
#ifndef ZKP_MIN_BLOCKS_FE
#define ZKP_MIN_BLOCKS_FE 3   // same for the final-exponentiation kernel (measured: 93.9 ms vs 96.4 at 2, 2^18)
#endif

using namespace zkp;

// ------------------------------------------------------------------ final-exponentiation scratch
//
// The final exponentiation is 13 launches: k_pairing (Miller loop and/or load, then fe_prepare), then for
// each of the six stages of pairing.cuh's fe_stage a k_fe_batch_inv (the stage's one Fp inversion,
// batched across pairings) followed by k_fe_stage.  Between launches each lane parks its half of the
// live values in `lanes` and the pair's norm in `norm` -- internal Montgomery limbs, never seen by the
// caller.  Layout: slot-major, lanes[slot][2 n] Fp, so that the 32 lanes of a warp touch one contiguous
// 1536-byte run per slot (coalesced 128-bit loads/stores).  Slots (Fp per lane):
//   0..5 m   6..11 y   12..23 the three snapshots   24 p1   25 p2   26 t
// stage 0's inputs reuse slots it overwrites itself afterwards (a lane only ever touches its own column):
// f in 6..11, the FeState in 12..15.  A stage moves only what it reads / changes: ~16 KB per pairing over
// the whole pipeline, 0.4 % of the step at HBM speed.
#define ZKP_FE_LANE_FP 27
#define ZKP_SLOT_M 0
#define ZKP_SLOT_Y 6
#define ZKP_SLOT_CEXP 12
#define ZKP_SLOT_F ZKP_SLOT_Y
#define ZKP_SLOT_FES ZKP_SLOT_CEXP
struct FeScratch {
    Fp *lanes;   // [ZKP_FE_LANE_FP][2 * n]
    Fp *norm;    // [n], replaced by its inverse in place
    size_t n2;   // 2 * n
};
extern "C" size_t zkp_fe_scratch_bytes(size_t n) { return n * (2 * ZKP_FE_LANE_FP + 1) * sizeof(Fp); }

ZKP_HD void park_fp12(const FeScratch &fs, size_t lane, int slot, const Fp12 &f) {
    const Fp2 *c = &f.c0.c0;
#pragma unroll
    for (int j = 0; j < 6; j++) fs.lanes[(size_t)(slot + j) * fs.n2 + lane] = c[j].c;
}
ZKP_HD void fetch_fp12(const FeScratch &fs, size_t lane, int slot, Fp12 &f) {
    Fp2 *c = &f.c0.c0;
#pragma unroll
    for (int j = 0; j < 6; j++) c[j].c = fs.lanes[(size_t)(slot + j) * fs.n2 + lane];
}
ZKP_HD void park_cexp(const FeScratch &fs, size_t lane, const CExp &c) {
    const Fp2 *z = &c.s[0][0];
#pragma unroll
    for (int j = 0; j < 12; j++) fs.lanes[(size_t)(ZKP_SLOT_CEXP + j) * fs.n2 + lane] = z[j].c;
    fs.lanes[(size_t)(ZKP_SLOT_CEXP + 12) * fs.n2 + lane] = c.p1.c;
    fs.lanes[(size_t)(ZKP_SLOT_CEXP + 13) * fs.n2 + lane] = c.p2.c;
    fs.lanes[(size_t)(ZKP_SLOT_CEXP + 14) * fs.n2 + lane] = c.t.c;
}
ZKP_HD void fetch_cexp(const FeScratch &fs, size_t lane, CExp &c) {
    Fp2 *z = &c.s[0][0];
#pragma unroll
    for (int j = 0; j < 12; j++) z[j].c = fs.lanes[(size_t)(ZKP_SLOT_CEXP + j) * fs.n2 + lane];
    c.p1.c = fs.lanes[(size_t)(ZKP_SLOT_CEXP + 12) * fs.n2 + lane];
    c.p2.c = fs.lanes[(size_t)(ZKP_SLOT_CEXP + 13) * fs.n2 + lane];
    c.t.c = fs.lanes[(size_t)(ZKP_SLOT_CEXP + 14) * fs.n2 + lane];
}

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

// one stage of the final exponentiation (pairing.cuh fe_stage): consumes the inverse the preceding
// k_fe_batch_inv left in norm[i], leaves the next norm there; the last stage stores the result
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS_FE)
k_fe_stage(int stage, FeScratch fs, uint64_t *__restrict__ out, uint8_t *__restrict__ is_one, size_t i0, size_t n) {
    // this launch covers the checks [i0, n) of the batch
    size_t i = i0 + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1);
    bool live = i < n;
    if (!live) i = n - 1;
    size_t lane = 2 * i + lane_par();
    FeWork w;
    Fp12 f;
    FeState s;
    if (stage == 0) {
        fetch_fp12(fs, lane, ZKP_SLOT_F, f);
        s.c.c0.c = fs.lanes[(size_t)(ZKP_SLOT_FES + 0) * fs.n2 + lane];
        s.c.c1.c = fs.lanes[(size_t)(ZKP_SLOT_FES + 1) * fs.n2 + lane];
        s.c.c2.c = fs.lanes[(size_t)(ZKP_SLOT_FES + 2) * fs.n2 + lane];
        s.t.c = fs.lanes[(size_t)(ZKP_SLOT_FES + 3) * fs.n2 + lane];
    } else {
        fetch_cexp(fs, lane, w.c);
        if (stage == 1 || stage == ZKP_FE_STAGES - 1) fetch_fp12(fs, lane, ZKP_SLOT_M, w.m);
        if (stage == 2 || stage == 3 || stage == ZKP_FE_STAGES - 1) fetch_fp12(fs, lane, ZKP_SLOT_Y, w.y);
    }
    Fp ninv = fs.norm[i];
    Fp nrm = fe_stage(stage, w, &f, &s, ninv, &f);
    if (stage == ZKP_FE_STAGES - 1) {
        bool one = store_fp12(out + 72 * i, f, live);
        if (is_one && live && lane_par() == 0) is_one[i] = one ? 1 : 0;
        return;
    }
    if (live) {
        park_cexp(fs, lane, w.c);
        if (stage == 0) park_fp12(fs, lane, ZKP_SLOT_M, w.m);
        if (stage >= 1 && stage <= 3) park_fp12(fs, lane, ZKP_SLOT_Y, w.y);
        if (lane_par() == 0) fs.norm[i] = nrm;
    }
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
    fs.n2 = 2 * n;
    dim3 g((unsigned)((2 * n + ZKP_TPB - 1) / ZKP_TPB)), b(ZKP_TPB);
    switch (pair_capacity(k)) {
        case 1: k_pairing<1><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        case 2: k_pairing<2><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        case 4: k_pairing<4><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
        default: k_pairing<8><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n, fs, (const Fp *)tab, tabinf, kf); break;
    }
    *launches = 1;
    if (mode & ZKP_DO_FINAL_EXP) {
        // The batch runs as two halves on two streams: while one half is in its (latency-bound) batched
        // inversion or in the tail of a stage kernel, the other half's stage kernel keeps the SMs busy.
        size_t na = n, nb = 0;
        if (n >= ZKP_FE_SPLIT_MIN) {
            na = ((n / 2) + 63) & ~(size_t)63;
            nb = n - na;
        }
        cudaStream_t s2 = nullptr;
        cudaEvent_t fork = nullptr, join = nullptr;
        if (nb) {
            cudaError_t e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(fork, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, fork, 0);
            if (e != cudaSuccess) return e;
        }
        dim3 ga((unsigned)((2 * na + ZKP_TPB - 1) / ZKP_TPB)), gb((unsigned)((2 * nb + ZKP_TPB - 1) / ZKP_TPB));
        size_t ta = (na + ZKP_INV_RUN - 1) / ZKP_INV_RUN, tb = (nb + ZKP_INV_RUN - 1) / ZKP_INV_RUN;
        for (int stage = 0; stage < ZKP_FE_STAGES; stage++) {
            k_fe_batch_inv<<<(unsigned)((ta + 127) / 128), 128, 0, st>>>(fs.norm, na);
            k_fe_stage<<<ga, b, 0, st>>>(stage, fs, out, is_one, 0, na);
            if (nb) {
                k_fe_batch_inv<<<(unsigned)((tb + 127) / 128), 128, 0, s2>>>(fs.norm + na, nb);
                k_fe_stage<<<gb, b, 0, s2>>>(stage, fs, out, is_one, na, n);
            }
        }
        *launches = 1 + 2 * ZKP_FE_STAGES * (nb ? 2 : 1);
        if (nb) {
            cudaEventRecord(join, s2);
            cudaStreamWaitEvent(st, join, 0);
            cudaEventDestroy(fork);
            cudaEventDestroy(join);
            cudaStreamDestroy(s2);   // returns at once; the stream's resources go when its work has drained
        }
    }
    return cudaGetLastError();
}
