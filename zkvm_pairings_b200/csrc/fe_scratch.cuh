// Scratch layout of the staged final exponentiation, shared by the two translation units that touch it:
// pairing_kernel.cu (k_pairing parks f and the FeState) and fe_kernel.cu (the stages).  Included inside each
// unit's own `zkp` namespace alias, after ops.cuh.
#pragma once
#include "ops.cuh"

// helper stream + fork/join events of the final exponentiation's two-stream split; one per context stream,
// created once in zkp_ctx_create (kernels.cu) so that no call creates or destroys CUDA objects
#ifndef ZKP_FE_AUX_DEFINED
#define ZKP_FE_AUX_DEFINED
struct ZkpFeAux {
    cudaStream_t s2;
    cudaEvent_t fork, join;
};
#endif

namespace zkp {

// // The final exponentiation is 13 launches: k_pairing (Miller loop and/or load, then fe_prepare), then for
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


}  // namespace zkp
