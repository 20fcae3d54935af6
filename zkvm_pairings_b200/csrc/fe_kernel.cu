// Final-exponentiation kernels (k_fe_batch_inv, k_fe_stage) in their own translation unit: like
// pairing_kernel.cu it is compiled warp-converged, but with its own copy of the device functions so that it
// can carry its own density of block-wide rendezvous points (ZKP_CODE_SYNC, fp.cuh).  The stage kernels run
// 4 blocks of 4 warps per SM -- one warp of a block per scheduler -- through ~60 KB of straight-line code
// against a 32 KB L1.5 instruction cache: keeping the four warps of a block within one Fp6-level body of
// each other lets them share the fetched lines -- and with that, four blocks per SM at 128 registers beat three
// at 168.  Final exponentiation only, 2^20 (profiles/r1l_code_sync_variants.txt, profiles/r1o_occupancy_variants.txt):
//   3 blocks/SM: no rendezvous 298.0 ms, per compressed squaring 283.3, per Fp6-level body 275.5, per Fp2 op 278.4
//   2 blocks/SM 310.9;  4 blocks/SM: Fp6-level 266.0 (kept), Fp2-level 266.4;  5 blocks 271.6;  6 blocks 279.5
#include <cuda_runtime.h>

#include <atomic>

#define ZKP_CONVERGED 1
#define zkp zkp_fe
#ifndef ZKP_FE_SYNC
#define ZKP_FE_SYNC 4
#endif
#define ZKP_LOOP_SYNC ZKP_FE_SYNC
#ifndef ZKP_FE_SMEM
#define ZKP_FE_SMEM 0          // 1: the running state of the compressed squaring chains in shared memory (pairing.cuh cexp_begin);
                               //    measured neutral (profiles/r2h_final_exp_variants.txt), off
#endif
#if ZKP_FE_SMEM
#define ZKP_CEXP_Z_SMEM 1
#endif
#include "../../include/zkpair.h"
#include "fe_scratch.cuh"

#ifndef ZKP_TPB_FE
#define ZKP_TPB_FE 128        // threads per block of the stage kernels
#endif
#undef ZKP_TPB
#define ZKP_TPB ZKP_TPB_FE
#ifndef ZKP_FE_SPLIT_MIN
#define ZKP_FE_SPLIT_MIN ((size_t)1 << 15)   // checks; smaller batches run their final exponentiation as one piece
#endif
#ifndef ZKP_MIN_BLOCKS_FE
#define ZKP_MIN_BLOCKS_FE 4   // resident blocks per SM
#endif

using namespace zkp;

#define ZKP_FE_SMEM_BYTES (ZKP_FE_SMEM ? (size_t)208 * ZKP_TPB : (size_t)0)

extern "C" void zkp_fe_geometry(int *tpb, int *blocks, int *sync) { *tpb = ZKP_TPB; *blocks = ZKP_MIN_BLOCKS_FE; *sync = ZKP_FE_SYNC; }
extern "C" size_t zkp_fe_scratch_bytes(size_t n) { return n * (2 * ZKP_FE_LANE_FP + 1) * sizeof(Fp); }

// norm[i] <- 1 / norm[i]: every thread inverts a run of ZKP_INV_RUN norms with Montgomery's trick -- 3 (run - 1) products
// and ONE inversion, a binary extended GCD on the ALU pipe (tower.cuh fp_inv): 0.245 ms per launch at 2^16 pairings against
// 0.69 ms for the Fermat ladder of round 1, and no work for the multiply pipe
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

// The six (batched inversion, stage) launch pairs over the state k_pairing parked in `scratch`.
// `aux` = the helper stream and the fork/join events of the two-stream split (owned by the context, created
// once in zkp_ctx_create; NULL members = run as one piece).
// where the batch splits into its two halves (n = no split)
size_t zkp_fe_split_point(size_t n, const ZkpFeAux *aux) {
    if (n >= ZKP_FE_SPLIT_MIN && aux && aux->s2 && aux->fork && aux->join) return ((n / 2) + 63) & ~(size_t)63;
    return n;
}

// `forked` = the caller has already made aux->s2 wait for `st` (the Miller kernels of the two halves ran on the two streams)
cudaError_t zkp_launch_fe_stages(void *scratch, size_t n, uint64_t *out, uint8_t *is_one, cudaStream_t st, const ZkpFeAux *aux,
                                 int *launches, int forked) {
    FeScratch fs;
    fs.lanes = (Fp *)scratch;
    fs.norm = fs.lanes + 2 * n * ZKP_FE_LANE_FP;
    fs.n2 = 2 * n;
    dim3 b(ZKP_TPB);
#ifdef ZKP_FE_CARVEOUT
    {   // experiment: an explicit L1 / shared-memory split for the stage kernels (percent of shared memory; per device)
        static std::atomic<unsigned long long> done{0};
        int cur = 0;
        cudaGetDevice(&cur);
        if (!((done.load() >> (cur & 63)) & 1ull)) {
            cudaFuncSetAttribute(k_fe_stage, cudaFuncAttributePreferredSharedMemoryCarveout, ZKP_FE_CARVEOUT);
            done.fetch_or(1ull << (cur & 63));
        }
    }
#endif
    // The batch runs as two halves on two streams: while one half is in its (latency-bound) batched
    // inversion or in the tail of a stage kernel, the other half's stage kernel keeps the SMs busy.
    size_t na = zkp_fe_split_point(n, aux), nb = n - na;
    *launches = 0;
    cudaStream_t s2 = nb ? aux->s2 : nullptr;
    if (nb && !forked) {
        cudaError_t e = cudaEventRecord(aux->fork, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, aux->fork, 0);
        if (e != cudaSuccess) return e;
    }
    dim3 ga((unsigned)((2 * na + ZKP_TPB - 1) / ZKP_TPB)), gb((unsigned)((2 * nb + ZKP_TPB - 1) / ZKP_TPB));
    size_t ta = (na + ZKP_INV_RUN - 1) / ZKP_INV_RUN, tb = (nb + ZKP_INV_RUN - 1) / ZKP_INV_RUN;
    for (int stage = 0; stage < ZKP_FE_STAGES; stage++) {
        k_fe_batch_inv<<<(unsigned)((ta + 127) / 128), 128, 0, st>>>(fs.norm, na);
        k_fe_stage<<<ga, b, ZKP_FE_SMEM_BYTES, st>>>(stage, fs, out, is_one, 0, na);
        if (nb) {
            k_fe_batch_inv<<<(unsigned)((tb + 127) / 128), 128, 0, s2>>>(fs.norm + na, nb);
            k_fe_stage<<<gb, b, ZKP_FE_SMEM_BYTES, s2>>>(stage, fs, out, is_one, na, n);
        }
    }
    *launches = 2 * ZKP_FE_STAGES * (nb ? 2 : 1);
    cudaError_t rc = cudaGetLastError();
    if (nb) {   // always rejoin, also after a failed launch: `st` must not run ahead of the helper stream
        cudaError_t e = cudaEventRecord(aux->join, s2);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, aux->join, 0);
        if (rc == cudaSuccess) rc = e;
    }
    return rc;
}
