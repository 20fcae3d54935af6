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
#define ZKP_MIN_BLOCKS 2      // resident blocks per SM the register allocator must allow
#endif

using namespace zkp;

// mode: bit0 Miller loop, bit1 final exponentiation.  One lane pair per check of k (<= K) pairs.
template <int K>
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_pairing(int mode, const uint64_t *__restrict__ g1, const uint8_t *__restrict__ g1inf,
          const uint64_t *__restrict__ g2, const uint8_t *__restrict__ g2inf, int k,
          const uint64_t *__restrict__ in12, uint64_t *__restrict__ out, uint8_t *__restrict__ is_one,
          uint32_t *err, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    bool live = i < n;
    if (!live) i = n - 1;   // stay converged: redo the last element, store nothing
    size_t e = i * (size_t)k;
    uint8_t s = pairing_one<K>(mode, g1 ? g1 + 12 * e : nullptr, g1inf ? g1inf + e : nullptr,
                               g2 ? g2 + 24 * e : nullptr, g2inf ? g2inf + e : nullptr, k,
                               in12 ? in12 + 72 * i : nullptr, out + 72 * i, is_one ? is_one + i : nullptr, live);
    if (s && err && live && lane_par() == 0) atomicOr(err, 1u);
}

static int pair_capacity(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

cudaError_t zkp_launch_k_pairing(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                 size_t n, int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one, uint32_t *err,
                                 cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    dim3 g((unsigned)((2 * n + ZKP_TPB - 1) / ZKP_TPB)), b(ZKP_TPB);
    switch (pair_capacity(k)) {
        case 1: k_pairing<1><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n); break;
        case 2: k_pairing<2><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n); break;
        case 4: k_pairing<4><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n); break;
        default: k_pairing<8><<<g, b, 0, st>>>(mode, g1, g1inf, g2, g2inf, k, in12, out, is_one, err, n); break;
    }
    return cudaGetLastError();
}
