// Multiply-pipe token probe (measurement tool, not part of libzkpair.so).
//
// Question: the warps that share a scheduler time-slice the integer-multiply pipe when several of them are inside a
// Montgomery product at once (every IMAD.WIDE stalls its warp for 4 cycles, the scheduler switches to the next ready
// warp), so they finish their products together, enter their linear phases together and the pipe idles.  Does a FIFO
// discipline -- a token passed round a ring of the warps of one scheduler through named barriers (bar.arrive /
// bar.sync, 64 threads), one or two tokens per ring -- keep the pipe busier than the hardware's time-slicing?
// Same loop as tools/mulmix_probe.cu (x = x*y; Fp2 add/sub; y = y^2; add/sub), ONE block of 128*WPS threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ring_probe tools/ring_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

// cfg: bit0-3 tokens per ring (0 = no ring); the scheduler of a warp is taken to be warp & 3 (the first run of this probe
// measured the other guess, warp / WPS, 10-15 % slower: rings across schedulers only serialise)
#define ZKP_CONVERGED 1   // all lanes on one path: plain full-mask SHFL, as in the pairing kernels
__constant__ int ring_cfg;
__device__ __forceinline__ void ring_edges(int &in, int &out, int &pos, int &wps) {
    int w = threadIdx.x >> 5;
    wps = blockDim.x >> 7;
    int grp = w & 3;
    pos = w >> 2;
    out = grp * 4 + pos;
    in = grp * 4 + (pos == 0 ? wps - 1 : pos - 1);
}
__device__ __forceinline__ void ring_acquire() {
    if ((ring_cfg & 15) == 0) return;
    int in, out, pos, wps;
    ring_edges(in, out, pos, wps);
    asm volatile("bar.sync %0, 64;" ::"r"(in) : "memory");
}
__device__ __forceinline__ void ring_release() {
    if ((ring_cfg & 15) == 0) return;
    int in, out, pos, wps;
    ring_edges(in, out, pos, wps);
    asm volatile("bar.arrive %0, 64;" ::"r"(out) : "memory");
}
#define ZKP_PIPE_ACQUIRE() ring_acquire()
#define ZKP_PIPE_RELEASE() ring_release()

#include "../zkvm_pairings_b200/csrc/tower.cuh"

using namespace zkp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ADDS, int TPB>
__global__ void __launch_bounds__(TPB, 512 / TPB) k_mix(uint32_t *sink, int iters) {
    int tokens = ring_cfg & 15;
    if (tokens) {   // hand out the initial tokens: the predecessor of every starting warp "releases" once
        int in, out, pos, wps;
        ring_edges(in, out, pos, wps);
        bool starts_next = tokens == 1 ? pos == wps - 1 : (pos & 1) == 1;
        if (wps > 1 && starts_next) asm volatile("bar.arrive %0, 64;" ::"r"(out) : "memory");
    }
    Fp2 x, y;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        x.c.l[i] = (threadIdx.x * 977 + i * 131 + blockIdx.x) & 0x0fffffff;
        y.c.l[i] = (threadIdx.x * 31 + i * 17 + 5) & 0x0fffffff;
    }
    for (int it = 0; it < iters; it++) {
        Fp2 t = fp2_mul(x, y);
        if (ADDS >= 1) { Fp2 s = fp2_add(t, x); Fp2 d = fp2_sub(t, y); x = fp2_sub(s, d); t = fp2_add(x, t); }
        if (ADDS >= 2) { Fp2 s = fp2_add(t, y); Fp2 d = fp2_sub(t, x); t = fp2_sub(s, d); t = fp2_add(x, t); }
        x = t;
        Fp2 q = fp2_sqr(y);
        if (ADDS >= 1) { q = fp2_add(q, x); q = fp2_sub(q, y); }
        y = q;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) s ^= x.c.l[i] ^ y.c.l[i];
    if (s == 0x12345678u) sink[0] = s;
}

template <int ADDS, int TPB>
static void run(int cfg, int sms, int iters, int wps) {
    CK(cudaMemcpyToSymbol(ring_cfg, &cfg, sizeof cfg));
    // occupancy by shared memory: `wps` warps per scheduler = one block of 128*wps threads, or wps blocks of 128
    int per_sm = TPB == 128 ? wps : 1;
    size_t smem = (size_t)(200 * 1024) / per_sm - 1024;
    CK(cudaFuncSetAttribute(k_mix<ADDS, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int blocks = sms * per_sm * 4;
    uint32_t *sink;
    CK(cudaMalloc(&sink, 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_mix<ADDS, TPB><<<blocks, TPB, smem>>>(sink, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double total = iters * 744.0 * TPB * blocks / (best * 1e-3);
    printf("adds=%-2d warps/smsp=%d block=%d tokens=%d  %.3f ms  %.2f T wide-MAC/s\n", ADDS == 0 ? 0 : ADDS == 1 ? 6 : 10, wps, TPB,
           cfg & 15, best, total / 1e12);
    fflush(stdout);
    cudaFree(sink);
}

template <int TPB>
static void sweep(int sms, int iters, int wps, int cfg) {
    run<0, TPB>(cfg, sms, iters, wps);
    run<1, TPB>(cfg, sms, iters, wps);
    run<2, TPB>(cfg, sms, iters, wps);
}

// usage: ring_probe WPS BLOCK TOKENS   (one configuration per process, so that a wedged ring cannot take the others along)
int main(int argc, char **argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 2000;
    int wps = argc > 1 ? atoi(argv[1]) : 4, tpb = argc > 2 ? atoi(argv[2]) : 128, tok = argc > 3 ? atoi(argv[3]) : 0;
    if (tpb == 128) sweep<128>(sms, iters, wps, 0);
    else if (tpb == 256 && wps == 2) sweep<256>(sms, iters, wps, tok);
    else if (tpb == 384 && wps == 3) sweep<384>(sms, iters, wps, tok);
    else if (tpb == 512 && wps == 4) sweep<512>(sms, iters, wps, tok);
    else { fprintf(stderr, "unsupported geometry\n"); return 2; }
    return 0;
}
