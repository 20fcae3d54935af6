// Integer-multiply issue probe for sm_100a (measurement tool, not part of libzkpair.so).
//
// Question it answers: how many IMAD.WIDE per clock does ONE SM sub-partition sustain as a function
// of (a) the number of resident warps per scheduler, (b) signed vs unsigned wide multiply,
// (c) the number of independent accumulator chains per thread, and (d) for the real Montgomery
// product of csrc/fp.cuh.  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_probe imad_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../zkvm_pairings_b200/csrc/fp.cuh"

using namespace zkp;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

template <int KIND, int NCH>
__global__ void __launch_bounds__(128) k_chain(uint64_t *sink, int iters, long long *cyc) {
    extern __shared__ uint8_t dyn[];
    uint32_t x = threadIdx.x * 2654435761u + 12345u, y = blockIdx.x * 40503u + 977u;
    uint64_t acc[NCH];
    uint32_t ys[NCH];   // a distinct multiplier per chain: ptxas cannot factor a*y + b*y
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        acc[c] = x + c;
        ys[c] = y * (2 * c + 3) + sink[8 + c];
    }
    long long t0 = clock64();
    uint32_t m = x;
    for (int it = 0; it < iters; it++) {
        m = m * 1664525u + 1013904223u;   // one multiplicand per iteration (nothing is loop invariant)
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(m), "r"(ys[c]));
            else if (KIND == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(m), "r"(ys[c]));
            else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(((uint32_t *)acc)[2 * c]) : "r"(m), "r"(ys[c]));
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) s ^= acc[c];
    if (s == 0x123456789abcdefull) sink[0] = s + dyn[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the real Montgomery product, back to back on dependent data (x = x*y), as one thread runs it
template <int SQR>
__global__ void __launch_bounds__(128) k_mont(uint64_t *sink, int iters, long long *cyc) {
    extern __shared__ uint8_t dyn[];
    Fp a, b;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        a.l[i] = (threadIdx.x * 977 + i * 131 + blockIdx.x) & ZKP_M28;
        b.l[i] = (threadIdx.x * 31 + i * 17 + 5) & ZKP_M28;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (SQR) a = mont_mul(a, a);
        else a = mont_mul(a, b);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) s ^= a.l[i];
    if (s == 0x12345678u) sink[0] = s + dyn[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F>
static void run(const char *name, F launch, int warps_per_smsp, double macs_per_thread, int sms) {
    // one block = 128 threads = 1 warp per scheduler; limit blocks/SM through dynamic shared memory
    size_t smem = (size_t)(200 * 1024) / warps_per_smsp - 1024;
    if (smem > 200 * 1024) smem = 200 * 1024;
    int blocks = sms * warps_per_smsp;
    long long *d_cyc, *h_cyc = (long long *)malloc(sizeof(long long) * blocks);
    uint64_t *sink;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * blocks));
    CK(cudaMalloc(&sink, 1024));
    CK(cudaMemset(sink, 1, 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        launch(blocks, smem, sink, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h_cyc[i];
    avg /= blocks;
    // per scheduler: warps_per_smsp warps each issuing macs_per_thread wide MACs in `avg` cycles
    double per_clk_smsp = macs_per_thread * warps_per_smsp / avg;
    double total = macs_per_thread * 128.0 * blocks / (best * 1e-3);
    printf("%-28s warps/smsp=%d  cyc=%.0f  warp-MAC/clk/SMSP=%.3f  (pipe peak 0.5)  %.2f T MAC/s\n", name, warps_per_smsp, avg,
           per_clk_smsp, total / 1e12);
    fflush(stdout);
    cudaFree(d_cyc);
    cudaFree(sink);
    free(h_cyc);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_chain<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<0, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<1, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_chain<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_mont<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_mont<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int iters = 4096;
    int ws[] = {1, 2, 3, 4, 6, 8, 16};
    for (int w : ws) {
        run("wide.u32 x8 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<0, 8><<<b, 128, s>>>(k, iters, c); }, w, iters * 8.0, sms);
        run("wide.s32 x8 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<1, 8><<<b, 128, s>>>(k, iters, c); }, w, iters * 8.0, sms);
        run("lo.u32   x8 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<2, 8><<<b, 128, s>>>(k, iters, c); }, w, iters * 8.0, sms);
        run("wide.u32 x14 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<0, 14><<<b, 128, s>>>(k, iters, c); }, w, iters * 14.0, sms);
        run("wide.s32 x14 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<1, 14><<<b, 128, s>>>(k, iters, c); }, w, iters * 14.0, sms);
        run("wide.u32 x2 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<0, 2><<<b, 128, s>>>(k, iters, c); }, w, iters * 2.0, sms);
        run("wide.u32 x4 chains", [&](int b, size_t s, uint64_t *k, long long *c) { k_chain<0, 4><<<b, 128, s>>>(k, iters, c); }, w, iters * 4.0, sms);
        run("mont_mul (406 MAC)", [&](int b, size_t s, uint64_t *k, long long *c) { k_mont<0><<<b, 128, s>>>(k, 512, c); }, w, 512 * 406.0, sms);
        run("mont_sqr (406 MAC)", [&](int b, size_t s, uint64_t *k, long long *c) { k_mont<1><<<b, 128, s>>>(k, 512, c); }, w, 512 * 406.0, sms);
    }
    return 0;
}
