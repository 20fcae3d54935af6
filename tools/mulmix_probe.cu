// Occupancy / phase-overlap probe (measurement tool, not part of libzkpair.so).
//
// Question: with NO instruction-cache pressure (tiny code), what fraction of the integer-multiply
// pipe does the real instruction mix reach as a function of resident warps per scheduler?  Every
// lane pair loops over  x = x*y ; four dependent Fp2 add/sub ; y = y^2 ; two add/sub -- roughly
// the add:mul ratio of the pairing.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mulmix_probe tools/mulmix_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../zkvm_pairings_b200/csrc/tower.cuh"

using namespace zkp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ADDS>
__global__ void __launch_bounds__(128) k_mix(uint32_t *sink, int iters, int delay, int sms) {
    extern __shared__ uint8_t dyn[];
    if (delay > 0 && ((blockIdx.x / sms) & 1)) {   // de-phase every second co-resident block
        long long t0 = clock64();
        while (clock64() - t0 < delay) {}
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
    if (s == 0x12345678u) sink[0] = s + dyn[0];
}

template <class F>
static void run(const char *name, F launch, int wps, double macs_per_thread, int sms) {
    size_t smem = (size_t)(200 * 1024) / wps - 1024;
    int blocks = sms * wps * 4;   // 4 waves of blocks, each block = 1 warp per scheduler
    uint32_t *sink;
    CK(cudaMalloc(&sink, 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        launch(blocks, smem, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double total = macs_per_thread * 128.0 * blocks / (best * 1e-3);
    printf("%-22s warps/smsp=%d  %.3f ms  %.2f T wide-MAC/s\n", name, wps, best, total / 1e12);
    fflush(stdout);
    cudaFree(sink);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_mix<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_mix<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_mix<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int iters = 2000;
    int ws[] = {1, 2, 3, 4, 6, 8};
    int delays[] = {0, 1500, 3000};
    for (int w : ws) {
        for (int d : delays) {
            if (w == 1 && d) continue;
            char name[64];
            snprintf(name, sizeof name, "mul+sqr      d=%d", d);
            run(name, [&](int b, size_t s, uint32_t *k) { k_mix<0><<<b, 128, s>>>(k, iters, d, sms); }, w, iters * 744.0, sms);
            snprintf(name, sizeof name, "mul+sqr+6add d=%d", d);
            run(name, [&](int b, size_t s, uint32_t *k) { k_mix<1><<<b, 128, s>>>(k, iters, d, sms); }, w, iters * 744.0, sms);
            snprintf(name, sizeof name, "mul+sqr+10ad d=%d", d);
            run(name, [&](int b, size_t s, uint32_t *k) { k_mix<2><<<b, 128, s>>>(k, iters, d, sms); }, w, iters * 744.0, sms);
        }
    }
    return 0;
}
