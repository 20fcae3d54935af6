// Is there multiply throughput on B200 beside the IMAD.WIDE pipe?  Measures, on all SMs:
//   kind 0: IMAD.WIDE.U32 chains alone          kind 1: DFMA chains alone
//   kind 2: both in the same warp (8 + 8 chains) kind 3: half of the warps each
// and prints giga-instructions per second per kind.  A 52 x 52 -> 104-bit product costs 2 DFMA (+1 DADD) in the
// double-precision limb technique against 4 IMAD.WIDE for the same bits (2704 vs 4 x 1024 bit^2), so DFMA throughput at
// or above the IMAD.WIDE rate on a SEPARATE pipe would be headroom the 32-bit limb design leaves unused.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dfma_probe tools/dfma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t *sink, int iters) {
    uint32_t x = threadIdx.x * 2654435761u + 12345u, y = blockIdx.x * 40503u + 977u + sink[1];
    uint64_t acc[8];
    double d[8], m = 1.0 + (double)(x & 1023) * 1e-9, c = (double)(y & 255) * 1e-12;
    uint32_t ys[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = x + i; ys[i] = y * (2 * i + 3); d[i] = 1.0 + i * 1e-3; }
    const bool do_i = KIND == 0 || KIND == 2 || (KIND == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_d = KIND == 1 || KIND == 2 || (KIND == 3 && ((threadIdx.x >> 5) & 1) == 1);
    uint32_t mm = x;
    for (int it = 0; it < iters; it++) {
        asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(mm));
        if (do_i) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(mm), "r"(ys[i]));
        }
        if (do_d) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
        }
    }
    uint64_t s = 0;
    double t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s ^= acc[i]; t += d[i]; }
    if (s == 0x123456789abcdefull || t == 1.2345e300) sink[0] = (uint32_t)s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink;
    cudaMalloc(&sink, 16);
    cudaMemset(sink, 0, 16);
    const int iters = 1 << 13, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[4] = {"IMAD.WIDE alone", "DFMA alone", "IMAD.WIDE + DFMA, same warp", "IMAD.WIDE warps + DFMA warps"};
    for (int kind = 0; kind < 4; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (kind == 0) k<0><<<blocks, threads>>>(sink, iters);
            else if (kind == 1) k<1><<<blocks, threads>>>(sink, iters);
            else if (kind == 2) k<2><<<blocks, threads>>>(sink, iters);
            else k<3><<<blocks, threads>>>(sink, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        double n_i = (kind == 0 || kind == 2) ? 8.0 : kind == 3 ? 4.0 : 0.0, n_d = (kind == 1 || kind == 2) ? 8.0 : kind == 3 ? 4.0 : 0.0;
        double tot = (double)iters * blocks * threads;
        printf("%-32s %8.3f ms   IMAD.WIDE %7.2f G/s   DFMA %7.2f G/s   (per SM per clock at 1.965 GHz: %.1f / %.1f lanes)\n", names[kind], best,
               n_i * tot / best / 1e6, n_d * tot / best / 1e6, n_i * tot / best / 1e6 / sms / 1.965, n_d * tot / best / 1e6 / sms / 1.965);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
