// Can ptxas hide independent Fp2 additions in the issue shadow of the IMAD.WIDE stream of an Fp2
// product when both live in ONE out-of-line function?  (measurement tool, not part of libzkpair.so)
//   A: x = mul(x, y); s = add(p, q); d = sub(p, q)            -- three calls/inlines, serial phases
//   B: (x, s, d) = mul_add_sub(x, y, p, q)                     -- one fused out-of-line body
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fused_probe tools/fused_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define ZKP_CONVERGED 1
#include "../zkvm_pairings_b200/csrc/tower.cuh"
using namespace zkp;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Tri { Fp2 m, s, d; };
__device__ __noinline__ Tri fused(Fp2 a, Fp2 b, Fp2 p, Fp2 q) {
    Tri r;
    Fp t = fp_xchg(fp_select(lane_par() != 0, fp_neg(a.c), a.c));
    Fp b0 = fp_bcast<0>(b.c), b1 = fp_bcast<1>(b.c);
    r.m.c = mont_mul2(a.c, b0, t, b1);
    r.s.c = fp_add(p.c, q.c);
    r.d.c = fp_sub(p.c, q.c);
    return r;
}
__device__ __noinline__ Fp2 add_ool(Fp2 a, Fp2 b) { return fp2_add(a, b); }
__device__ __noinline__ Fp2 sub_ool(Fp2 a, Fp2 b) { return fp2_sub(a, b); }

template <int MODE>
__global__ void __launch_bounds__(128, 2) k(uint32_t *sink, int iters) {
    Fp2 x, y, p, q;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) {
        x.c.l[i] = (threadIdx.x * 977 + i * 131 + blockIdx.x) & 0x0fffffff;
        y.c.l[i] = (threadIdx.x * 31 + i * 17 + 5) & 0x0fffffff;
        p.c.l[i] = (threadIdx.x * 3 + i * 7 + 1) & 0x0fffffff;
        q.c.l[i] = (threadIdx.x * 5 + i * 11 + 2) & 0x0fffffff;
    }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {           // product only
            x = fp2_mul(x, y);
        } else if (MODE == 1) {    // product, then inlined add + sub (what the tower does today)
            x = fp2_mul(x, y);
            Fp2 s = fp2_add(p, q), d = fp2_sub(p, q);
            p = s; q = d;
        } else {                   // fused
            Tri r = fused(x, y, p, q);
            x = r.m; p = r.s; q = r.d;
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ZKP_NL; i++) s ^= x.c.l[i] ^ p.c.l[i] ^ q.c.l[i] ^ y.c.l[i];
    if (s == 0x12345678u) sink[0] = s;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t *sink;
    CK(cudaMalloc(&sink, 64));
    const int iters = 4000, blocks = sms * 2 * 4;
    const char *names[3] = {"mul only", "mul ; add ; sub", "fused mul|add|sub"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            if (mode == 0) k<0><<<blocks, 128>>>(sink, iters);
            else if (mode == 1) k<1><<<blocks, 128>>>(sink, iters);
            else k<2><<<blocks, 128>>>(sink, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("%-20s %.3f ms   %.2f T wide-MAC/s\n", names[mode], best, 444.0 * iters * 128.0 * blocks / (best * 1e-3) / 1e12);
    }
    return 0;
}
