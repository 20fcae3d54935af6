// libzkpair.so: CUDA kernels (sm_100a) + the C ABI declared in include/zkpair.h.
//
// This translation unit holds the C ABI, the host-side pipeline and the kernels whose lane pairs may
// diverge from each other (tower ops, products, point generation: pair-masked shuffles).  The
// pairing kernel itself is pairing_kernel.cu (warp-converged, full-mask shuffles).
//
// A LANE PAIR (two adjacent threads) owns one pairing check (k pairs -> shared-accumulator Miller
// loop -> final exponentiation): every Fp2 of the tower is split over the pair (tower.cuh), so a
// warp runs 16 independent checks in lock-step (the control flow is data independent) and the
// integer-multiply pipe is kept busy by the carry-chained 32-bit-limb Montgomery products of
// fp.cuh.  Independent checks shard in contiguous slices over the context's devices;
// the only cross-device traffic is the 576-byte Fp12 partial of zkp_multi_miller_product.
//
// There is deliberately no CPU implementation in this library: with no CUDA device every entry
// point fails with ZKP_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zkpair.h"
#include "fe_scratch.cuh"   // ops.cuh + the context-owned helper objects of the final exponentiation

#ifndef ZKP_TPB
#define ZKP_TPB 128           // threads per block of the pairing kernels
#endif
#ifndef ZKP_MIN_BLOCKS
#define ZKP_MIN_BLOCKS 2      // resident blocks per SM the register allocator must allow
#endif
#ifndef ZKP_PRODUCT_GROUP
#define ZKP_PRODUCT_GROUP 4   // pairs per shared-accumulator Miller loop in zkp_multi_miller_product (k_pairing<4>)
#endif
#ifndef ZKP_CHUNK
#define ZKP_CHUNK (1u << 18)  // checks per host<->device pipeline chunk (e2e at 2^20: 1.735 / 1.776 / 1.794 M/s for 2^16 / 2^17 / 2^18)
#endif

using namespace zkp;

// ================================================================== kernels

__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_tower_op(int op, const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint64_t *__restrict__ out,
           uint8_t *__restrict__ status, uint32_t *err, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;   // element = lane pair
    if (i >= n) return;
    int na, nb, nr;
    tower_op_shape(op, na, nb, nr);
    uint8_t s = tower_op_one(op, a + (size_t)6 * na * i, nb ? b + (size_t)6 * nb * i : nullptr, out + (size_t)6 * nr * i);
    if (lane_par() == 0) {
        if (status) status[i] = s;
        if ((s & 1) && err) atomicOr(err, 1u);
    }
}

// out[t] = in[t] * in[t+m] * in[t+2m] * ...   (t < m <= n), canonical limbs in and out
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_fp12_product(const uint64_t *__restrict__ in, size_t n, uint64_t *__restrict__ out, size_t m, uint32_t *err) {
    size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    if (t >= m) return;
    bool bad = false;
    Fp12 acc, x;
    load_fp12(acc, in + 72 * t, bad);
    for (size_t j = t + m; j < n; j += m) {
        load_fp12(x, in + 72 * j, bad);
        fp12_mul(acc, acc, x);
    }
    store_fp12(out + 72 * t, acc);
    if (lane_or(bad) && err && lane_par() == 0) atomicOr(err, 1u);
}

__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_gen_points(uint64_t seed, uint64_t first, size_t n, uint64_t *__restrict__ g1, uint8_t *__restrict__ g1inf,
             uint64_t *__restrict__ g2, uint8_t *__restrict__ g2inf) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    if (i >= n) return;
    uint64_t a = splitmix64_at(seed, 2 * (first + i));
    uint64_t b = splitmix64_at(seed, 2 * (first + i) + 1);
    if (a == 0) a = 1;
    if (b == 0) b = 1;
    gen_g1_one(a, g1 + 12 * i, g1inf + i);
    gen_g2_one(b, g2 + 24 * i, g2inf + i);
}

// G2Prepared-style line tables: out = [point][68][3][lane parity] Montgomery Fp
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_g2_prepare(const uint64_t *__restrict__ g2, Fp *__restrict__ out, uint32_t *err, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    if (i >= n) return;
    bool bad = false;
    G2A q;
    load_g2(q, g2 + 24 * i, bad);
    g2_prepare(q, out + i * (ZKP_LINE_STEPS * 3 * 2) + lane_par());
    if (lane_or(bad) && err && lane_par() == 0) atomicOr(err, 1u);
}

// Group-level batch ops: subgroup/on-curve validation and scalar multiplication.
// gop = GroupOp; points are 12 (G1) or 24 (G2) u64 each; flag = status (checks), is_infinity (mul) or bit0 identity | bit1 undefined (add).
__global__ void __launch_bounds__(ZKP_TPB, ZKP_MIN_BLOCKS)
k_group_op(int gop, const uint64_t *__restrict__ pts, const uint8_t *__restrict__ inf, const uint64_t *__restrict__ scalars,
           uint64_t *__restrict__ out_pts, uint8_t *__restrict__ flag, uint32_t *err, size_t n, const uint8_t *__restrict__ inf2) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    if (i >= n) return;
    const int w = (gop & 1) ? 24 : 12;
    bool bad = false;
    uint8_t is_inf = inf ? inf[i] : 0, f = 0;
    switch (gop) {
        case GOP_G1_CHECK: f = g1_check_one(pts + (size_t)w * i, is_inf, bad); break;
        case GOP_G2_CHECK: f = g2_check_one(pts + (size_t)w * i, is_inf, bad); break;
        case GOP_G1_MUL: g1_mul_one(pts + (size_t)w * i, is_inf, scalars + 4 * i, out_pts + (size_t)w * i, flag + i, bad); break;
        case GOP_G2_MUL: g2_mul_one(pts + (size_t)w * i, is_inf, scalars + 4 * i, out_pts + (size_t)w * i, flag + i, bad); break;
        // additions: the second operand travels in `scalars` (w u64 per element), its flags in inf2
        case GOP_G1_ADD: g1_add_one(pts + (size_t)w * i, is_inf, scalars + (size_t)w * i, inf2 ? inf2[i] : 0, out_pts + (size_t)w * i, flag + i, bad); break;
        default: g2_add_one(pts + (size_t)w * i, is_inf, scalars + (size_t)w * i, inf2 ? inf2[i] : 0, out_pts + (size_t)w * i, flag + i, bad); break;
    }
    bad = lane_or(bad);
    if (lane_par() == 0) {
        if (gop <= GOP_G2_CHECK) flag[i] = f;
        if (bad && err) atomicOr(err, 1u);
    }
}

// Byte (de)serialisation at the boundary: 48 big-endian bytes <-> six little-endian u64
// limbs, Fp::from_bytes / Fp::to_bytes (src/fp.rs:165-207).  Pure byte shuffling, HBM bound: one
// thread moves one element with three 128-bit loads and three 128-bit stores (a warp covers 1536
// contiguous bytes on both sides).  dir 0: bytes -> limbs (+ canonical flag), dir 1: limbs -> bytes.
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    uint32_t lo = __byte_perm((uint32_t)x, 0, 0x0123), hi = __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
    return ((uint64_t)lo << 32) | hi;
}
__global__ void __launch_bounds__(256) k_fp_bytes(int dir, const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                  uint8_t *__restrict__ ok, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
    // reversing all 48 bytes: the three 16-byte chunks swap ends, each chunk is byte-reversed
    uint64_t w[6] = {((uint64_t)a.y << 32) | a.x, ((uint64_t)a.w << 32) | a.z, ((uint64_t)b.y << 32) | b.x,
                     ((uint64_t)b.w << 32) | b.z, ((uint64_t)c.y << 32) | c.x, ((uint64_t)c.w << 32) | c.z};
    uint64_t r[6];
#pragma unroll
    for (int k = 0; k < 6; k++) r[k] = bswap64(w[5 - k]);
    out[3 * i] = make_uint4((uint32_t)r[0], (uint32_t)(r[0] >> 32), (uint32_t)r[1], (uint32_t)(r[1] >> 32));
    out[3 * i + 1] = make_uint4((uint32_t)r[2], (uint32_t)(r[2] >> 32), (uint32_t)r[3], (uint32_t)(r[3] >> 32));
    out[3 * i + 2] = make_uint4((uint32_t)r[4], (uint32_t)(r[4] >> 32), (uint32_t)r[5], (uint32_t)(r[5] >> 32));
    if (dir == 0 && ok) {   // canonical: value < p (src/fp.rs:176-190)
        uint64_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            uint64_t pk = ((uint64_t)ZKP_P[2 * k + 1] << 32) | ZKP_P[2 * k];
            uint64_t d = r[k] - pk;
            borrow = (r[k] < pk) | (d < borrow);
        }
        ok[i] = (uint8_t)borrow;   // 1 = canonical (Ok), 0 = Err(())
    }
}

// Integer-multiply roofline probe.  Every chain gets its own multiplier and the multiplicand changes
// every iteration, so nothing is loop invariant (ptxas would otherwise hoist the product and leave
// 64-bit adds).  KIND 0: 8 independent IMAD.WIDE.U32 accumulate chains per thread; KIND 1: 8
// independent 32-bit IMAD chains; KIND 2: the carry-chained wide MACs exactly as the Montgomery
// rows of fp.cuh issue them (row_mad: two 6-MAC carry chains per step).
template <int KIND>
__global__ void __launch_bounds__(512) k_imad_peak(uint32_t *sink, int iters) {
    uint32_t x = threadIdx.x * 2654435761u + 12345u, y = blockIdx.x * 40503u + 977u;
    uint32_t m = x;
    if (KIND == 0) {
        uint64_t acc[8];
        uint32_t ys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) { acc[c] = x + c; ys[c] = y * (2 * c + 3) + sink[1]; }
        for (int it = 0; it < iters; it++) {
            asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(m));   // rotate + add (SHF, IADD3): only the counted MACs touch the multiply pipe
            m += 0x9E3779B9u;
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(m), "r"(ys[c]));
        }
        uint64_t s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) s ^= acc[c];
        if (s == 0x123456789abcdefull) sink[0] = (uint32_t)s;
    } else if (KIND == 1) {
        uint32_t acc[8], ys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) { acc[c] = x + c; ys[c] = y * (2 * c + 3) + sink[1]; }
        for (int it = 0; it < iters; it++) {
            asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(m));   // rotate + add (SHF, IADD3): only the counted MACs touch the multiply pipe
            m += 0x9E3779B9u;
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(m), "r"(ys[c]));
        }
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) s ^= acc[c];
        if (s == 0x12345678u) sink[0] = s;
    } else {
        uint32_t ev[12], od[12], a[12];
#pragma unroll
        for (int c = 0; c < 12; c++) { ev[c] = x + c; od[c] = y + c; a[c] = x * (c + 3) + y + sink[1]; }
        for (int it = 0; it < iters; it++) {
            asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(m));   // rotate + add (SHF, IADD3): only the counted MACs touch the multiply pipe
            m += 0x9E3779B9u;
            row_mad(ev, od, a, m);
        }
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < 12; c++) s ^= ev[c] ^ od[c];
        if (s == 0x12345678u) sink[0] = s;
    }
}

// ================================================================== host side

static thread_local std::string g_err;
static int32_t fail(int32_t code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ZKP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// pinned host staging for user buffers that are pageable (a plain Vec / numpy array): an async copy to or
// from pageable memory blocks the issuing thread until it is done, which would serialise the two buffer
// sets of the pipeline; staged through these, the copies of one set overlap the kernels of the other
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};
struct PendingOut {   // staged device->host copy still to be moved into the caller's buffer
    void *user;
    const void *stage;
    size_t bytes;
};

enum { B_G1 = 0, B_G1INF, B_G2, B_G2INF, B_IN, B_OUT, B_FLAG, B_TAB, B_NBUF, B_TABINF = B_NBUF, B_NPIN };
#ifndef ZKP_STAGE_MIN
#define ZKP_STAGE_MIN ((size_t)256 << 10)   // smaller transfers go straight from / to the caller's memory
#endif

struct DevState {
    int id = 0;
    int sms = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    // helper stream + fork/join events of the final exponentiation's two-stream split: one set per context
    // stream and a third for launches on caller streams (created once here, never in a call)
    ZkpFeAux fe_aux[3] = {};
    uint32_t *d_err = nullptr;
    cudaMemPool_t pool = nullptr;   // stream-ordered scratch of the final exponentiation (kept, never trimmed)
    DevBuf buf[2][B_NBUF];     // double-buffered pipeline scratch
    PinBuf pin_in[2][B_NPIN], pin_out[2][B_NPIN];   // pinned staging per buffer set (only for pageable user buffers)
    std::vector<PendingOut> pend[2];
    DevBuf scratch, partial;   // product reduction
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers;
    double timed_ms = 0;
    uint64_t timed_launches = 0;
};

struct zkp_ctx {
    std::vector<DevState> devs;
    std::mutex mu;
    std::atomic<uint64_t> launches{0};
    bool timing = false;
};

// Entry points select the device they work on; the caller's current device is restored on every exit path
// (a torch process keeps allocating on ITS device after a call into a multi-device context).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// two threads (one lane pair) per element
static inline unsigned grid_for(size_t n) { return (unsigned)((2 * n + ZKP_TPB - 1) / ZKP_TPB); }

// the pairing kernels live in pairing_kernel.cu (compiled with warp-converged shuffles)
extern "C" size_t zkp_fe_scratch_bytes(size_t n);
cudaError_t zkp_launch_k_pairing(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                 size_t n, int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one, uint32_t *err,
                                 void *scratch, const void *tab, const uint8_t *tabinf, int kf, cudaStream_t st, const ZkpFeAux *aux,
                                 int *launches);

static cudaError_t launch_pairing(zkp_ctx *ctx, DevState &d, int mode, const uint64_t *g1, const uint8_t *g1inf,
                                  const uint64_t *g2, const uint8_t *g2inf, size_t n, int k, const uint64_t *in12,
                                  uint64_t *out, uint8_t *is_one, uint32_t *err, cudaStream_t st,
                                  const void *tab = nullptr, const uint8_t *tabinf = nullptr, int kf = 0) {
    if (n == 0) return cudaSuccess;
    // the final exponentiation parks its state between launches in stream-ordered scratch memory
    void *scratch = nullptr;
    if (mode & 2) {
        cudaError_t e = cudaMallocFromPoolAsync(&scratch, zkp_fe_scratch_bytes(n), d.pool, st);
        if (e != cudaSuccess) return e;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->timing) {   // (created after the only early return, so that no event is left behind)
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    int nl = 0;
    const ZkpFeAux *aux = &d.fe_aux[st == d.stream[0] ? 0 : st == d.stream[1] ? 1 : 2];
    cudaError_t rc = zkp_launch_k_pairing(mode, g1, g1inf, g2, g2inf, n, k, in12, out, is_one, err, scratch, tab, tabinf, kf, st, aux, &nl);
    ctx->launches += nl;
    if (scratch) cudaFreeAsync(scratch, st);
    if (ctx->timing) {
        cudaEventRecord(e1, st);
        d.timers.emplace_back(e0, e1);
    }
    return rc;
}

// product of n canonical Fp12 at d_in -> d_out (1 element).  scratch: zkp_product_scratch_elems(n)
static const size_t PROD_M1 = 8192, PROD_M2 = 64;
extern "C" size_t zkp_product_scratch_elems(size_t n) {
    (void)n;
    return PROD_M1 + PROD_M2;
}
static cudaError_t launch_product(zkp_ctx *ctx, const uint64_t *d_in, size_t n, uint64_t *d_scratch, uint64_t *d_out,
                                  uint32_t *err, cudaStream_t st) {
    const uint64_t *cur = d_in;
    size_t cnt = n;
    uint64_t *s1 = d_scratch, *s2 = d_scratch + 72 * PROD_M1;
    if (cnt > PROD_M1) {
        k_fp12_product<<<grid_for(PROD_M1), ZKP_TPB, 0, st>>>(cur, cnt, s1, PROD_M1, err);
        ctx->launches++;
        cur = s1;
        cnt = PROD_M1;
    }
    if (cnt > PROD_M2) {
        k_fp12_product<<<grid_for(PROD_M2), ZKP_TPB, 0, st>>>(cur, cnt, s2, PROD_M2, err);
        ctx->launches++;
        cur = s2;
        cnt = PROD_M2;
    }
    k_fp12_product<<<1, ZKP_TPB, 0, st>>>(cur, cnt, d_out, 1, err);
    ctx->launches++;
    return cudaGetLastError();
}

extern "C" {

const char *zkp_last_error(void) { return g_err.c_str(); }

extern "C" void zkp_miller_geometry(int *tpb, int *blocks, int *sync);   // pairing_kernel.cu
extern "C" void zkp_fe_geometry(int *tpb, int *blocks, int *sync);       // fe_kernel.cu
const char *zkp_version(void) {
    static char buf[224];
    int mt, mb, ms, ft, fb, fs;
    zkp_miller_geometry(&mt, &mb, &ms);
    zkp_fe_geometry(&ft, &fb, &fs);
    snprintf(buf, sizeof buf, "zkpair 0.3 (sm_100a, 2 lanes/pairing, 12x32 CIOS; miller %dx%d/SM sync %d, final exp %dx%d/SM sync %d, chunk=%u)",
             mt, mb, ms, ft, fb, fs, (unsigned)ZKP_CHUNK);
    return buf;
}

int32_t zkp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t zkp_ctx_create(const int *devices, int n_devices, zkp_ctx **out) {
    if (!out) return fail(ZKP_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int avail = zkp_device_count();
    if (avail <= 0) return fail(ZKP_ERR_NO_DEVICE, "no CUDA device available (libzkpair has no CPU path)");
    std::vector<int> ids;
    if (!devices || n_devices <= 0) {
        for (int i = 0; i < avail; i++) ids.push_back(i);
    } else {
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= avail) return fail(ZKP_ERR_INVALID_ARG, "device index out of range");
            ids.push_back(devices[i]);
        }
    }
    DeviceGuard guard;
    zkp_ctx *c = new zkp_ctx();
    c->devs.resize(ids.size());
    for (size_t i = 0; i < ids.size(); i++) {
        DevState &d = c->devs[i];
        d.id = ids[i];
        cudaError_t e = cudaSetDevice(d.id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream[0], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream[1], cudaStreamNonBlocking);
        for (int a = 0; a < 3 && e == cudaSuccess; a++) {
            e = cudaStreamCreateWithFlags(&d.fe_aux[a].s2, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.fe_aux[a].fork, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.fe_aux[a].join, cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaMalloc(&d.d_err, 4 * sizeof(uint32_t));   // [0] error flag, [1] zero word read by the probes
        if (e == cudaSuccess) e = cudaMemset(d.d_err, 0, 4 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.id);
        if (e == cudaSuccess) {
            // a private pool whose memory is never handed back at synchronisation points: the scratch
            // of one launch is reused by the next instead of being re-allocated (1.1 GB at 2^20)
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = d.id;
            e = cudaMemPoolCreate(&d.pool, &props);
            uint64_t keep = ~0ull;
            if (e == cudaSuccess) e = cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        if (e != cudaSuccess) {
            std::string msg = std::string("device init failed: ") + cudaGetErrorString(e);
            zkp_ctx_destroy(c);
            return fail(ZKP_ERR_CUDA, msg);
        }
    }
    *out = c;
    return ZKP_OK;
}

void zkp_ctx_destroy(zkp_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    for (DevState &d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        for (int a = 0; a < 3; a++) {
            if (d.fe_aux[a].fork) cudaEventDestroy(d.fe_aux[a].fork);
            if (d.fe_aux[a].join) cudaEventDestroy(d.fe_aux[a].join);
            if (d.fe_aux[a].s2) cudaStreamDestroy(d.fe_aux[a].s2);
        }
        for (auto &t : d.timers) {
            cudaEventDestroy(t.first);
            cudaEventDestroy(t.second);
        }
        for (int s = 0; s < 2; s++) {
            for (int b = 0; b < B_NBUF; b++) d.buf[s][b].release();
            for (int b = 0; b < B_NPIN; b++) {
                d.pin_in[s][b].release();
                d.pin_out[s][b].release();
            }
            if (d.stream[s]) cudaStreamDestroy(d.stream[s]);
        }
        d.scratch.release();
        d.partial.release();
        if (d.d_err) cudaFree(d.d_err);
        if (d.pool) {
            cudaDeviceSynchronize();
            cudaMemPoolDestroy(d.pool);
        }
    }
    delete ctx;
}

int32_t zkp_ctx_num_devices(const zkp_ctx *ctx) { return ctx ? (int32_t)ctx->devs.size() : 0; }
uint64_t zkp_launch_count(const zkp_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int32_t zkp_set_kernel_timing(zkp_ctx *ctx, int32_t enabled) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->timing = enabled != 0;
    return ZKP_OK;
}

int32_t zkp_last_kernel_ms(zkp_ctx *ctx, int32_t dev, double *total_ms, uint64_t *launches) {
    if (!ctx || dev < 0 || dev >= (int)ctx->devs.size()) return fail(ZKP_ERR_INVALID_ARG, "bad ctx/dev");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    for (auto &t : d.timers) {
        CU(cudaEventSynchronize(t.second));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, t.first, t.second));
        d.timed_ms += ms;
        d.timed_launches++;
        cudaEventDestroy(t.first);
        cudaEventDestroy(t.second);
    }
    d.timers.clear();
    if (total_ms) *total_ms = d.timed_ms;
    if (launches) *launches = d.timed_launches;
    d.timed_ms = 0;
    d.timed_launches = 0;
    return ZKP_OK;
}

// ------------------------------------------------------------------ device-resident entry points

static int32_t check_dev(zkp_ctx *ctx, int32_t dev) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    if (dev < 0 || dev >= (int)ctx->devs.size()) return fail(ZKP_ERR_INVALID_ARG, "device index out of range");
    return ZKP_OK;
}

int32_t zkp_pairing_dev(zkp_ctx *ctx, int32_t dev, int32_t mode, const uint64_t *d_g1_xy, const uint8_t *d_g1_inf,
                        const uint64_t *d_g2_xy, const uint8_t *d_g2_inf, size_t n_checks, int32_t pairs_per_check,
                        const uint64_t *d_in_fp12, uint64_t *d_out, uint8_t *d_is_one, uint32_t *d_err, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (!(mode == 1 || mode == 2 || mode == 3 || mode == 5) || !d_out) return fail(ZKP_ERR_INVALID_ARG, "bad mode / NULL output");
    if ((mode & 1) && (!d_g1_xy || !d_g2_xy || pairs_per_check < 1)) return fail(ZKP_ERR_INVALID_ARG, "NULL point buffers");
    if (!(mode & 1) && !d_in_fp12) return fail(ZKP_ERR_INVALID_ARG, "NULL Fp12 input");
    if (pairs_per_check > ZKP_MAX_PAIRS_PER_CHECK) return fail(ZKP_ERR_TOO_MANY_PAIRS, "pairs_per_check > 8");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    CU(launch_pairing(ctx, d, mode, d_g1_xy, d_g1_inf, d_g2_xy, d_g2_inf, n_checks, (mode & 1) ? pairs_per_check : 1,
                      d_in_fp12, d_out, d_is_one, d_err, st));
    return ZKP_OK;
}

int32_t zkp_tower_op_dev(zkp_ctx *ctx, int32_t dev, int32_t op, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out,
                         uint8_t *d_status, uint32_t *d_err, size_t n, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    int na, nb, nr;
    tower_op_shape(op, na, nb, nr);
    if (!d_a || !d_out || (nb && !d_b)) return fail(ZKP_ERR_INVALID_ARG, "NULL operand");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    if (n) {
        k_tower_op<<<grid_for(n), ZKP_TPB, 0, st>>>(op, d_a, d_b, d_out, d_status, d_err, n);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ZKP_OK;
}

int32_t zkp_fp12_product_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_in, size_t n, uint64_t *d_scratch,
                             uint64_t *d_out, uint32_t *d_err, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (!d_in || !d_out || !d_scratch || n == 0) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer or n == 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    CU(launch_product(ctx, d_in, n, d_scratch, d_out, d_err, st));
    return ZKP_OK;
}

int32_t zkp_gen_points_dev(zkp_ctx *ctx, int32_t dev, uint64_t seed, uint64_t first, size_t n, uint64_t *d_g1_xy,
                           uint8_t *d_g1_inf, uint64_t *d_g2_xy, uint8_t *d_g2_inf, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (!d_g1_xy || !d_g1_inf || !d_g2_xy || !d_g2_inf) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    if (n) {
        k_gen_points<<<grid_for(n), ZKP_TPB, 0, st>>>(seed, first, n, d_g1_xy, d_g1_inf, d_g2_xy, d_g2_inf);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ZKP_OK;
}

// ------------------------------------------------------------------ host-buffer entry points

// Split [0,n) into one contiguous slice per device.
static void slice_of(size_t n, size_t ndev, size_t d, size_t &lo, size_t &hi) {
    lo = n * d / ndev;
    hi = n * (d + 1) / ndev;
}

struct HostJob {
    int mode = 0;   // 1 miller, 2 final exp, 3 pairing ; 16 = tower op ; 32 = gen points ; 48 = group op ; 64 = Miller product
    const uint64_t *pts = nullptr, *scalars = nullptr;   // group op inputs (inf in g1inf, outputs in out / flags)
    const uint64_t *tab = nullptr;                       // prepared G2 line tables shared by all checks (kf of them)
    const uint8_t *tabinf = nullptr;
    int kf = 0;
    const uint64_t *g1 = nullptr, *g2 = nullptr, *in12 = nullptr, *a = nullptr, *b = nullptr;
    const uint8_t *g1inf = nullptr, *g2inf = nullptr;
    uint64_t *out = nullptr, *og1 = nullptr, *og2 = nullptr;
    uint8_t *flags = nullptr, *og1inf = nullptr, *og2inf = nullptr;
    int k = 1, op = 0;
    int miller_mode = 1;   // mode 64: 1 = SURVEY 9.1 line scaling (the Miller product is returned), 5 = free scaling (only Gt is)
    uint64_t seed = 0, first = 0;
};

// Runs elements [lo,hi) of a job on one device with a two-stage copy/compute pipeline.
static int32_t run_slice(zkp_ctx *ctx, DevState &d, const HostJob &j, size_t lo, size_t hi, std::string &msg) {
#define CUS(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            msg = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return ZKP_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)
    DeviceGuard guard;
    CUS(cudaSetDevice(d.id));
    CUS(cudaMemsetAsync(d.d_err, 0, sizeof(uint32_t), d.stream[0]));
    CUS(cudaStreamSynchronize(d.stream[0]));
    int na = 0, nb = 0, nr = 0;
    if (j.mode == 16) tower_op_shape(j.op, na, nb, nr);
    size_t chunk = ZKP_CHUNK;
    int s = 0;
    cudaStream_t st = nullptr;
    d.pend[0].clear();   // (a job that failed half-way leaves nothing behind for the next one)
    d.pend[1].clear();
    if (j.mode == 64) CUS(d.partial.ensure(((hi - lo + chunk - 1) / chunk + 1) * 576));
    auto pageable = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return true;
        }
        return a.type == cudaMemoryTypeUnregistered;
    };
    auto h2d = [&](int which, void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (bytes >= ZKP_STAGE_MIN && pageable(src)) {
            PinBuf &pb = d.pin_in[s][which];
            cudaError_t e = pb.ensure(bytes);
            if (e != cudaSuccess) return e;
            memcpy(pb.p, src, bytes);
            src = pb.p;
        }
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    };
    auto d2h = [&](int which, void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (bytes >= ZKP_STAGE_MIN && pageable(dst)) {
            PinBuf &pb = d.pin_out[s][which];
            cudaError_t e = pb.ensure(bytes);
            if (e != cudaSuccess) return e;
            d.pend[s].push_back(PendingOut{dst, pb.p, bytes});
            dst = pb.p;
        }
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    auto flush = [&](int set) {   // after that set's stream has drained
        for (const PendingOut &o : d.pend[set]) memcpy(o.user, o.stage, o.bytes);
        d.pend[set].clear();
    };
    for (size_t c0 = lo; c0 < hi; c0 += chunk, s ^= 1) {
        size_t cn = hi - c0 < chunk ? hi - c0 : chunk;
        st = d.stream[s];
        DevBuf *B = d.buf[s];
        CUS(cudaStreamSynchronize(st));   // this buffer set's previous chunk is fully drained
        flush(s);
        if (j.mode == 16) {
            CUS(B[B_IN].ensure(cn * na * 48));
            CUS(B[B_OUT].ensure(cn * nr * 48));
            CUS(B[B_FLAG].ensure(cn));
            CUS(h2d(B_IN, B[B_IN].p, j.a + c0 * na * 6, cn * na * 48));
            if (nb) {
                CUS(B[B_G2].ensure(cn * nb * 48));
                CUS(h2d(B_G2, B[B_G2].p, j.b + c0 * nb * 6, cn * nb * 48));
            }
            k_tower_op<<<grid_for(cn), ZKP_TPB, 0, st>>>(j.op, (const uint64_t *)B[B_IN].p, nb ? (const uint64_t *)B[B_G2].p : nullptr,
                                                        (uint64_t *)B[B_OUT].p, (uint8_t *)B[B_FLAG].p, d.d_err, cn);
            ctx->launches++;
            CUS(cudaGetLastError());
            CUS(d2h(B_OUT, j.out + c0 * nr * 6, B[B_OUT].p, cn * nr * 48));
            if (j.flags) CUS(d2h(B_FLAG, j.flags + c0, B[B_FLAG].p, cn));
        } else if (j.mode == 48) {
            const size_t w = (j.op & 1) ? 24 : 12;
            const bool is_mul = j.op >= GOP_G1_MUL;          // has a second operand and an output point
            const bool is_add = j.op >= GOP_G1_ADD;
            const size_t w2 = is_add ? w : 4;                 // u64 per element of the second operand
            CUS(B[B_IN].ensure(cn * w * 8));
            CUS(B[B_FLAG].ensure(cn));
            CUS(h2d(B_IN, B[B_IN].p, j.pts + c0 * w, cn * w * 8));
            const uint8_t *dinf = nullptr, *dinf2 = nullptr;
            if (j.g1inf) {
                CUS(B[B_G1INF].ensure(cn));
                CUS(h2d(B_G1INF, B[B_G1INF].p, j.g1inf + c0, cn));
                dinf = (const uint8_t *)B[B_G1INF].p;
            }
            if (is_add && j.g2inf) {
                CUS(B[B_G2INF].ensure(cn));
                CUS(h2d(B_G2INF, B[B_G2INF].p, j.g2inf + c0, cn));
                dinf2 = (const uint8_t *)B[B_G2INF].p;
            }
            if (is_mul) {
                CUS(B[B_G2].ensure(cn * w2 * 8));
                CUS(B[B_OUT].ensure(cn * w * 8));
                CUS(h2d(B_G2, B[B_G2].p, j.scalars + c0 * w2, cn * w2 * 8));
            }
            k_group_op<<<grid_for(cn), ZKP_TPB, 0, st>>>(j.op, (const uint64_t *)B[B_IN].p, dinf, is_mul ? (const uint64_t *)B[B_G2].p : nullptr,
                                                        is_mul ? (uint64_t *)B[B_OUT].p : nullptr, (uint8_t *)B[B_FLAG].p, d.d_err, cn, dinf2);
            ctx->launches++;
            CUS(cudaGetLastError());
            if (is_mul) CUS(d2h(B_OUT, j.out + c0 * w, B[B_OUT].p, cn * w * 8));
            CUS(d2h(B_FLAG, j.flags + c0, B[B_FLAG].p, cn));
        } else if (j.mode == 32) {
            CUS(B[B_G1].ensure(cn * 96));
            CUS(B[B_G2].ensure(cn * 192));
            CUS(B[B_G1INF].ensure(cn));
            CUS(B[B_G2INF].ensure(cn));
            k_gen_points<<<grid_for(cn), ZKP_TPB, 0, st>>>(j.seed, j.first + c0, cn, (uint64_t *)B[B_G1].p, (uint8_t *)B[B_G1INF].p,
                                                          (uint64_t *)B[B_G2].p, (uint8_t *)B[B_G2INF].p);
            ctx->launches++;
            CUS(cudaGetLastError());
            CUS(d2h(B_G1, j.og1 + c0 * 12, B[B_G1].p, cn * 96));
            CUS(d2h(B_G2, j.og2 + c0 * 24, B[B_G2].p, cn * 192));
            CUS(d2h(B_G1INF, j.og1inf + c0, B[B_G1INF].p, cn));
            CUS(d2h(B_G2INF, j.og2inf + c0, B[B_G2INF].p, cn));
        } else if (j.mode == 64) {
            // One product over all pairs of the slice: the pairs of a chunk are grouped four to a lane pair and
            // run as shared-accumulator Miller loops (one Fp12 squaring chain per FOUR pairs: 4,684 instead of
            // 6,916 Fp products per pair), the chunk's outputs are folded to one Fp12 in
            // d.partial[chunk].  Same two-stream pipeline as the independent pairings: the copies of one chunk
            // overlap the kernels of the other; nothing returns to the host per chunk.
            size_t ci = (c0 - lo) / chunk, nc4 = cn / ZKP_PRODUCT_GROUP, rem = cn % ZKP_PRODUCT_GROUP, ne = nc4 + (rem ? 1 : 0);
            const uint8_t *di1 = nullptr, *di2 = nullptr;
            CUS(B[B_G1].ensure(cn * 96));
            CUS(B[B_G2].ensure(cn * 192));
            CUS(B[B_OUT].ensure(ne * 576));
            CUS(B[B_IN].ensure(zkp_product_scratch_elems(ne) * 576));   // this buffer set's product scratch
            CUS(h2d(B_G1, B[B_G1].p, j.g1 + c0 * 12, cn * 96));
            CUS(h2d(B_G2, B[B_G2].p, j.g2 + c0 * 24, cn * 192));
            if (j.g1inf) {
                CUS(B[B_G1INF].ensure(cn));
                CUS(h2d(B_G1INF, B[B_G1INF].p, j.g1inf + c0, cn));
                di1 = (const uint8_t *)B[B_G1INF].p;
            }
            if (j.g2inf) {
                CUS(B[B_G2INF].ensure(cn));
                CUS(h2d(B_G2INF, B[B_G2INF].p, j.g2inf + c0, cn));
                di2 = (const uint8_t *)B[B_G2INF].p;
            }
            if (nc4)
                CUS(launch_pairing(ctx, d, j.miller_mode, (const uint64_t *)B[B_G1].p, di1, (const uint64_t *)B[B_G2].p, di2, nc4, ZKP_PRODUCT_GROUP,
                                   nullptr, (uint64_t *)B[B_OUT].p, nullptr, d.d_err, st));
            if (rem) {
                size_t o = nc4 * ZKP_PRODUCT_GROUP;
                CUS(launch_pairing(ctx, d, j.miller_mode, (const uint64_t *)B[B_G1].p + 12 * o, di1 ? di1 + o : nullptr,
                                   (const uint64_t *)B[B_G2].p + 24 * o, di2 ? di2 + o : nullptr, 1, (int)rem, nullptr,
                                   (uint64_t *)B[B_OUT].p + 72 * nc4, nullptr, d.d_err, st));
            }
            CUS(launch_product(ctx, (const uint64_t *)B[B_OUT].p, ne, (uint64_t *)B[B_IN].p, (uint64_t *)d.partial.p + 72 * ci, d.d_err, st));
        } else {
            size_t np = cn * (size_t)j.k, p0 = c0 * (size_t)j.k;
            size_t nq = cn * (size_t)(j.k - j.kf), q0 = c0 * (size_t)(j.k - j.kf);   // per-check G2 points
            const uint8_t *di1 = nullptr, *di2 = nullptr, *dti = nullptr;
            if (j.mode & 1) {
                CUS(B[B_G1].ensure(np * 96));
                CUS(B[B_G2].ensure(nq * 192 + 16));
                CUS(h2d(B_G1, B[B_G1].p, j.g1 + p0 * 12, np * 96));
                if (nq) CUS(h2d(B_G2, B[B_G2].p, j.g2 + q0 * 24, nq * 192));
                if (j.kf) {
                    size_t tb = (size_t)j.kf * ZKP_LINE_STEPS * 3 * 2 * sizeof(Fp);
                    CUS(B[B_TAB].ensure(tb + 16));
                    CUS(h2d(B_TAB, B[B_TAB].p, j.tab, tb));
                    if (j.tabinf) {
                        CUS(h2d(B_TABINF, (uint8_t *)B[B_TAB].p + tb, j.tabinf, j.kf));
                        dti = (const uint8_t *)B[B_TAB].p + tb;
                    }
                }
                if (j.g1inf) {
                    CUS(B[B_G1INF].ensure(np));
                    CUS(h2d(B_G1INF, B[B_G1INF].p, j.g1inf + p0, np));
                    di1 = (const uint8_t *)B[B_G1INF].p;
                }
                if (j.g2inf && nq) {
                    CUS(B[B_G2INF].ensure(nq));
                    CUS(h2d(B_G2INF, B[B_G2INF].p, j.g2inf + q0, nq));
                    di2 = (const uint8_t *)B[B_G2INF].p;
                }
            } else {
                CUS(B[B_IN].ensure(cn * 576));
                CUS(h2d(B_IN, B[B_IN].p, j.in12 + c0 * 72, cn * 576));
            }
            CUS(B[B_OUT].ensure(cn * 576));
            if (j.flags) CUS(B[B_FLAG].ensure(cn));
            CUS(launch_pairing(ctx, d, j.mode, (const uint64_t *)B[B_G1].p, di1, (const uint64_t *)B[B_G2].p, di2, cn, j.k,
                               (const uint64_t *)B[B_IN].p, (uint64_t *)B[B_OUT].p, j.flags ? (uint8_t *)B[B_FLAG].p : nullptr,
                               d.d_err, st, j.kf ? B[B_TAB].p : nullptr, dti, j.kf));
            CUS(d2h(B_OUT, j.out + c0 * 72, B[B_OUT].p, cn * 576));
            if (j.flags) CUS(d2h(B_FLAG, j.flags + c0, B[B_FLAG].p, cn));
        }
    }
    CUS(cudaStreamSynchronize(d.stream[0]));
    flush(0);
    CUS(cudaStreamSynchronize(d.stream[1]));
    flush(1);
    if (j.mode == 64) {   // fold this device's chunk partials into the slot after the last one
        size_t nchunks = (hi - lo + chunk - 1) / chunk;
        CUS(d.buf[0][B_IN].ensure(zkp_product_scratch_elems(nchunks) * 576));
        CUS(launch_product(ctx, (const uint64_t *)d.partial.p, nchunks, (uint64_t *)d.buf[0][B_IN].p,
                           (uint64_t *)d.partial.p + 72 * nchunks, d.d_err, d.stream[0]));
        CUS(cudaStreamSynchronize(d.stream[0]));
    }
    uint32_t herr = 0;
    CUS(cudaMemcpy(&herr, d.d_err, sizeof herr, cudaMemcpyDeviceToHost));
    if (herr & 1) {
        msg = "input limb vector >= p (non-canonical field element)";
        return ZKP_ERR_NONCANONICAL;
    }
    return ZKP_OK;
#undef CUS
}

static int32_t run_host_job(zkp_ctx *ctx, const HostJob &j, size_t n) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    if (n == 0) return ZKP_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    size_t nd = ctx->devs.size();
    std::vector<int32_t> rcs(nd, ZKP_OK);
    std::vector<std::string> msgs(nd);
    if (nd == 1 || n < 2 * nd) {
        rcs[0] = run_slice(ctx, ctx->devs[0], j, 0, n, msgs[0]);
    } else {
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; d++) {
            size_t lo, hi;
            slice_of(n, nd, d, lo, hi);
            th.emplace_back([&, d, lo, hi]() { rcs[d] = run_slice(ctx, ctx->devs[d], j, lo, hi, msgs[d]); });
        }
        for (auto &t : th) t.join();
    }
    for (size_t d = 0; d < nd; d++)
        if (rcs[d] != ZKP_OK) return fail(rcs[d], msgs[d]);
    return ZKP_OK;
}

int32_t zkp_tower_op_batch(zkp_ctx *ctx, int32_t op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint8_t *status, size_t n) {
    int na, nb, nr;
    tower_op_shape(op, na, nb, nr);
    bool known = (op >= 0 && op <= OP_FP_SQRT) || (op >= OP_FP2_ADD && op <= OP_FP2_POW) || (op >= OP_FP6_ADD && op <= OP_FP6_MUL_BY_01) ||
                 (op >= OP_FP12_ADD && op <= OP_FP12_POW);
    if (!known) return fail(ZKP_ERR_INVALID_ARG, "unknown tower op");
    if (n && (!a || !out || (nb && !b))) return fail(ZKP_ERR_INVALID_ARG, "NULL operand");
    HostJob j;
    j.mode = 16; j.op = op; j.a = a; j.b = b; j.out = out; j.flags = status;
    return run_host_job(ctx, j, n);
}
int32_t zkp_fp_mul_batch(zkp_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    return zkp_tower_op_batch(ctx, OP_FP_MUL, a, b, out, nullptr, n);
}
// The reference's zkVM precompile FFI, one Fp operation per call (src/fp.rs:126,376,443).  The precompiles carry no
// handle: ctx = NULL selects a process-wide context on device 0 created on first use (never destroyed).
static zkp_ctx *default_ctx(int32_t &rc) {
    static std::mutex mu;
    static zkp_ctx *ctx = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    rc = ZKP_OK;
    if (!ctx) {
        int dev0 = 0;
        rc = zkp_ctx_create(&dev0, 1, &ctx);
        if (rc != ZKP_OK) ctx = nullptr;
    }
    return ctx;
}
int32_t zkp_sys_bigint(zkp_ctx *ctx, uint32_t *result, uint32_t op, const uint32_t *lhs, const uint32_t *rhs) {
    if (!result || !lhs || !rhs) return fail(ZKP_ERR_INVALID_ARG, "NULL limb pointer");
    if (op > 1) return fail(ZKP_ERR_INVALID_ARG, "op must be 0 (mul) or 1 (add)");
    if (!ctx) {
        int32_t rc;
        ctx = default_ctx(rc);
        if (!ctx) return rc;
    }
    // twelve u32 limbs = the little-endian [u64; 6] of src/fp.rs:24 (the crate transmutes, src/fp.rs:124-128)
    uint64_t a[6], b[6], o[6];
    memcpy(a, lhs, sizeof a);
    memcpy(b, rhs, sizeof b);
    int32_t rc = zkp_tower_op_batch(ctx, op == 0 ? OP_FP_MUL : OP_FP_ADD, a, b, o, nullptr, 1);
    if (rc == ZKP_OK) memcpy(result, o, sizeof o);
    return rc;
}
int32_t zkp_syscall_fp_mulmod(zkp_ctx *ctx, uint32_t *lhs, const uint32_t *rhs) { return zkp_sys_bigint(ctx, lhs, 0, lhs, rhs); }
int32_t zkp_fp12_mul_batch(zkp_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    return zkp_tower_op_batch(ctx, OP_FP12_MUL, a, b, out, nullptr, n);
}
int32_t zkp_fp12_mul_by_014_batch(zkp_ctx *ctx, const uint64_t *f, const uint64_t *c, uint64_t *out, size_t n) {
    return zkp_tower_op_batch(ctx, OP_FP12_MUL_BY_014, f, c, out, nullptr, n);
}

static int32_t pairs_job(zkp_ctx *ctx, int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                         size_t n, int32_t k, uint64_t *out, uint8_t *flags) {
    if (k < 1) return fail(ZKP_ERR_INVALID_ARG, "pairs_per_check < 1");
    if (k > ZKP_MAX_PAIRS_PER_CHECK) return fail(ZKP_ERR_TOO_MANY_PAIRS, "pairs_per_check > 8 (use zkp_multi_miller_product)");
    if (n && (!g1 || !g2 || !out)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    HostJob j;
    j.mode = mode; j.g1 = g1; j.g1inf = g1inf; j.g2 = g2; j.g2inf = g2inf; j.k = k; j.out = out; j.flags = flags;
    return run_host_job(ctx, j, n);
}
int32_t zkp_miller_loop_batch(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf, size_t n, uint64_t *out) {
    return pairs_job(ctx, 1, g1, g1inf, g2, g2inf, n, 1, out, nullptr);
}
int32_t zkp_pairing_batch(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf, size_t n, uint64_t *out) {
    return pairs_job(ctx, 3, g1, g1inf, g2, g2inf, n, 1, out, nullptr);
}
int32_t zkp_multi_miller_loop_batch(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                    size_t n_checks, int32_t k, uint64_t *out) {
    return pairs_job(ctx, 1, g1, g1inf, g2, g2inf, n_checks, k, out, nullptr);
}
int32_t zkp_multi_pairing_batch(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                size_t n_checks, int32_t k, uint64_t *out, uint8_t *is_one) {
    return pairs_job(ctx, 3, g1, g1inf, g2, g2inf, n_checks, k, out, is_one);
}
int32_t zkp_final_exp_batch(zkp_ctx *ctx, const uint64_t *in, size_t n, uint64_t *out) {
    if (n && (!in || !out)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    HostJob j;
    j.mode = 2; j.in12 = in; j.out = out;
    return run_host_job(ctx, j, n);
}
int32_t zkp_gen_points(zkp_ctx *ctx, uint64_t seed, uint64_t first, size_t n, uint64_t *g1, uint8_t *g1inf, uint64_t *g2, uint8_t *g2inf) {
    if (n && (!g1 || !g1inf || !g2 || !g2inf)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    HostJob j;
    j.mode = 32; j.seed = seed; j.first = first; j.og1 = g1; j.og1inf = g1inf; j.og2 = g2; j.og2inf = g2inf;
    return run_host_job(ctx, j, n);
}

// ---- prepared G2 points
int32_t zkp_g2_prepare_batch(zkp_ctx *ctx, const uint64_t *g2_xy, size_t n, uint64_t *out_tables) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    if (n == 0) return ZKP_OK;
    if (!g2_xy || !out_tables) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[0];   // a handful of verifying-key points: one device
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = d.stream[0];
    DevBuf *B = d.buf[0];
    const size_t tb = (size_t)ZKP_G2_PREPARED_U64 * 8;
    CU(cudaMemsetAsync(d.d_err, 0, sizeof(uint32_t), st));
    CU(B[B_G2].ensure(n * 192));
    CU(B[B_TAB].ensure(n * tb + 16));
    CU(cudaMemcpyAsync(B[B_G2].p, g2_xy, n * 192, cudaMemcpyHostToDevice, st));
    k_g2_prepare<<<grid_for(n), ZKP_TPB, 0, st>>>((const uint64_t *)B[B_G2].p, (Fp *)B[B_TAB].p, d.d_err, n);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_tables, B[B_TAB].p, n * tb, cudaMemcpyDeviceToHost, st));
    uint32_t herr = 0;
    CU(cudaMemcpyAsync(&herr, d.d_err, sizeof herr, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (herr & 1) return fail(ZKP_ERR_NONCANONICAL, "input limb vector >= p (non-canonical field element)");
    return ZKP_OK;
}
int32_t zkp_g2_prepare_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_g2_xy, size_t n, uint64_t *d_out_tables, uint32_t *d_err, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (n && (!d_g2_xy || !d_out_tables)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    if (n) {
        k_g2_prepare<<<grid_for(n), ZKP_TPB, 0, st>>>(d_g2_xy, (Fp *)d_out_tables, d_err, n);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ZKP_OK;
}
int32_t zkp_multi_pairing_prepared_batch(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                         size_t n_checks, int32_t k, const uint64_t *tables, const uint8_t *tables_inf, int32_t kf,
                                         uint64_t *out, uint8_t *is_one) {
    if (k < 1 || kf < 0 || kf > k) return fail(ZKP_ERR_INVALID_ARG, "need 0 <= prepared pairs <= pairs_per_check");
    if (k > ZKP_MAX_PAIRS_PER_CHECK) return fail(ZKP_ERR_TOO_MANY_PAIRS, "pairs_per_check > 8");
    if (n_checks && (!g1 || !out || (kf < k && !g2) || (kf && !tables))) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    HostJob j;
    j.mode = 3; j.g1 = g1; j.g1inf = g1inf; j.g2 = g2; j.g2inf = g2inf; j.k = k; j.out = out; j.flags = is_one;
    j.tab = tables; j.tabinf = tables_inf; j.kf = kf;
    return run_host_job(ctx, j, n_checks);
}
int32_t zkp_multi_pairing_prepared_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_g1, const uint8_t *d_g1inf, const uint64_t *d_g2,
                                       const uint8_t *d_g2inf, size_t n_checks, int32_t k, const uint64_t *d_tables,
                                       const uint8_t *d_tables_inf, int32_t kf, uint64_t *d_out, uint8_t *d_is_one, uint32_t *d_err,
                                       void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (k < 1 || kf < 0 || kf > k || k > ZKP_MAX_PAIRS_PER_CHECK) return fail(ZKP_ERR_INVALID_ARG, "bad pairs_per_check / prepared count");
    if (!d_g1 || !d_out || (kf < k && !d_g2) || (kf && !d_tables)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    CU(launch_pairing(ctx, d, 3, d_g1, d_g1inf, d_g2, d_g2inf, n_checks, k, nullptr, d_out, d_is_one, d_err, st, d_tables, d_tables_inf, kf));
    return ZKP_OK;
}

// bytes <-> limbs on device-resident buffers (dir 0: from_bytes, 1: to_bytes)
int32_t zkp_fp_bytes_dev(zkp_ctx *ctx, int32_t dev, int32_t dir, const void *d_in, void *d_out, uint8_t *d_ok, size_t n, void *stream) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (n && (!d_in || !d_out)) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(ZKP_ERR_INVALID_ARG, "buffers must be 16-byte aligned");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (CUDA convention)
    if (n) {
        k_fp_bytes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dir, (const uint4 *)d_in, (uint4 *)d_out, d_ok, n);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ZKP_OK;
}
static int32_t bytes_job(zkp_ctx *ctx, int dir, const void *in, void *out, uint8_t *ok, size_t n) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    if (n == 0) return ZKP_OK;
    if (!in || !out) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[0];   // HBM-bound byte shuffling: one device is already PCIe-limited
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = d.stream[0];
    DevBuf *B = d.buf[0];
    const size_t chunk = (size_t)ZKP_CHUNK * 8;
    for (size_t c0 = 0; c0 < n; c0 += chunk) {
        size_t cn = n - c0 < chunk ? n - c0 : chunk;
        CU(B[B_IN].ensure(cn * 48));
        CU(B[B_OUT].ensure(cn * 48));
        CU(B[B_FLAG].ensure(cn));
        CU(cudaMemcpyAsync(B[B_IN].p, (const uint8_t *)in + c0 * 48, cn * 48, cudaMemcpyHostToDevice, st));
        k_fp_bytes<<<(unsigned)((cn + 255) / 256), 256, 0, st>>>(dir, (const uint4 *)B[B_IN].p, (uint4 *)B[B_OUT].p,
                                                                  ok ? (uint8_t *)B[B_FLAG].p : nullptr, cn);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync((uint8_t *)out + c0 * 48, B[B_OUT].p, cn * 48, cudaMemcpyDeviceToHost, st));
        if (ok && dir == 0) CU(cudaMemcpyAsync(ok + c0, B[B_FLAG].p, cn, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ZKP_OK;
}
int32_t zkp_fp_from_bytes_batch(zkp_ctx *ctx, const uint8_t *bytes, size_t n, uint64_t *out_limbs, uint8_t *ok) {
    return bytes_job(ctx, 0, bytes, out_limbs, ok, n);
}
int32_t zkp_fp_to_bytes_batch(zkp_ctx *ctx, const uint64_t *limbs, size_t n, uint8_t *out_bytes) {
    return bytes_job(ctx, 1, limbs, out_bytes, nullptr, n);
}

static int32_t group_job(zkp_ctx *ctx, int gop, const uint64_t *pts, const uint8_t *inf, const uint64_t *scalars, size_t n,
                         uint64_t *out, uint8_t *flags, const uint8_t *inf2 = nullptr) {
    if (n && (!pts || !flags || (gop >= GOP_G1_MUL && (!scalars || !out)))) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer");
    HostJob j;
    j.mode = 48; j.op = gop; j.pts = pts; j.g1inf = inf; j.g2inf = inf2; j.scalars = scalars; j.out = out; j.flags = flags;
    return run_host_job(ctx, j, n);
}
int32_t zkp_g1_add_batch(zkp_ctx *ctx, const uint64_t *a_xy, const uint8_t *a_inf, const uint64_t *b_xy, const uint8_t *b_inf, size_t n,
                         uint64_t *out_xy, uint8_t *out_flag) {
    return group_job(ctx, GOP_G1_ADD, a_xy, a_inf, b_xy, n, out_xy, out_flag, b_inf);
}
int32_t zkp_g2_add_batch(zkp_ctx *ctx, const uint64_t *a_xy, const uint8_t *a_inf, const uint64_t *b_xy, const uint8_t *b_inf, size_t n,
                         uint64_t *out_xy, uint8_t *out_flag) {
    return group_job(ctx, GOP_G2_ADD, a_xy, a_inf, b_xy, n, out_xy, out_flag, b_inf);
}
int32_t zkp_g1_check_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, size_t n, uint8_t *status) {
    return group_job(ctx, GOP_G1_CHECK, g1_xy, g1_inf, nullptr, n, nullptr, status);
}
int32_t zkp_g2_check_batch(zkp_ctx *ctx, const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint8_t *status) {
    return group_job(ctx, GOP_G2_CHECK, g2_xy, g2_inf, nullptr, n, nullptr, status);
}
int32_t zkp_g1_mul_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, const uint64_t *scalars, size_t n,
                         uint64_t *out_xy, uint8_t *out_inf) {
    return group_job(ctx, GOP_G1_MUL, g1_xy, g1_inf, scalars, n, out_xy, out_inf);
}
int32_t zkp_g2_mul_batch(zkp_ctx *ctx, const uint64_t *g2_xy, const uint8_t *g2_inf, const uint64_t *scalars, size_t n,
                         uint64_t *out_xy, uint8_t *out_inf) {
    return group_job(ctx, GOP_G2_MUL, g2_xy, g2_inf, scalars, n, out_xy, out_inf);
}

// One large product: per-device shared-accumulator Miller loops over a contiguous slice -> per-device Fp12
// partial -> gather the 576-byte partials on the first device -> multiply -> ONE final exponentiation.
int32_t zkp_multi_miller_product(zkp_ctx *ctx, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                                 size_t n, uint64_t *out_miller_product, uint64_t *out_gt) {
    if (!ctx) return fail(ZKP_ERR_INVALID_ARG, "ctx is NULL");
    if (!g1 || !g2 || n == 0) return fail(ZKP_ERR_INVALID_ARG, "NULL buffer or n == 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard guard;
    size_t nd = ctx->devs.size();
    if (n < 2 * nd) nd = 1;
    HostJob j;
    j.mode = 64; j.g1 = g1; j.g1inf = g1inf; j.g2 = g2; j.g2inf = g2inf;
    j.miller_mode = out_miller_product ? 1 : 5;   // the un-exponentiated product is only well defined with SURVEY 9.1's lines
    std::vector<int32_t> rcs(nd, ZKP_OK);
    std::vector<std::string> msgs(nd);
    if (nd == 1) {
        rcs[0] = run_slice(ctx, ctx->devs[0], j, 0, n, msgs[0]);
    } else {
        std::vector<std::thread> th;
        for (size_t di = 0; di < nd; di++) {
            size_t lo, hi;
            slice_of(n, nd, di, lo, hi);
            th.emplace_back([&, di, lo, hi]() { rcs[di] = run_slice(ctx, ctx->devs[di], j, lo, hi, msgs[di]); });
        }
        for (auto &t : th) t.join();
    }
    for (size_t di = 0; di < nd; di++)
        if (rcs[di] != ZKP_OK) return fail(rcs[di], msgs[di]);
    // gather: one 576-byte partial per device -> first device (peer copy over NVLink when available)
    DevState &d0 = ctx->devs[0];
    CU(cudaSetDevice(d0.id));
    DevBuf &G = d0.buf[1][B_IN];
    CU(G.ensure((nd + 2) * 576));
    for (size_t di = 0; di < nd; di++) {
        DevState &d = ctx->devs[di];
        size_t lo, hi;
        slice_of(n, nd, di, lo, hi);
        size_t nchunks = (hi - lo + ZKP_CHUNK - 1) / ZKP_CHUNK;
        const uint64_t *src = (const uint64_t *)d.partial.p + 72 * nchunks;
        if (d.id == d0.id)
            CU(cudaMemcpyAsync((uint64_t *)G.p + 72 * di, src, 576, cudaMemcpyDeviceToDevice, d0.stream[0]));
        else
            CU(cudaMemcpyPeerAsync((uint64_t *)G.p + 72 * di, d0.id, src, d.id, 576, d0.stream[0]));
    }
    uint64_t *prod = (uint64_t *)G.p + 72 * nd, *gt = (uint64_t *)G.p + 72 * (nd + 1);
    CU(d0.scratch.ensure(zkp_product_scratch_elems(nd) * 576));
    CU(launch_product(ctx, (const uint64_t *)G.p, nd, (uint64_t *)d0.scratch.p, prod, d0.d_err, d0.stream[0]));
    if (out_gt) CU(launch_pairing(ctx, d0, 2, nullptr, nullptr, nullptr, nullptr, 1, 1, prod, gt, nullptr, d0.d_err, d0.stream[0]));
    if (out_miller_product) CU(cudaMemcpyAsync(out_miller_product, prod, 576, cudaMemcpyDeviceToHost, d0.stream[0]));
    if (out_gt) CU(cudaMemcpyAsync(out_gt, gt, 576, cudaMemcpyDeviceToHost, d0.stream[0]));
    CU(cudaStreamSynchronize(d0.stream[0]));
    return ZKP_OK;
}

// ------------------------------------------------------------------ measurement

int32_t zkp_imad_peak(zkp_ctx *ctx, int32_t dev, int32_t kind, double *macs_per_second) {
    int32_t rc = check_dev(ctx, dev);
    if (rc) return rc;
    if (!macs_per_second || kind < 0 || kind > 2) return fail(ZKP_ERR_INVALID_ARG, "bad kind / NULL out");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState &d = ctx->devs[dev];
    DeviceGuard guard;
    CU(cudaSetDevice(d.id));
    cudaStream_t st = d.stream[0];
    uint32_t *sink = d.d_err;
    // every issued multiply is counted (the multiplicand update is a shift + add on the ALU pipe); the launch
    // geometry is swept and the best sustained rate is the roofline denominator
    const int iters = 1 << 13;
    const double per_thread = (double)iters * (kind == 2 ? 12.0 : 8.0);
    static const int geom[][2] = {{8, 256}, {4, 256}, {16, 128}, {8, 128}, {4, 512}};   // blocks per SM, threads per block
    struct EventPair {   // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    CU(cudaEventCreate(&ev.a));
    CU(cudaEventCreate(&ev.b));
    double best_rate = 0;
    for (const auto &g : geom) {
        const int blocks = d.sms * g[0], threads = g[1];
        for (int rep = 0; rep < 3; rep++) {
            CU(cudaEventRecord(ev.a, st));
            if (kind == 0) k_imad_peak<0><<<blocks, threads, 0, st>>>(sink + 0, iters);
            else if (kind == 1) k_imad_peak<1><<<blocks, threads, 0, st>>>(sink + 0, iters);
            else k_imad_peak<2><<<blocks, threads, 0, st>>>(sink + 0, iters);
            ctx->launches++;
            CU(cudaEventRecord(ev.b, st));
            CU(cudaEventSynchronize(ev.b));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
            double rate = per_thread * blocks * threads / (ms * 1e-3);
            if (rep > 0 && rate > best_rate) best_rate = rate;
        }
    }
    CU(cudaMemsetAsync(d.d_err, 0, sizeof(uint32_t), st));
    CU(cudaStreamSynchronize(st));
    *macs_per_second = best_rate;
    return ZKP_OK;
}

}  // extern "C"
