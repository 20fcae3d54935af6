/*
 * zkpair.h -- C ABI of the B200-native batched BLS12-381 pairing engine (libzkpair.so).
 *
 * This is the drop-in boundary for the pairing hot path of 0xWOLAND/zkvm-pairings.  Citations are
 * into /root/reference/ (the Rust crate being replaced on this path).
 *
 * What it replaces
 *   - the per-operation accelerator boundary the reference has today, the SP1 zkVM precompile FFI
 *       bls12381_sys_bigint(result:&mut[u32;12], op:u32 /0=mul,1=add/, lhs:&[u32;12], rhs:&[u32;12])   src/fp.rs:376,443
 *       syscall_bls12381_fp_mulmod(lhs:*mut u32, rhs:*const u32)                                       src/fp.rs:126
 *     -> zkp_fp_mul_batch / zkp_tower_op_batch (same canonical 12xu32 = 6xu64 little-endian limbs,
 *        but batched so one call amortises the launch); zkp_sys_bigint / zkp_syscall_fp_mulmod keep the
 *        precompiles' one-operation shape for code that is not batched yet;
 *   - the tower methods on the path (Fp/Fp2/Fp6/Fp12 mul, square, invert, frobenius_map, conjugate,
 *     mul_by_1 / mul_by_01 / mul_by_014: src/fp2.rs:147-313, src/fp6.rs:102-309, src/fp12.rs:99-210)
 *     -> zkp_tower_op_batch with the ZKP_OP_* code of the method;
 *   - the module the reference declares but leaves EMPTY (src/lib.rs:12, src/pairings.rs = 0 bytes):
 *     pairing / miller_loop / multi_miller_loop / final_exponentiation and the batch entry points
 *     the north star adds -> zkp_miller_loop_batch, zkp_final_exp_batch, zkp_pairing_batch,
 *     zkp_multi_miller_loop_batch, zkp_multi_pairing_batch, zkp_multi_miller_product.
 *
 * Conventions
 *   - Every function returns int32_t: ZKP_OK (0) or a negative ZKP_ERR_* code; zkp_last_error()
 *     returns a thread-local message.  Nothing panics or aborts.  There is NO CPU fallback: with no
 *     CUDA device zkp_ctx_create fails with ZKP_ERR_NO_DEVICE.
 *   - The caller owns every buffer.  Element layout is array-of-structs of canonical (non-
 *     Montgomery, value in [0,p)) little-endian u64 limbs, exactly `Fp.0` (src/fp.rs:24):
 *       Fp = 6 u64; Fp2 = c0|c1 (12); Fp6 = c0|c1|c2 (36); Fp12 / Gt = c0|c1 (72 u64 = 576 bytes);
 *       G1 point = x|y (12 u64), G2 point = x.c0|x.c1|y.c0|y.c1 (24 u64); the `is_infinity` flag of
 *       G1Affine/G2Affine (src/g1.rs:7-11, src/g2.rs:8-12) travels in a separate uint8_t array
 *       (NULL = no point is at infinity).  A pair with either point at infinity contributes
 *       Fp12::one() (zkcrypto lineage behaviour).
 *   - Inputs with a limb vector >= p are rejected with ZKP_ERR_NONCANONICAL (the reference's neg is
 *     undefined there, src/fp.rs:383-405); outputs are then unspecified.
 *   - Host entry points (no suffix) copy host->device, run, copy device->host, and shard the batch
 *     in contiguous slices over the devices of the context (one host thread and two streams per device,
 *     2^18-element chunks double-buffered).  Pinned (page-locked or registered) caller buffers are used in
 *     place; pageable ones are staged through pinned memory so that copies and kernels still overlap.  *_dev entry points take DEVICE
 *     pointers valid on device `dev` (an index into the context's device list), enqueue on
 *     `stream` (a cudaStream_t; NULL = the legacy default stream, as everywhere in CUDA, so the launch is
 *     ordered after the caller's earlier default-stream work) and return without synchronising; their
 *     error/status words are device resident.  Every entry point restores the caller's current device.
 *   - Calls on one zkp_ctx are serialised by an internal mutex; distinct contexts are independent.
 *   - include/zkpair.hpp is a header-only C++17 mirror of the crate's value types (Fp .. Fp12, G1Affine,
 *     G2Affine, pairings::*) over these entry points.
 */
#ifndef ZKPAIR_H
#define ZKPAIR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKP_OK 0
#define ZKP_ERR_INVALID_ARG (-1)
#define ZKP_ERR_CUDA (-2)
#define ZKP_ERR_NONCANONICAL (-3)
#define ZKP_ERR_NO_DEVICE (-4)
#define ZKP_ERR_TOO_MANY_PAIRS (-5)

#define ZKP_MAX_PAIRS_PER_CHECK 8

/* element-wise tower operations (zkp_tower_op_batch); reference method in the comment */
enum zkp_tower_op {
    ZKP_OP_FP_ADD = 0,        /* Fp::add            src/fp.rs:351-368 */
    ZKP_OP_FP_SUB = 1,        /* Fp::sub            src/fp.rs:407-411 */
    ZKP_OP_FP_NEG = 2,        /* Fp::neg            src/fp.rs:381-405 */
    ZKP_OP_FP_MUL = 3,        /* Fp::mul            src/fp.rs:413-434 */
    ZKP_OP_FP_SQR = 4,        /* Fp::square         src/fp.rs:452-455 */
    ZKP_OP_FP_INV = 5,        /* Fp::invert         src/fp.rs:306-319 (status bit1 set for zero) */
    ZKP_OP_FP_POW = 6,        /* Fp::pow_vartime    src/fp.rs:264-276; b = exponent, six RAW little-endian u64 */
    ZKP_OP_FP_SQRT = 7,       /* Fp::sqrt           src/fp.rs:280-300 (status bit1 set for a non-residue = Err(())) */
    ZKP_OP_FP2_ADD = 16, ZKP_OP_FP2_SUB = 17, ZKP_OP_FP2_NEG = 18,
    ZKP_OP_FP2_MUL = 19,      /* src/fp2.rs:192-209 */
    ZKP_OP_FP2_SQR = 20,      /* src/fp2.rs:171-189 */
    ZKP_OP_FP2_INV = 21,      /* src/fp2.rs:278-296 */
    ZKP_OP_FP2_MUL_NR = 22,   /* mul_by_nonresidue  src/fp2.rs:161-168 */
    ZKP_OP_FP2_CONJ = 23,     /* conjugate = frobenius_map  src/fp2.rs:147-157 */
    ZKP_OP_FP2_POW = 24,      /* Fp2::pow_vartime   src/fp2.rs:301-313; b = exponent, six raw u64 */
    ZKP_OP_FP6_ADD = 32, ZKP_OP_FP6_SUB = 33, ZKP_OP_FP6_NEG = 34,
    ZKP_OP_FP6_MUL = 35,      /* src/fp6.rs:188-267 */
    ZKP_OP_FP6_SQR = 36,      /* src/fp6.rs:274-288 */
    ZKP_OP_FP6_INV = 37,      /* src/fp6.rs:291-309 */
    ZKP_OP_FP6_MUL_NR = 38,   /* src/fp6.rs:128-139 */
    ZKP_OP_FP6_FROB = 39,     /* TRUE a^p; src/fp6.rs:142-176 has wrong constants */
    ZKP_OP_FP6_MUL_BY_1 = 40, /* b = c1 (Fp2)            src/fp6.rs:102-108 */
    ZKP_OP_FP6_MUL_BY_01 = 41,/* b = c0|c1 (2 Fp2)       src/fp6.rs:110-125 */
    ZKP_OP_FP12_ADD = 48, ZKP_OP_FP12_SUB = 49, ZKP_OP_FP12_NEG = 50,
    ZKP_OP_FP12_MUL = 51,     /* src/fp12.rs:193-210 */
    ZKP_OP_FP12_SQR = 52,     /* src/fp12.rs:173-184 */
    ZKP_OP_FP12_INV = 53,     /* src/fp12.rs:186-190 */
    ZKP_OP_FP12_CONJ = 54,    /* src/fp12.rs:123-125 */
    ZKP_OP_FP12_FROB = 55,    /* TRUE a^p (src/fp12.rs:143-170 inherits the Fp6 defect) */
    ZKP_OP_FP12_MUL_BY_014 = 56, /* b = c0|c1|c4 (3 Fp2) src/fp12.rs:99-111 */
    ZKP_OP_FP12_CYC_SQR = 57, /* Granger-Scott cyclotomic squaring */
    ZKP_OP_FP12_CYC_EXP = 58, /* f^x for the curve parameter x (negative): conj(f^|x|) */
    ZKP_OP_FP12_FROB2 = 59,   /* a^(p^2) */
    ZKP_OP_FP12_FROB3 = 60,   /* a^(p^3) */
    ZKP_OP_FP12_POW = 61      /* Fp12::pow_vartime  src/fp12.rs:127-139; b = exponent, six raw u64 */
};

typedef struct zkp_ctx zkp_ctx;

/* ---- context ------------------------------------------------------------------------------ */

/* Number of CUDA devices visible (0 when there is none / no driver). */
int32_t zkp_device_count(void);
/* devices == NULL or n_devices <= 0: use every visible device.  Fails with ZKP_ERR_NO_DEVICE when
 * no CUDA device is usable (there is no CPU path). */
int32_t zkp_ctx_create(const int *devices, int n_devices, zkp_ctx **out);
void zkp_ctx_destroy(zkp_ctx *ctx);
int32_t zkp_ctx_num_devices(const zkp_ctx *ctx);
/* Thread-local description of the last error returned to this thread. */
const char *zkp_last_error(void);
/* Library / build description string (arch, launch geometry). */
const char *zkp_version(void);

/* ---- tower operations (parity / test surface for the tower rows of SURVEY 8a) -------------- */

/* out[i] = op(a[i], b[i]).  b may be NULL for unary ops.  status (optional, n bytes): bit0 =
 * non-canonical input, bit1 = inverse of zero requested (the reference returns None; out = 0). */
int32_t zkp_tower_op_batch(zkp_ctx *ctx, int32_t op, const uint64_t *a, const uint64_t *b,
                           uint64_t *out, uint8_t *status, size_t n);
/* Named wrappers for the three ops SURVEY 8b lists. */
int32_t zkp_fp_mul_batch(zkp_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zkp_fp12_mul_batch(zkp_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zkp_fp12_mul_by_014_batch(zkp_ctx *ctx, const uint64_t *f, const uint64_t *c0_c1_c4,
                                  uint64_t *out, size_t n);

/* Scalar drop-ins with the argument meaning of the reference's zkVM precompile FFI (sp1_zkvm::syscalls, called at
 * src/fp.rs:126, :376, :443): ONE Fp operation on twelve little-endian u32 limbs (= the transmuted [u64; 6] of
 * src/fp.rs:124-128), canonical in and out.  zkp_sys_bigint: op 0 = lhs * rhs mod p (src/fp.rs:443), op 1 = lhs + rhs
 * mod p (src/fp.rs:376), written to `result` (which may alias an operand).  zkp_syscall_fp_mulmod: lhs <- lhs * rhs
 * mod p in place (src/fp.rs:126).  ctx = NULL uses a process-wide context on device 0, created on first use (the
 * precompiles carry no handle); unlike the precompiles they return a status (limbs >= p: ZKP_ERR_NONCANONICAL).
 * One launch per call: they exist so that the crate's `cfg(target_os = "zkvm")` bodies link unchanged; anything
 * hot belongs on the batch entry points. */
int32_t zkp_sys_bigint(zkp_ctx *ctx, uint32_t *result, uint32_t op, const uint32_t *lhs, const uint32_t *rhs);
int32_t zkp_syscall_fp_mulmod(zkp_ctx *ctx, uint32_t *lhs, const uint32_t *rhs);

/* ---- pairing path (host buffers) ---------------------------------------------------------- */

/* n independent Miller loops: out_fp12[i] = miller_loop(G1[i], G2[i])   (SURVEY 9.1 convention) */
int32_t zkp_miller_loop_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                              const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint64_t *out_fp12);
/* out_fp12[i] = in_fp12[i]^((p^12-1)/r) */
int32_t zkp_final_exp_batch(zkp_ctx *ctx, const uint64_t *in_fp12, size_t n, uint64_t *out_fp12);
/* out_gt[i] = e(G1[i], G2[i]).  Points must be curve points (zkp_g1_check_batch / zkp_g2_check_batch validate them): the
 * fused path uses line steps that rely on the curve equation, so for off-curve inputs -- where no pairing is defined -- it need
 * not agree with zkp_final_exp_batch(zkp_miller_loop_batch(.)), which applies SURVEY 9.1's formulas to whatever it is given. */
int32_t zkp_pairing_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                          const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint64_t *out_gt);
/* n_checks checks of pairs_per_check (<= ZKP_MAX_PAIRS_PER_CHECK) pairs each, stored check-major;
 * every check runs ONE shared-accumulator Miller loop over its pairs. */
int32_t zkp_multi_miller_loop_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                                    const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n_checks,
                                    int32_t pairs_per_check, uint64_t *out_fp12);
/* ... followed by one final exponentiation per check.  out_is_one (optional, n_checks bytes) is 1
 * where the product of pairings equals Gt one (e.g. a Groth16 verification equation holds). */
int32_t zkp_multi_pairing_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                                const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n_checks,
                                int32_t pairs_per_check, uint64_t *out_gt, uint8_t *out_is_one);
/* prod_i e(G1[i], G2[i]) for one large product: per-device shared-accumulator Miller loops (four pairs per
 * loop) over a contiguous slice, streamed in double-buffered chunks, per-device Fp12 partial product, gather
 * of the 576-byte partials to the first device, multiply, ONE final exponentiation.  out_miller_product
 * (optional) receives the un-exponentiated product in SURVEY 9.1's line scaling (a field element: independent of the
 * grouping and of the number of devices); with out_miller_product == NULL the cheaper line steps are used. */
int32_t zkp_multi_miller_product(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                                 const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n,
                                 uint64_t *out_miller_product, uint64_t *out_gt);

/* ---- device-resident entry points (asynchronous; pointers are device pointers on `dev`) ---- */

/* mode: 1 = Miller loop only, 2 = final exponentiation only (d_in_fp12 -> d_out), 3 = pairing, 5 = Miller loop whose
 * output is only meant for a later final exponentiation (or for products of such outputs): it equals the mode-1 value up
 * to a factor from a proper subfield, which the final exponentiation removes -- the cheaper homogeneous line steps the
 * fused pairing path uses (mode 3 == mode 2 after mode 1 == mode 2 after mode 5, bit for bit).
 * d_err (optional): device uint32_t that is OR-ed with 1 when an input is non-canonical. */
int32_t zkp_pairing_dev(zkp_ctx *ctx, int32_t dev, int32_t mode, const uint64_t *d_g1_xy,
                        const uint8_t *d_g1_inf, const uint64_t *d_g2_xy, const uint8_t *d_g2_inf,
                        size_t n_checks, int32_t pairs_per_check, const uint64_t *d_in_fp12,
                        uint64_t *d_out, uint8_t *d_is_one, uint32_t *d_err, void *stream);
int32_t zkp_tower_op_dev(zkp_ctx *ctx, int32_t dev, int32_t op, const uint64_t *d_a, const uint64_t *d_b,
                         uint64_t *d_out, uint8_t *d_status, uint32_t *d_err, size_t n, void *stream);
/* d_out (72 u64) = product of the n Fp12 elements at d_in (canonical limbs); d_scratch must hold
 * zkp_product_scratch_elems(n) * 72 u64. */
size_t zkp_product_scratch_elems(size_t n);
int32_t zkp_fp12_product_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_in, size_t n,
                             uint64_t *d_scratch, uint64_t *d_out, uint32_t *d_err, void *stream);
/* Synthetic valid inputs: G1[i] = a_i * G1gen, G2[i] = b_i * G2gen with 64-bit scalars
 * a_i = splitmix64(seed, 2*(first+i)), b_i = splitmix64(seed, 2*(first+i)+1) (zero mapped to 1). */
int32_t zkp_gen_points_dev(zkp_ctx *ctx, int32_t dev, uint64_t seed, uint64_t first, size_t n,
                           uint64_t *d_g1_xy, uint8_t *d_g1_inf, uint64_t *d_g2_xy, uint8_t *d_g2_inf,
                           void *stream);
/* Host-buffer version of the generator. */
int32_t zkp_gen_points(zkp_ctx *ctx, uint64_t seed, uint64_t first, size_t n, uint64_t *g1_xy,
                       uint8_t *g1_inf, uint64_t *g2_xy, uint8_t *g2_inf);

/* ---- prepared G2 points (`G2Prepared` of the zkcrypto lineage, absent from the reference) --
 * The 68 line-coefficient triples the Miller loop derives from a G2 point (63 doubling + 5 addition
 * steps), computed once for points that are fixed across checks -- the verifying-key points of a
 * Groth16-style check -- so that those pairs skip the G2 arithmetic in every check.  A table is an
 * opaque blob of ZKP_G2_PREPARED_U64 u64 per point in the library's internal (Montgomery) limb format,
 * valid for this build of the library only; results are bit-identical to the unprepared calls. */
#define ZKP_G2_PREPARED_U64 (68 * 3 * 12)
int32_t zkp_g2_prepare_batch(zkp_ctx *ctx, const uint64_t *g2_xy, size_t n, uint64_t *out_tables);
int32_t zkp_g2_prepare_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_g2_xy, size_t n, uint64_t *d_out_tables,
                           uint32_t *d_err, void *stream);
/* n_checks products of k pairs with a shared final exponentiation, like zkp_multi_pairing_batch, where
 * the LAST kf pairs of every check take their G2 point from `tables` (kf tables shared by all checks;
 * tables_inf = their is_infinity flags or NULL) and only the first k - kf pairs have per-check G2
 * points: g1_xy holds n_checks * k points, g2_xy holds n_checks * (k - kf). */
int32_t zkp_multi_pairing_prepared_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf,
                                         const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n_checks,
                                         int32_t pairs_per_check, const uint64_t *tables, const uint8_t *tables_inf,
                                         int32_t prepared_pairs, uint64_t *out_gt, uint8_t *out_is_one);
int32_t zkp_multi_pairing_prepared_dev(zkp_ctx *ctx, int32_t dev, const uint64_t *d_g1_xy, const uint8_t *d_g1_inf,
                                       const uint64_t *d_g2_xy, const uint8_t *d_g2_inf, size_t n_checks,
                                       int32_t pairs_per_check, const uint64_t *d_tables, const uint8_t *d_tables_inf,
                                       int32_t prepared_pairs, uint64_t *d_out_gt, uint8_t *d_is_one, uint32_t *d_err,
                                       void *stream);

/* ---- byte (de)serialisation ----------------------------------------------------------
 * Fp::from_bytes (src/fp.rs:165-191): n x 48 big-endian bytes -> n x 6 little-endian u64 limbs;
 * ok[i] = 1 when the value is canonical (< p, the reference's Ok), 0 for the reference's Err(())
 * (the limbs are still written).  Fp::to_bytes (src/fp.rs:195-207) is the inverse, no check.  An Fp2 /
 * Fp6 / Fp12 / point is a run of Fp, so the same calls serialise them coefficient by coefficient. */
int32_t zkp_fp_from_bytes_batch(zkp_ctx *ctx, const uint8_t *bytes, size_t n, uint64_t *out_limbs, uint8_t *ok);
int32_t zkp_fp_to_bytes_batch(zkp_ctx *ctx, const uint64_t *limbs, size_t n, uint8_t *out_bytes);
/* device-resident form: dir 0 = from_bytes, 1 = to_bytes; buffers 16-byte aligned; d_ok may be NULL */
int32_t zkp_fp_bytes_dev(zkp_ctx *ctx, int32_t dev, int32_t dir, const void *d_in, void *d_out, uint8_t *d_ok,
                         size_t n, void *stream);

/* ---- group-level batch operations (the callers either side of a pairing) ------------ */

/* status byte per point */
#define ZKP_POINT_OK 0                 /* G1Affine::is_valid / G2Affine::is_valid == Ok(())            */
#define ZKP_POINT_NOT_ON_CURVE 1       /* Err("Point is not on curve")        src/g1.rs:54, src/g2.rs:61 */
#define ZKP_POINT_NOT_TORSION_FREE 2   /* Err("Point is not torsion free")    src/g1.rs:57, src/g2.rs:64 */

/* G1Affine::is_valid (src/g1.rs:49-62): identity -> ok; y^2 = x^3 + 4 (src/g1.rs:95-101); subgroup test
 * -[x^2]P == (beta x, y) (src/g1.rs:103-115).  g1_inf may be NULL. */
int32_t zkp_g1_check_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, size_t n, uint8_t *status);
/* G2Affine::is_valid (src/g2.rs:57-69): y^2 = x^3 + 4(1+u) (src/g2.rs:109-120); psi(Q) == -[|x|]Q
 * (src/g2.rs:126-170). */
int32_t zkp_g2_check_batch(zkp_ctx *ctx, const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint8_t *status);
/* [k_i]P_i, k_i = 4 little-endian u64 (the limbs of an Fr, src/fr.rs): `&G1Affine * &Fr` (src/g1.rs:130-153)
 * and `&G2Affine * &Fr` (src/g2.rs:185-208) as the correct double-and-add over all 256 bits (the reference's
 * G1 loop drops bit 0 of the scalar, src/g1.rs:138-142 -- deliberately not reproduced).  Results are affine;
 * the identity is (0, 1) with out_inf = 1 (src/g1.rs:25-31). */
int32_t zkp_g1_mul_batch(zkp_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, const uint64_t *scalars,
                         size_t n, uint64_t *out_xy, uint8_t *out_inf);
int32_t zkp_g2_mul_batch(zkp_ctx *ctx, const uint64_t *g2_xy, const uint8_t *g2_inf, const uint64_t *scalars,
                         size_t n, uint64_t *out_xy, uint8_t *out_inf);

/* `&G1Affine + &G1Affine` (src/g1.rs:155-187, double() :74-91) and `&G2Affine + &G2Affine` (src/g2.rs:210-242,
 * :81-105): the affine chord-and-tangent law, element-wise a[i] + b[i].  An identity operand passes the other one
 * through, equal points double.  out_flag[i]: bit0 = the sum is the identity (0, 1); bit1 = the reference divides
 * by zero and panics here (P + (-P) src/g1.rs:177, or doubling a point with y = 0) -- the identity is returned.
 * `a - b` is a + (-b) with -b = (x, p - y) (src/g1.rs:118-128): negate on the host or with ZKP_OP_FP_NEG. */
int32_t zkp_g1_add_batch(zkp_ctx *ctx, const uint64_t *a_xy, const uint8_t *a_inf, const uint64_t *b_xy, const uint8_t *b_inf,
                         size_t n, uint64_t *out_xy, uint8_t *out_flag);
int32_t zkp_g2_add_batch(zkp_ctx *ctx, const uint64_t *a_xy, const uint8_t *a_inf, const uint64_t *b_xy, const uint8_t *b_inf,
                         size_t n, uint64_t *out_xy, uint8_t *out_flag);

/* ---- measurement helpers -------------------------------------------------------------------- */

/* Integer-multiply roofline probe: runs independent IMAD.WIDE.U32 (kind 0), IMAD lo (kind 1) or
 * carry-chained IMAD.WIDE.U32.X (kind 2) chains on every SM of device `dev` and reports the
 * sustained rate in multiply-accumulates per second (CUDA-event timed). */
int32_t zkp_imad_peak(zkp_ctx *ctx, int32_t dev, int32_t kind, double *macs_per_second);
/* Number of kernel launches issued by this context so far (for bench.py's gpu_launches). */
uint64_t zkp_launch_count(const zkp_ctx *ctx);
/* Name and average duration (ms, CUDA events on the launching stream) of the most recent timed
 * pairing-path kernel on device `dev`; timing is enabled with zkp_set_kernel_timing(ctx, 1). */
int32_t zkp_set_kernel_timing(zkp_ctx *ctx, int32_t enabled);
int32_t zkp_last_kernel_ms(zkp_ctx *ctx, int32_t dev, double *total_ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* ZKPAIR_H */
