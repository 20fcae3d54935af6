/*
 * CPU ORACLE (test infrastructure, NOT product code) -- plain-C restatement.
 *
 * Restates the hot-path arithmetic of the reference crate 0xWOLAND/zkvm-pairings
 * (/root/reference/src) and the pairing its empty src/pairings.rs never implemented.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libzkpair.so) never links or calls it.
 *
 * Parity status: tower + groups are PINNED to the reference's own known-answer vectors
 * (tests/golden/reference_kats.json) and cross-checked against oracle/pyref.py; Miller loop /
 * final exponentiation / Gt are PARITY UNPINNED by the reference (no implementation, no vector)
 * and follow SURVEY.md section 9 (zkcrypto lineage), pinned to the published e(G1,G2) value.
 *
 * The reference's host arithmetic is exact big-integer mul/add followed by "% p"
 * (num-bigint 0.4.6; src/fp.rs:351-368, :415-434).  Any exact modular arithmetic is
 * bit-identical, so this file keeps values in Montgomery form internally (6 x 64-bit limbs,
 * unsigned __int128 products) and converts at the boundary; boundary values are canonical
 * little-endian u64 limbs exactly like Fp.0 (src/fp.rs:24).  The TOWER STRUCTURE follows the
 * reference (schoolbook Fp2 mul src/fp2.rs:192-209, 36-mul interleaved Fp6 mul
 * src/fp6.rs:188-267, ...), so as a timed CPU baseline it does the reference's operation count
 * with a faster Fp primitive than the reference's heap-allocating BigUint -- i.e. it flatters
 * the reference.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 l[6]; } fp;
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

/* src/common.rs:74-81 */
static const fp MODULUS = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                            0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
/* src/common.rs:147 */
static const u64 INV = 0x89f3fffcfffcfffdULL;
/* src/common.rs:150-157 : R = 2^384 mod p = Montgomery form of 1 */
static const fp R1 = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                       0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
/* src/common.rs:72 */
#define BLS_X 0xd201000000010000ULL

static fp R2;                 /* 2^768 mod p, computed at init */
static fp2 FROB6_C1, FROB6_C2, FROB12_C1;   /* true Frobenius coefficients, computed */
static fp2 PSI_X, PSI_Y;      /* src/g2.rs:128-157, computed */
static fp BETA_M;             /* src/common.rs:83-90 */
static fp2 B2_M;              /* 4+4u, src/common.rs:70-71 */
static fp B1_M;               /* 4,    src/common.rs:69 */
static fp G1X, G1Y;           /* src/common.rs:92-108 */
static fp2 G2X, G2Y;          /* src/common.rs:110-144 */
static pthread_once_t once = PTHREAD_ONCE_INIT;

/* ------------------------------------------------------------------ Fp (src/fp.rs) */

static inline int fp_geq_p(const fp *a) {
    for (int i = 5; i >= 0; i--) {
        if (a->l[i] > MODULUS.l[i]) return 1;
        if (a->l[i] < MODULUS.l[i]) return 0;
    }
    return 1;
}
static inline void fp_sub_p(fp *a) {
    u64 br = 0;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a->l[i] - MODULUS.l[i] - br;
        a->l[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
}
static inline void fp_add(fp *r, const fp *a, const fp *b) { /* src/fp.rs:351-368 */
    u64 c = 0;
    for (int i = 0; i < 6; i++) {
        u128 s = (u128)a->l[i] + b->l[i] + c;
        r->l[i] = (u64)s;
        c = (u64)(s >> 64);
    }
    if (fp_geq_p(r)) fp_sub_p(r);   /* p < 2^381 so no carry-out */
}
static inline void fp_sub(fp *r, const fp *a, const fp *b) { /* src/fp.rs:407-411 */
    u64 br = 0;
    fp t;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - br;
        t.l[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
    if (br) {
        u64 c = 0;
        for (int i = 0; i < 6; i++) {
            u128 s = (u128)t.l[i] + MODULUS.l[i] + c;
            t.l[i] = (u64)s;
            c = (u64)(s >> 64);
        }
    }
    *r = t;
}
static inline int fp_is_zero(const fp *a) {
    return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0;
}
static inline void fp_neg(fp *r, const fp *a) { /* src/fp.rs:381-405 */
    fp z = {{0, 0, 0, 0, 0, 0}};
    fp_sub(r, &z, a);
}
static inline int fp_eq(const fp *a, const fp *b) { return memcmp(a, b, sizeof(fp)) == 0; }

/* Montgomery product a*b/2^384 mod p (value-equivalent to src/fp.rs:413-434 on canonical values) */
static void fp_mul(fp *r, const fp *a, const fp *b) {
    u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        u64 c = 0;
        for (int j = 0; j < 6; j++) {
            u128 s = (u128)a->l[j] * b->l[i] + t[j] + c;
            t[j] = (u64)s;
            c = (u64)(s >> 64);
        }
        u128 s = (u128)t[6] + c;
        t[6] = (u64)s;
        t[7] = (u64)(s >> 64);
        u64 m = t[0] * INV;
        s = (u128)m * MODULUS.l[0] + t[0];
        c = (u64)(s >> 64);
        for (int j = 1; j < 6; j++) {
            s = (u128)m * MODULUS.l[j] + t[j] + c;
            t[j - 1] = (u64)s;
            c = (u64)(s >> 64);
        }
        s = (u128)t[6] + c;
        t[5] = (u64)s;
        t[6] = t[7] + (u64)(s >> 64);
    }
    fp o;
    memcpy(&o, t, sizeof(fp));
    if (t[6] || fp_geq_p(&o)) fp_sub_p(&o);
    *r = o;
}
static inline void fp_sqr(fp *r, const fp *a) { fp_mul(r, a, a); } /* src/fp.rs:452-455 */

static void fp_from_canon(fp *r, const u64 *l) { fp t; memcpy(&t, l, 48); fp_mul(r, &t, &R2); }
static void fp_to_canon(u64 *l, const fp *a) {
    fp one = {{1, 0, 0, 0, 0, 0}}, t;
    fp_mul(&t, a, &one);
    memcpy(l, &t, 48);
}
static int canon_ok(const u64 *l) { fp t; memcpy(&t, l, 48); return !fp_geq_p(&t); }

/* src/fp.rs:264-276 : square every bit of six u64 limbs, multiply on set bits */
static void fp_pow_vartime(fp *r, const fp *a, const u64 by[6]) {
    fp res = R1;
    for (int e = 5; e >= 0; e--)
        for (int i = 63; i >= 0; i--) {
            fp_sqr(&res, &res);
            if ((by[e] >> i) & 1) fp_mul(&res, &res, a);
        }
    *r = res;
}
/* src/fp.rs:306-319 ; returns 0 when a == 0 */
static int fp_inv(fp *r, const fp *a) {
    u64 e[6];
    memcpy(e, MODULUS.l, 48);
    e[0] -= 2;
    fp_pow_vartime(r, a, e);
    return !fp_is_zero(a);
}
/* src/fp.rs:280-300 ; returns 0 when not a residue */
static int fp_sqrt(fp *r, const fp *a) {
    static const u64 e[6] = {0xee7fbfffffffeaabULL, 0x07aaffffac54ffffULL, 0xd9cc34a83dac3d89ULL,
                             0xd91dd2e13ce144afULL, 0x92c6e9ed90d2eb35ULL, 0x0680447a8e5ff9a6ULL};
    fp s, q;
    fp_pow_vartime(&s, a, e);
    fp_sqr(&q, &s);
    *r = s;
    return fp_eq(&q, a);
}

/* ------------------------------------------------------------------ Fp2 (src/fp2.rs) */

static inline void fp2_add(fp2 *r, const fp2 *a, const fp2 *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(fp2 *r, const fp2 *a, const fp2 *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(fp2 *r, const fp2 *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_dbl(fp2 *r, const fp2 *a) { fp2_add(r, a, a); }
static inline void fp2_conj(fp2 *r, const fp2 *a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); } /* :155-157 */
static inline int fp2_is_zero(const fp2 *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const fp2 *a, const fp2 *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static inline void fp2_mul_nr(fp2 *r, const fp2 *a) { /* src/fp2.rs:161-168 */
    fp t0, t1;
    fp_sub(&t0, &a->c0, &a->c1);
    fp_add(&t1, &a->c0, &a->c1);
    r->c0 = t0; r->c1 = t1;
}
static void fp2_sqr(fp2 *r, const fp2 *a) { /* src/fp2.rs:171-189 */
    fp s, d, c;
    fp_add(&s, &a->c0, &a->c1);
    fp_sub(&d, &a->c0, &a->c1);
    fp_add(&c, &a->c0, &a->c0);
    fp t1;
    fp_mul(&t1, &c, &a->c1);
    fp_mul(&r->c0, &s, &d);
    r->c1 = t1;
}
static void fp2_mul(fp2 *r, const fp2 *a, const fp2 *b) { /* src/fp2.rs:192-209 (schoolbook) */
    fp t0, t1, t2, t3;
    fp_mul(&t0, &a->c0, &b->c0);
    fp_mul(&t1, &a->c1, &b->c1);
    fp_mul(&t2, &a->c0, &b->c1);
    fp_mul(&t3, &a->c1, &b->c0);
    fp_sub(&r->c0, &t0, &t1);
    fp_add(&r->c1, &t2, &t3);
}
static inline void fp2_mul_fp(fp2 *r, const fp2 *a, const fp *k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); } /* :95-102 */
static int fp2_inv(fp2 *r, const fp2 *a) { /* src/fp2.rs:278-296 */
    fp t0, t1, t;
    fp_sqr(&t0, &a->c0);
    fp_sqr(&t1, &a->c1);
    fp_add(&t0, &t0, &t1);
    int ok = fp_inv(&t, &t0);
    fp nt;
    fp_neg(&nt, &t);
    fp_mul(&r->c0, &a->c0, &t);
    fp_mul(&r->c1, &a->c1, &nt);
    return ok;
}
static void fp2_pow_vartime(fp2 *r, const fp2 *a, const u64 by[6]) { /* src/fp2.rs:301-313 */
    fp2 res; res.c0 = R1; memset(&res.c1, 0, sizeof(fp));
    for (int e = 5; e >= 0; e--)
        for (int i = 63; i >= 0; i--) {
            fp2_sqr(&res, &res);
            if ((by[e] >> i) & 1) fp2_mul(&res, &res, a);
        }
    *r = res;
}

/* ------------------------------------------------------------------ Fp6 (src/fp6.rs) */

static inline void fp6_add(fp6 *r, const fp6 *a, const fp6 *b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static inline void fp6_sub(fp6 *r, const fp6 *a, const fp6 *b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static inline void fp6_neg(fp6 *r, const fp6 *a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static inline void fp6_mul_nr(fp6 *r, const fp6 *a) { /* src/fp6.rs:128-139 */
    fp2 t;
    fp2_mul_nr(&t, &a->c2);
    fp2 a0 = a->c0, a1 = a->c1;
    r->c0 = t; r->c1 = a0; r->c2 = a1;
}
static void fp6_mul_by_1(fp6 *r, const fp6 *a, const fp2 *c1) { /* src/fp6.rs:102-108 */
    fp2 t0, t1, t2;
    fp2_mul(&t0, &a->c2, c1);
    fp2_mul_nr(&t0, &t0);
    fp2_mul(&t1, &a->c0, c1);
    fp2_mul(&t2, &a->c1, c1);
    r->c0 = t0; r->c1 = t1; r->c2 = t2;
}
static void fp6_mul_by_01(fp6 *r, const fp6 *a, const fp2 *c0, const fp2 *c1) { /* src/fp6.rs:110-125 */
    fp2 a_a, b_b, t1, t2, t3, s0, s1;
    fp2_mul(&a_a, &a->c0, c0);
    fp2_mul(&b_b, &a->c1, c1);
    fp2_mul(&t1, &a->c2, c1);
    fp2_mul_nr(&t1, &t1);
    fp2_add(&t1, &t1, &a_a);
    fp2_add(&s0, c0, c1);
    fp2_add(&s1, &a->c0, &a->c1);
    fp2_mul(&t2, &s0, &s1);
    fp2_sub(&t2, &t2, &a_a);
    fp2_sub(&t2, &t2, &b_b);
    fp2_mul(&t3, &a->c2, c0);
    fp2_add(&t3, &t3, &b_b);
    r->c0 = t1; r->c1 = t2; r->c2 = t3;
}
/* src/fp6.rs:188-267 : mul_interleaved, 36 Fp mul, six sums of six products */
static void sop6(fp *r, const fp *a[6], const fp *b[6], const int sign[6]) {
    fp acc, t;
    memset(&acc, 0, sizeof acc);
    for (int i = 0; i < 6; i++) {
        fp_mul(&t, a[i], b[i]);
        if (sign[i] > 0) fp_add(&acc, &acc, &t); else fp_sub(&acc, &acc, &t);
    }
    *r = acc;
}
static void fp6_mul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp b10p, b10m, b20p, b20m;
    fp_add(&b10p, &b->c1.c0, &b->c1.c1);
    fp_sub(&b10m, &b->c1.c0, &b->c1.c1);
    fp_add(&b20p, &b->c2.c0, &b->c2.c1);
    fp_sub(&b20m, &b->c2.c0, &b->c2.c1);
    const fp *A[6] = {&a->c0.c0, &a->c0.c1, &a->c1.c0, &a->c1.c1, &a->c2.c0, &a->c2.c1};
    static const int pm[6] = {1, -1, 1, -1, 1, -1}, pp[6] = {1, 1, 1, 1, 1, 1};
    fp6 o;
    { const fp *B[6] = {&b->c0.c0, &b->c0.c1, &b20m, &b20p, &b10m, &b10p}; sop6(&o.c0.c0, A, B, pm); }
    { const fp *B[6] = {&b->c0.c1, &b->c0.c0, &b20p, &b20m, &b10p, &b10m}; sop6(&o.c0.c1, A, B, pp); }
    { const fp *B[6] = {&b->c1.c0, &b->c1.c1, &b->c0.c0, &b->c0.c1, &b20m, &b20p}; sop6(&o.c1.c0, A, B, pm); }
    { const fp *B[6] = {&b->c1.c1, &b->c1.c0, &b->c0.c1, &b->c0.c0, &b20p, &b20m}; sop6(&o.c1.c1, A, B, pp); }
    { const fp *B[6] = {&b->c2.c0, &b->c2.c1, &b->c1.c0, &b->c1.c1, &b->c0.c0, &b->c0.c1}; sop6(&o.c2.c0, A, B, pm); }
    { const fp *B[6] = {&b->c2.c1, &b->c2.c0, &b->c1.c1, &b->c1.c0, &b->c0.c1, &b->c0.c0}; sop6(&o.c2.c1, A, B, pp); }
    *r = o;
}
static void fp6_sqr(fp6 *r, const fp6 *a) { /* src/fp6.rs:274-288 */
    fp2 s0, ab, s1, s2, bc, s3, s4, t;
    fp2_sqr(&s0, &a->c0);
    fp2_mul(&ab, &a->c0, &a->c1);
    fp2_add(&s1, &ab, &ab);
    fp2_sub(&t, &a->c0, &a->c1);
    fp2_add(&t, &t, &a->c2);
    fp2_sqr(&s2, &t);
    fp2_mul(&bc, &a->c1, &a->c2);
    fp2_add(&s3, &bc, &bc);
    fp2_sqr(&s4, &a->c2);
    fp6 o;
    fp2_mul_nr(&t, &s3); fp2_add(&o.c0, &t, &s0);
    fp2_mul_nr(&t, &s4); fp2_add(&o.c1, &t, &s1);
    fp2_add(&t, &s1, &s2); fp2_add(&t, &t, &s3); fp2_sub(&t, &t, &s0); fp2_sub(&o.c2, &t, &s4);
    *r = o;
}
static int fp6_inv(fp6 *r, const fp6 *a) { /* src/fp6.rs:291-309 */
    fp2 c0, c1, c2, t, u;
    fp2_mul(&t, &a->c1, &a->c2); fp2_mul_nr(&t, &t); fp2_sqr(&c0, &a->c0); fp2_sub(&c0, &c0, &t);
    fp2_sqr(&c1, &a->c2); fp2_mul_nr(&c1, &c1); fp2_mul(&t, &a->c0, &a->c1); fp2_sub(&c1, &c1, &t);
    fp2_sqr(&c2, &a->c1); fp2_mul(&t, &a->c0, &a->c2); fp2_sub(&c2, &c2, &t);
    fp2_mul(&t, &a->c1, &c2); fp2_mul(&u, &a->c2, &c1); fp2_add(&t, &t, &u); fp2_mul_nr(&t, &t);
    fp2_mul(&u, &a->c0, &c0); fp2_add(&t, &t, &u);
    int ok = fp2_inv(&u, &t);
    fp2_mul(&r->c0, &u, &c0); fp2_mul(&r->c1, &u, &c1); fp2_mul(&r->c2, &u, &c2);
    return ok;
}
/* TRUE a^p; src/fp6.rs:142-176 has the shape but wrong (p^2) constants */
static void fp6_frob(fp6 *r, const fp6 *a) {
    fp2 t;
    fp2_conj(&r->c0, &a->c0);
    fp2_conj(&t, &a->c1); fp2_mul(&r->c1, &t, &FROB6_C1);
    fp2_conj(&t, &a->c2); fp2_mul(&r->c2, &t, &FROB6_C2);
}

/* ------------------------------------------------------------------ Fp12 (src/fp12.rs) */

static void fp12_one(fp12 *r) { memset(r, 0, sizeof *r); r->c0.c0.c0 = R1; }
static inline void fp12_conj(fp12 *r, const fp12 *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); } /* :123-125 */
static void fp12_mul(fp12 *r, const fp12 *a, const fp12 *b) { /* src/fp12.rs:193-210 */
    fp6 aa, bb, o, c1, c0;
    fp6_mul(&aa, &a->c0, &b->c0);
    fp6_mul(&bb, &a->c1, &b->c1);
    fp6_add(&o, &b->c0, &b->c1);
    fp6_add(&c1, &a->c1, &a->c0);
    fp6_mul(&c1, &c1, &o);
    fp6_sub(&c1, &c1, &aa);
    fp6_sub(&c1, &c1, &bb);
    fp6_mul_nr(&c0, &bb);
    fp6_add(&c0, &c0, &aa);
    r->c0 = c0; r->c1 = c1;
}
static void fp12_sqr(fp12 *r, const fp12 *a) { /* src/fp12.rs:173-184 */
    fp6 ab, c0c1, c0, c1, t;
    fp6_mul(&ab, &a->c0, &a->c1);
    fp6_add(&c0c1, &a->c0, &a->c1);
    fp6_mul_nr(&c0, &a->c1);
    fp6_add(&c0, &c0, &a->c0);
    fp6_mul(&c0, &c0, &c0c1);
    fp6_sub(&c0, &c0, &ab);
    fp6_add(&c1, &ab, &ab);
    fp6_mul_nr(&t, &ab);
    fp6_sub(&c0, &c0, &t);
    r->c0 = c0; r->c1 = c1;
}
static void fp12_mul_by_014(fp12 *r, const fp12 *a, const fp2 *c0, const fp2 *c1, const fp2 *c4) { /* src/fp12.rs:99-111 */
    fp6 aa, bb, t, r1, r0;
    fp2 o;
    fp6_mul_by_01(&aa, &a->c0, c0, c1);
    fp6_mul_by_1(&bb, &a->c1, c4);
    fp2_add(&o, c1, c4);
    fp6_add(&t, &a->c1, &a->c0);
    fp6_mul_by_01(&r1, &t, c0, &o);
    fp6_sub(&r1, &r1, &aa);
    fp6_sub(&r1, &r1, &bb);
    fp6_mul_nr(&r0, &bb);
    fp6_add(&r0, &r0, &aa);
    r->c0 = r0; r->c1 = r1;
}
static int fp12_inv(fp12 *r, const fp12 *a) { /* src/fp12.rs:186-190 */
    fp6 t0, t1, t;
    fp6_sqr(&t0, &a->c0);
    fp6_sqr(&t1, &a->c1);
    fp6_mul_nr(&t1, &t1);
    fp6_sub(&t0, &t0, &t1);
    int ok = fp6_inv(&t, &t0);
    fp6 nt;
    fp6_neg(&nt, &t);
    fp6_mul(&r->c0, &a->c0, &t);
    fp6_mul(&r->c1, &a->c1, &nt);
    return ok;
}
static void fp12_frob(fp12 *r, const fp12 *a) { /* src/fp12.rs:143-170 (with the true Fp6 map) */
    fp6 c0, c1, k;
    fp6_frob(&c0, &a->c0);
    fp6_frob(&c1, &a->c1);
    memset(&k, 0, sizeof k);
    k.c0 = FROB12_C1;
    fp6_mul(&c1, &c1, &k);            /* the reference does a FULL Fp6 mul by (k,0,0) */
    r->c0 = c0; r->c1 = c1;
}

/* ------------------------------------------------------------------ groups (src/g1.rs, src/g2.rs) */

typedef struct { fp x, y; int inf; } g1a;
typedef struct { fp2 x, y; int inf; } g2a;

static void g1_double(g1a *r, const g1a *p) { /* src/g1.rs:74-91 (affine) */
    if (p->inf) { memset(r, 0, sizeof *r); r->y = R1; r->inf = 1; return; }
    fp n, d, s, xr, yr, t;
    fp_sqr(&n, &p->x); fp_add(&t, &n, &n); fp_add(&n, &t, &n);
    fp_add(&d, &p->y, &p->y);
    fp_inv(&d, &d); fp_mul(&s, &n, &d);
    fp_sqr(&xr, &s); fp_sub(&xr, &xr, &p->x); fp_sub(&xr, &xr, &p->x);
    fp_sub(&t, &p->x, &xr); fp_mul(&yr, &s, &t); fp_sub(&yr, &yr, &p->y);
    r->x = xr; r->y = yr; r->inf = 0;
}
static void g1_add(g1a *r, const g1a *p, const g1a *q) { /* src/g1.rs:155-187 */
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    if (fp_eq(&p->x, &q->x) && fp_eq(&p->y, &q->y)) { g1_double(r, p); return; }
    /* P + (-P): the reference divides by zero and panics (src/g1.rs:177); return the identity */
    if (fp_eq(&p->x, &q->x)) { memset(r, 0, sizeof *r); r->y = R1; r->inf = 1; return; }
    fp n, d, s, xr, yr, t;
    fp_sub(&n, &q->y, &p->y); fp_sub(&d, &q->x, &p->x);
    fp_inv(&d, &d); fp_mul(&s, &n, &d);
    fp_sqr(&xr, &s); fp_sub(&xr, &xr, &p->x); fp_sub(&xr, &xr, &q->x);
    fp_sub(&t, &p->x, &xr); fp_mul(&yr, &s, &t); fp_sub(&yr, &yr, &p->y);
    r->x = xr; r->y = yr; r->inf = 0;
}
static void g2_double(g2a *r, const g2a *p) { /* src/g2.rs:81-105 */
    if (p->inf || fp2_is_zero(&p->y)) { memset(r, 0, sizeof *r); r->y.c0 = R1; r->inf = 1; return; }
    fp2 n, d, s, xr, yr, t;
    fp2_sqr(&n, &p->x); fp2_add(&t, &n, &n); fp2_add(&n, &t, &n);
    fp2_add(&d, &p->y, &p->y);
    fp2_inv(&d, &d); fp2_mul(&s, &n, &d);
    fp2_sqr(&xr, &s); fp2_sub(&xr, &xr, &p->x); fp2_sub(&xr, &xr, &p->x);
    fp2_sub(&t, &p->x, &xr); fp2_mul(&yr, &s, &t); fp2_sub(&yr, &yr, &p->y);
    r->x = xr; r->y = yr; r->inf = 0;
}
static void g2_add(g2a *r, const g2a *p, const g2a *q) { /* src/g2.rs:210-242 */
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    if (fp2_eq(&p->x, &q->x) && fp2_eq(&p->y, &q->y)) { g2_double(r, p); return; }
    if (fp2_eq(&p->x, &q->x)) { memset(r, 0, sizeof *r); r->y.c0 = R1; r->inf = 1; return; } /* src/g2.rs:232 panics */
    fp2 n, d, s, xr, yr, t;
    fp2_sub(&n, &q->y, &p->y); fp2_sub(&d, &q->x, &p->x);
    fp2_inv(&d, &d); fp2_mul(&s, &n, &d);
    fp2_sqr(&xr, &s); fp2_sub(&xr, &xr, &p->x); fp2_sub(&xr, &xr, &q->x);
    fp2_sub(&t, &p->x, &xr); fp2_mul(&yr, &s, &t); fp2_sub(&yr, &yr, &p->y);
    r->x = xr; r->y = yr; r->inf = 0;
}
/* MSB-first double-and-add over a 256-bit scalar (src/g2.rs:185-208; for G1 the CORRECT
 * algorithm, not src/g1.rs:130-153 which drops bit 0 -- SURVEY section 2) */
static void g1_mul(g1a *r, const g1a *p, const u64 k[4]) {
    g1a acc; memset(&acc, 0, sizeof acc); acc.y = R1; acc.inf = 1;
    for (int i = 255; i >= 0; i--) {
        g1_double(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) g1_add(&acc, &acc, p);
    }
    *r = acc;
}
static void g2_mul(g2a *r, const g2a *p, const u64 k[4]) {
    g2a acc; memset(&acc, 0, sizeof acc); acc.y.c0 = R1; acc.inf = 1;
    for (int i = 255; i >= 0; i--) {
        g2_double(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) g2_add(&acc, &acc, p);
    }
    *r = acc;
}
static int g1_on_curve(const g1a *p) { /* src/g1.rs:95-101 */
    fp l, r;
    fp_sqr(&l, &p->y); fp_sqr(&r, &p->x); fp_mul(&r, &r, &p->x); fp_add(&r, &r, &B1_M);
    return fp_eq(&l, &r);
}
static int g2_on_curve(const g2a *p) { /* src/g2.rs:109-120 */
    fp2 l, r;
    fp2_sqr(&l, &p->y); fp2_sqr(&r, &p->x); fp2_mul(&r, &r, &p->x); fp2_add(&r, &r, &B2_M);
    return fp2_eq(&l, &r);
}
static int g1_torsion_free(const g1a *p) { /* src/g1.rs:103-115 */
    u64 k[4] = {BLS_X, 0, 0, 0};
    g1a t, e;
    g1_mul(&t, p, k); g1_mul(&t, &t, k);
    fp_neg(&t.y, &t.y);
    fp_mul(&e.x, &p->x, &BETA_M); e.y = p->y;
    return fp_eq(&t.x, &e.x) && fp_eq(&t.y, &e.y);
}
static int g2_torsion_free(const g2a *p) { /* src/g2.rs:126-170 */
    u64 k[4] = {BLS_X, 0, 0, 0};
    g2a t, e;
    fp2 c;
    fp2_conj(&c, &p->x); fp2_mul(&e.x, &c, &PSI_X);
    fp2_conj(&c, &p->y); fp2_mul(&e.y, &c, &PSI_Y);
    g2_mul(&t, p, k);
    fp2_neg(&t.y, &t.y);
    return fp2_eq(&t.x, &e.x) && fp2_eq(&t.y, &e.y);
}

/* ------------------------------------------------------------------ pairing */

typedef struct { fp2 x, y, z; } g2p;

static void doubling_step(g2p *r, fp2 co[3]) {
    fp2 t0, t1, t2, t3, t4, t5, t6, zz, xn, yn, zn, t;
    fp2_sqr(&t0, &r->x);
    fp2_sqr(&t1, &r->y);
    fp2_sqr(&t2, &t1);
    fp2_add(&t, &t1, &r->x); fp2_sqr(&t3, &t); fp2_sub(&t3, &t3, &t0); fp2_sub(&t3, &t3, &t2);
    fp2_dbl(&t3, &t3);
    fp2_add(&t4, &t0, &t0); fp2_add(&t4, &t4, &t0);
    fp2_add(&t6, &r->x, &t4);
    fp2_sqr(&t5, &t4);
    fp2_sqr(&zz, &r->z);
    fp2_sub(&xn, &t5, &t3); fp2_sub(&xn, &xn, &t3);
    fp2_add(&t, &r->z, &r->y); fp2_sqr(&zn, &t); fp2_sub(&zn, &zn, &t1); fp2_sub(&zn, &zn, &zz);
    fp2_sub(&t, &t3, &xn); fp2_mul(&yn, &t, &t4);
    fp2_dbl(&t2, &t2); fp2_dbl(&t2, &t2); fp2_dbl(&t2, &t2);
    fp2_sub(&yn, &yn, &t2);
    fp2_mul(&co[1], &t4, &zz); fp2_dbl(&co[1], &co[1]); fp2_neg(&co[1], &co[1]);
    fp2_sqr(&co[2], &t6); fp2_sub(&co[2], &co[2], &t0); fp2_sub(&co[2], &co[2], &t5);
    fp2_dbl(&t1, &t1); fp2_dbl(&t1, &t1);
    fp2_sub(&co[2], &co[2], &t1);
    fp2_mul(&co[0], &zn, &zz); fp2_dbl(&co[0], &co[0]);
    r->x = xn; r->y = yn; r->z = zn;
}
static void addition_step(g2p *r, const fp2 *qx, const fp2 *qy, fp2 co[3]) {
    fp2 zz, yy, t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, xn, yn, zn, t;
    fp2_sqr(&zz, &r->z);
    fp2_sqr(&yy, qy);
    fp2_mul(&t0, &zz, qx);
    fp2_add(&t, qy, &r->z); fp2_sqr(&t1, &t); fp2_sub(&t1, &t1, &yy); fp2_sub(&t1, &t1, &zz); fp2_mul(&t1, &t1, &zz);
    fp2_sub(&t2, &t0, &r->x);
    fp2_sqr(&t3, &t2);
    fp2_dbl(&t4, &t3); fp2_dbl(&t4, &t4);
    fp2_mul(&t5, &t4, &t2);
    fp2_sub(&t6, &t1, &r->y); fp2_sub(&t6, &t6, &r->y);
    fp2_mul(&t9, &t6, qx);
    fp2_mul(&t7, &t4, &r->x);
    fp2_sqr(&xn, &t6); fp2_sub(&xn, &xn, &t5); fp2_sub(&xn, &xn, &t7); fp2_sub(&xn, &xn, &t7);
    fp2_add(&t, &r->z, &t2); fp2_sqr(&zn, &t); fp2_sub(&zn, &zn, &zz); fp2_sub(&zn, &zn, &t3);
    fp2_add(&t10, qy, &zn);
    fp2_sub(&t, &t7, &xn); fp2_mul(&t8, &t, &t6);
    fp2_mul(&t0, &r->y, &t5); fp2_dbl(&t0, &t0);
    fp2_sub(&yn, &t8, &t0);
    fp2_sqr(&t10, &t10); fp2_sub(&t10, &t10, &yy);
    fp2_sqr(&t, &zn); fp2_sub(&t10, &t10, &t);
    fp2_dbl(&t9, &t9); fp2_sub(&t9, &t9, &t10);
    fp2_dbl(&co[0], &zn);
    fp2_neg(&t6, &t6); fp2_dbl(&co[1], &t6);
    co[2] = t9;
    r->x = xn; r->y = yn; r->z = zn;
}
static void ell(fp12 *f, const fp2 co[3], const g1a *p) {
    fp2 a, b;
    fp2_mul_fp(&a, &co[0], &p->y);
    fp2_mul_fp(&b, &co[1], &p->x);
    fp12_mul_by_014(f, f, &co[2], &b, &a);
}
/* k pairs sharing one accumulator; pairs with a point at infinity contribute one */
static void multi_miller(fp12 *out, const g1a *ps, const g2a *qs, int k) {
    g2p *rs = (g2p *)malloc(sizeof(g2p) * (k ? k : 1));
    fp2 co[3];
    fp12 f;
    fp12_one(&f);
    for (int j = 0; j < k; j++) { rs[j].x = qs[j].x; rs[j].y = qs[j].y; memset(&rs[j].z, 0, sizeof(fp2)); rs[j].z.c0 = R1; }
    int found = 0;
    for (int b = 63; b >= 0; b--) {
        int i = (int)(((BLS_X >> 1) >> b) & 1);
        if (!found) { found = i; continue; }
        for (int j = 0; j < k; j++) if (!ps[j].inf && !qs[j].inf) { doubling_step(&rs[j], co); ell(&f, co, &ps[j]); }
        if (i) for (int j = 0; j < k; j++) if (!ps[j].inf && !qs[j].inf) { addition_step(&rs[j], &qs[j].x, &qs[j].y, co); ell(&f, co, &ps[j]); }
        fp12_sqr(&f, &f);
    }
    for (int j = 0; j < k; j++) if (!ps[j].inf && !qs[j].inf) { doubling_step(&rs[j], co); ell(&f, co, &ps[j]); }
    fp12_conj(out, &f);
    free(rs);
}
static void fp4_square(fp2 *c0, fp2 *c1, const fp2 *a, const fp2 *b) {
    fp2 t0, t1, t2;
    fp2_sqr(&t0, a);
    fp2_sqr(&t1, b);
    fp2_mul_nr(&t2, &t1);
    fp2_add(c0, &t2, &t0);
    fp2_add(&t2, a, b); fp2_sqr(&t2, &t2); fp2_sub(&t2, &t2, &t0); fp2_sub(c1, &t2, &t1);
}
static void cyclotomic_square(fp12 *r, const fp12 *f) {
    fp2 z0 = f->c0.c0, z4 = f->c0.c1, z3 = f->c0.c2, z2 = f->c1.c0, z1 = f->c1.c1, z5 = f->c1.c2;
    fp2 t0, t1, t2, t3;
    fp4_square(&t0, &t1, &z0, &z1);
    fp2_sub(&z0, &t0, &z0); fp2_dbl(&z0, &z0); fp2_add(&z0, &z0, &t0);
    fp2_add(&z1, &t1, &z1); fp2_dbl(&z1, &z1); fp2_add(&z1, &z1, &t1);
    fp4_square(&t0, &t1, &z2, &z3);
    fp4_square(&t2, &t3, &z4, &z5);
    fp2_sub(&z4, &t0, &z4); fp2_dbl(&z4, &z4); fp2_add(&z4, &z4, &t0);
    fp2_add(&z5, &t1, &z5); fp2_dbl(&z5, &z5); fp2_add(&z5, &z5, &t1);
    fp2_mul_nr(&t0, &t3);
    fp2_add(&z2, &t0, &z2); fp2_dbl(&z2, &z2); fp2_add(&z2, &z2, &t0);
    fp2_sub(&z3, &t2, &z3); fp2_dbl(&z3, &z3); fp2_add(&z3, &z3, &t2);
    r->c0.c0 = z0; r->c0.c1 = z4; r->c0.c2 = z3; r->c1.c0 = z2; r->c1.c1 = z1; r->c1.c2 = z5;
}
static void cyclotomic_exp(fp12 *r, const fp12 *f) {
    fp12 tmp;
    fp12_one(&tmp);
    int found = 0;
    for (int b = 63; b >= 0; b--) {
        int i = (int)((BLS_X >> b) & 1);
        if (found) cyclotomic_square(&tmp, &tmp); else found = i;
        if (i) fp12_mul(&tmp, &tmp, f);
    }
    fp12_conj(r, &tmp);
}
static int final_exp(fp12 *r, const fp12 *f) {
    fp12 t0, t1, t2, t3, t4, t5, t6, t;
    t0 = *f;
    for (int i = 0; i < 6; i++) fp12_frob(&t0, &t0);
    if (!fp12_inv(&t1, f)) { memset(r, 0, sizeof *r); return 0; }
    fp12_mul(&t2, &t0, &t1);
    t1 = t2;
    fp12_frob(&t2, &t2); fp12_frob(&t2, &t2);
    fp12_mul(&t2, &t2, &t1);
    cyclotomic_square(&t1, &t2); fp12_conj(&t1, &t1);
    cyclotomic_exp(&t3, &t2);
    cyclotomic_square(&t4, &t3);
    fp12_mul(&t5, &t1, &t3);
    cyclotomic_exp(&t1, &t5);
    cyclotomic_exp(&t0, &t1);
    cyclotomic_exp(&t6, &t0);
    fp12_mul(&t6, &t6, &t4);
    cyclotomic_exp(&t4, &t6);
    fp12_conj(&t5, &t5);
    fp12_mul(&t, &t5, &t2); fp12_mul(&t4, &t4, &t);
    fp12_conj(&t5, &t2);
    fp12_mul(&t1, &t1, &t2);
    fp12_frob(&t1, &t1); fp12_frob(&t1, &t1); fp12_frob(&t1, &t1);
    fp12_mul(&t6, &t6, &t5);
    fp12_frob(&t6, &t6);
    fp12_mul(&t3, &t3, &t0);
    fp12_frob(&t3, &t3); fp12_frob(&t3, &t3);
    fp12_mul(&t3, &t3, &t1);
    fp12_mul(&t3, &t3, &t6);
    fp12_mul(r, &t3, &t4);
    return 1;
}

/* ------------------------------------------------------------------ init + boundary */

static void set_fp2_pow(fp2 *r, const u64 e[6]) {
    fp2 b; b.c0 = R1; b.c1 = R1; /* u+1 */
    fp2_pow_vartime(r, &b, e);
}
/* e = (p-1)/d * m  for tiny d, m, computed with schoolbook limb ops */
static void exp_of(u64 e[6], unsigned mul, unsigned div) {
    u64 t[6]; memcpy(t, MODULUS.l, 48); t[0] -= 1;
    u64 c = 0;
    for (int i = 0; i < 6; i++) { u128 s = (u128)t[i] * mul + c; t[i] = (u64)s; c = (u64)(s >> 64); }
    u128 rem = 0;
    for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | t[i]; e[i] = (u64)(cur / div); rem = cur % div; }
}
static void do_init(void) {
    /* R2 = 2^768 mod p : double R1 (=2^384 mod p, canonical integer) 384 times */
    fp t = R1;
    for (int i = 0; i < 384; i++) fp_add(&t, &t, &t);
    R2 = t;
    u64 e[6];
    exp_of(e, 1, 3); set_fp2_pow(&FROB6_C1, e);
    exp_of(e, 2, 3); set_fp2_pow(&FROB6_C2, e);
    exp_of(e, 1, 6); set_fp2_pow(&FROB12_C1, e);
    /* psi coefficients src/g2.rs:128-157 : 1/(u+1)^((p-1)/3), 1/(u+1)^((p-1)/2) */
    fp2_inv(&PSI_X, &FROB6_C1);
    exp_of(e, 1, 2); set_fp2_pow(&PSI_Y, e); fp2_inv(&PSI_Y, &PSI_Y);
    static const u64 beta[6] = {0x2e01fffffffefffeULL, 0xde17d813620a0002ULL, 0xddb3a93be6f89688ULL,
                                0xba69c6076a0f77eaULL, 0x5f19672fdf76ce51ULL, 0x0ULL};
    fp_from_canon(&BETA_M, beta);
    static const u64 four[6] = {4, 0, 0, 0, 0, 0};
    fp_from_canon(&B1_M, four); B2_M.c0 = B1_M; B2_M.c1 = B1_M;
    static const u64 g1x[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL, 0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const u64 g1y[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL, 0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    static const u64 g2x0[6] = {0xd48056c8c121bdb8ULL, 0x0bac0326a805bbefULL, 0xb4510b647ae3d177ULL, 0xc6e47ad4fa403b02ULL, 0x260805272dc51051ULL, 0x024aa2b2f08f0a91ULL};
    static const u64 g2x1[6] = {0xe5ac7d055d042b7eULL, 0x334cf11213945d57ULL, 0xb5da61bbdc7f5049ULL, 0x596bd0d09920b61aULL, 0x7dacd3a088274f65ULL, 0x13e02b6052719f60ULL};
    static const u64 g2y0[6] = {0xe193548608b82801ULL, 0x923ac9cc3baca289ULL, 0x6d429a695160d12cULL, 0xadfd9baa8cbdd3a7ULL, 0x8cc9cdc6da2e351aULL, 0x0ce5d527727d6e11ULL};
    static const u64 g2y1[6] = {0xaaa9075ff05f79beULL, 0x3f370d275cec1da1ULL, 0x267492ab572e99abULL, 0xcb3e287e85a763afULL, 0x32acd2b02bc28b99ULL, 0x0606c4a02ea734ccULL};
    fp_from_canon(&G1X, g1x); fp_from_canon(&G1Y, g1y);
    fp_from_canon(&G2X.c0, g2x0); fp_from_canon(&G2X.c1, g2x1);
    fp_from_canon(&G2Y.c0, g2y0); fp_from_canon(&G2Y.c1, g2y1);
}
static void init(void) { pthread_once(&once, do_init); }

static void load_fp2(fp2 *r, const u64 *l) { fp_from_canon(&r->c0, l); fp_from_canon(&r->c1, l + 6); }
static void store_fp2(u64 *l, const fp2 *a) { fp_to_canon(l, &a->c0); fp_to_canon(l + 6, &a->c1); }
static void load_fp6(fp6 *r, const u64 *l) { load_fp2(&r->c0, l); load_fp2(&r->c1, l + 12); load_fp2(&r->c2, l + 24); }
static void store_fp6(u64 *l, const fp6 *a) { store_fp2(l, &a->c0); store_fp2(l + 12, &a->c1); store_fp2(l + 24, &a->c2); }
static void load_fp12(fp12 *r, const u64 *l) { load_fp6(&r->c0, l); load_fp6(&r->c1, l + 36); }
static void store_fp12(u64 *l, const fp12 *a) { store_fp6(l, &a->c0); store_fp6(l + 36, &a->c1); }
static void load_g1(g1a *p, const u64 *xy, uint8_t inf) { fp_from_canon(&p->x, xy); fp_from_canon(&p->y, xy + 6); p->inf = inf != 0; }
static void load_g2(g2a *p, const u64 *xy, uint8_t inf) { load_fp2(&p->x, xy); load_fp2(&p->y, xy + 12); p->inf = inf != 0; }
static void store_g1(u64 *xy, uint8_t *inf, const g1a *p) { fp_to_canon(xy, &p->x); fp_to_canon(xy + 6, &p->y); *inf = (uint8_t)p->inf; }
static void store_g2(u64 *xy, uint8_t *inf, const g2a *p) { store_fp2(xy, &p->x); store_fp2(xy + 12, &p->y); *inf = (uint8_t)p->inf; }

/* ------------------------------------------------------------------ thread pool helper */

typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }
static void parallel_for(range_fn fn, void *ctx, size_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    if (nthreads == 1) { fn(ctx, 0, n); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * (size_t)t / nthreads; jobs[t].hi = n * (size_t)(t + 1) / nthreads;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------ exported API
 * All buffers: canonical little-endian u64 limbs, array-of-structs; Fp=6, Fp2=12, Fp6=36,
 * Fp12=72 u64; G1 = x|y (12 u64), G2 = x.c0|x.c1|y.c0|y.c1 (24 u64), infinity flags separate.
 * Return 0 on success, -1 if any input limb vector is >= p.                                   */

#define EXPORT __attribute__((visibility("default")))

static int all_canon(const u64 *l, size_t n_fp) { for (size_t i = 0; i < n_fp; i++) if (!canon_ok(l + 6 * i)) return 0; return 1; }

/* op codes for zo_tower_op */
enum { OP_FP_ADD = 0, OP_FP_SUB, OP_FP_NEG, OP_FP_MUL, OP_FP_SQR, OP_FP_INV, OP_FP_SQRT,
       OP_FP2_ADD = 16, OP_FP2_SUB, OP_FP2_NEG, OP_FP2_MUL, OP_FP2_SQR, OP_FP2_INV, OP_FP2_MUL_NR, OP_FP2_CONJ,
       OP_FP6_ADD = 32, OP_FP6_SUB, OP_FP6_NEG, OP_FP6_MUL, OP_FP6_SQR, OP_FP6_INV, OP_FP6_MUL_NR, OP_FP6_FROB, OP_FP6_MUL_BY_1, OP_FP6_MUL_BY_01,
       OP_FP12_ADD = 48, OP_FP12_SUB, OP_FP12_NEG, OP_FP12_MUL, OP_FP12_SQR, OP_FP12_INV, OP_FP12_CONJ, OP_FP12_FROB, OP_FP12_MUL_BY_014, OP_FP12_CYC_SQR, OP_FP12_CYC_EXP };

/* Generic element-wise tower op over n elements.  a, b, c: operand arrays (b/c may be NULL when
 * unused).  For MUL_BY_1: b = c1 (Fp2).  MUL_BY_01: b = c0|c1 (2 Fp2).  MUL_BY_014: b = c0|c1|c4
 * (3 Fp2).  ok (optional, n bytes): 0 where an inverse / sqrt does not exist.                  */
EXPORT int zo_tower_op(int op, const u64 *a, const u64 *b, u64 *out, uint8_t *ok, size_t n) {
    init();
    for (size_t i = 0; i < n; i++) {
        int good = 1;
        switch (op) {
#define FP_BIN(OP, F) case OP: { fp x, y, r; fp_from_canon(&x, a + 6 * i); fp_from_canon(&y, b + 6 * i); F(&r, &x, &y); fp_to_canon(out + 6 * i, &r); } break;
            FP_BIN(OP_FP_ADD, fp_add) FP_BIN(OP_FP_SUB, fp_sub) FP_BIN(OP_FP_MUL, fp_mul)
            case OP_FP_NEG: { fp x, r; fp_from_canon(&x, a + 6 * i); fp_neg(&r, &x); fp_to_canon(out + 6 * i, &r); } break;
            case OP_FP_SQR: { fp x, r; fp_from_canon(&x, a + 6 * i); fp_sqr(&r, &x); fp_to_canon(out + 6 * i, &r); } break;
            case OP_FP_INV: { fp x, r; fp_from_canon(&x, a + 6 * i); good = fp_inv(&r, &x); fp_to_canon(out + 6 * i, &r); } break;
            case OP_FP_SQRT: { fp x, r; fp_from_canon(&x, a + 6 * i); good = fp_sqrt(&r, &x); fp_to_canon(out + 6 * i, &r); } break;
#define FP2_BIN(OP, F) case OP: { fp2 x, y, r; load_fp2(&x, a + 12 * i); load_fp2(&y, b + 12 * i); F(&r, &x, &y); store_fp2(out + 12 * i, &r); } break;
#define FP2_UN(OP, F) case OP: { fp2 x, r; load_fp2(&x, a + 12 * i); F(&r, &x); store_fp2(out + 12 * i, &r); } break;
            FP2_BIN(OP_FP2_ADD, fp2_add) FP2_BIN(OP_FP2_SUB, fp2_sub) FP2_BIN(OP_FP2_MUL, fp2_mul)
            FP2_UN(OP_FP2_NEG, fp2_neg) FP2_UN(OP_FP2_SQR, fp2_sqr) FP2_UN(OP_FP2_MUL_NR, fp2_mul_nr) FP2_UN(OP_FP2_CONJ, fp2_conj)
            case OP_FP2_INV: { fp2 x, r; load_fp2(&x, a + 12 * i); good = fp2_inv(&r, &x); store_fp2(out + 12 * i, &r); } break;
#define FP6_BIN(OP, F) case OP: { fp6 x, y, r; load_fp6(&x, a + 36 * i); load_fp6(&y, b + 36 * i); F(&r, &x, &y); store_fp6(out + 36 * i, &r); } break;
#define FP6_UN(OP, F) case OP: { fp6 x, r; load_fp6(&x, a + 36 * i); F(&r, &x); store_fp6(out + 36 * i, &r); } break;
            FP6_BIN(OP_FP6_ADD, fp6_add) FP6_BIN(OP_FP6_SUB, fp6_sub) FP6_BIN(OP_FP6_MUL, fp6_mul)
            FP6_UN(OP_FP6_NEG, fp6_neg) FP6_UN(OP_FP6_SQR, fp6_sqr) FP6_UN(OP_FP6_MUL_NR, fp6_mul_nr) FP6_UN(OP_FP6_FROB, fp6_frob)
            case OP_FP6_INV: { fp6 x, r; load_fp6(&x, a + 36 * i); good = fp6_inv(&r, &x); store_fp6(out + 36 * i, &r); } break;
            case OP_FP6_MUL_BY_1: { fp6 x, r; fp2 c1; load_fp6(&x, a + 36 * i); load_fp2(&c1, b + 12 * i); fp6_mul_by_1(&r, &x, &c1); store_fp6(out + 36 * i, &r); } break;
            case OP_FP6_MUL_BY_01: { fp6 x, r; fp2 c0, c1; load_fp6(&x, a + 36 * i); load_fp2(&c0, b + 24 * i); load_fp2(&c1, b + 24 * i + 12); fp6_mul_by_01(&r, &x, &c0, &c1); store_fp6(out + 36 * i, &r); } break;
#define FP12_BIN(OP, F) case OP: { fp12 x, y, r; load_fp12(&x, a + 72 * i); load_fp12(&y, b + 72 * i); F(&r, &x, &y); store_fp12(out + 72 * i, &r); } break;
#define FP12_UN(OP, F) case OP: { fp12 x, r; load_fp12(&x, a + 72 * i); F(&r, &x); store_fp12(out + 72 * i, &r); } break;
            case OP_FP12_ADD: { fp12 x, y, r; load_fp12(&x, a + 72 * i); load_fp12(&y, b + 72 * i); fp6_add(&r.c0, &x.c0, &y.c0); fp6_add(&r.c1, &x.c1, &y.c1); store_fp12(out + 72 * i, &r); } break;
            case OP_FP12_SUB: { fp12 x, y, r; load_fp12(&x, a + 72 * i); load_fp12(&y, b + 72 * i); fp6_sub(&r.c0, &x.c0, &y.c0); fp6_sub(&r.c1, &x.c1, &y.c1); store_fp12(out + 72 * i, &r); } break;
            case OP_FP12_NEG: { fp12 x, r; load_fp12(&x, a + 72 * i); fp6_neg(&r.c0, &x.c0); fp6_neg(&r.c1, &x.c1); store_fp12(out + 72 * i, &r); } break;
            FP12_BIN(OP_FP12_MUL, fp12_mul)
            FP12_UN(OP_FP12_SQR, fp12_sqr) FP12_UN(OP_FP12_CONJ, fp12_conj) FP12_UN(OP_FP12_FROB, fp12_frob)
            FP12_UN(OP_FP12_CYC_SQR, cyclotomic_square) FP12_UN(OP_FP12_CYC_EXP, cyclotomic_exp)
            case OP_FP12_INV: { fp12 x, r; load_fp12(&x, a + 72 * i); good = fp12_inv(&r, &x); store_fp12(out + 72 * i, &r); } break;
            case OP_FP12_MUL_BY_014: { fp12 x, r; fp2 c0, c1, c4; load_fp12(&x, a + 72 * i); load_fp2(&c0, b + 36 * i); load_fp2(&c1, b + 36 * i + 12); load_fp2(&c4, b + 36 * i + 24); fp12_mul_by_014(&r, &x, &c0, &c1, &c4); store_fp12(out + 72 * i, &r); } break;
            default: return -2;
        }
        if (ok) ok[i] = (uint8_t)good;
    }
    return 0;
}

typedef struct {
    const u64 *g1, *g2; const uint8_t *g1inf, *g2inf;
    const u64 *in12; u64 *out; uint8_t *flags; int k; int mode;
} pjob;

static void pair_range(void *vc, size_t lo, size_t hi) {
    pjob *c = (pjob *)vc;
    int k = c->k;
    g1a *ps = (g1a *)malloc(sizeof(g1a) * (k ? k : 1));
    g2a *qs = (g2a *)malloc(sizeof(g2a) * (k ? k : 1));
    for (size_t i = lo; i < hi; i++) {
        fp12 f, g;
        if (c->mode == 1) { /* final exponentiation only */
            load_fp12(&f, c->in12 + 72 * i);
            final_exp(&g, &f);
            store_fp12(c->out + 72 * i, &g);
            continue;
        }
        for (int j = 0; j < k; j++) {
            size_t e = i * (size_t)k + j;
            load_g1(&ps[j], c->g1 + 12 * e, c->g1inf ? c->g1inf[e] : 0);
            load_g2(&qs[j], c->g2 + 24 * e, c->g2inf ? c->g2inf[e] : 0);
        }
        multi_miller(&f, ps, qs, k);
        if (c->mode == 0) { store_fp12(c->out + 72 * i, &f); continue; }   /* Miller only */
        final_exp(&g, &f);                                               /* mode 2: pairing */
        store_fp12(c->out + 72 * i, &g);
        if (c->flags) { fp12 one; fp12_one(&one); c->flags[i] = memcmp(&g, &one, sizeof one) == 0; }
    }
    free(ps); free(qs);
}

/* n checks of k pairs each (k=1: independent pairs).  mode 0: Miller loop outputs, 2: Gt. */
static int run_pairs(int mode, const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf,
                     size_t n, int k, u64 *out, uint8_t *flags, int nthreads) {
    init();
    if (!all_canon(g1, n * k * 2) || !all_canon(g2, n * k * 4)) return -1;
    pjob c = {g1, g2, g1inf, g2inf, NULL, out, flags, k, mode};
    parallel_for(pair_range, &c, n, nthreads);
    return 0;
}
EXPORT int zo_miller_loop_batch(const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf, size_t n, u64 *out, int nthreads) {
    return run_pairs(0, g1, g1inf, g2, g2inf, n, 1, out, NULL, nthreads);
}
EXPORT int zo_pairing_batch(const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf, size_t n, u64 *out, int nthreads) {
    return run_pairs(2, g1, g1inf, g2, g2inf, n, 1, out, NULL, nthreads);
}
EXPORT int zo_multi_miller_batch(const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf, size_t n_checks, int k, u64 *out, int nthreads) {
    return run_pairs(0, g1, g1inf, g2, g2inf, n_checks, k, out, NULL, nthreads);
}
EXPORT int zo_multi_pairing_batch(const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf, size_t n_checks, int k, u64 *out, uint8_t *is_one, int nthreads) {
    return run_pairs(2, g1, g1inf, g2, g2inf, n_checks, k, out, is_one, nthreads);
}
EXPORT int zo_final_exp_batch(const u64 *in, size_t n, u64 *out, int nthreads) {
    init();
    if (!all_canon(in, n * 12)) return -1;
    pjob c = {NULL, NULL, NULL, NULL, in, out, NULL, 1, 1};
    parallel_for(pair_range, &c, n, nthreads);
    return 0;
}
/* product of n Miller loops (each pair its own loop, multiplied) followed by ONE final exp.
 * Equal bit-for-bit to the shared-accumulator loop.                               */
EXPORT int zo_miller_product(const u64 *g1, const uint8_t *g1inf, const u64 *g2, const uint8_t *g2inf, size_t n, u64 *out_miller_prod, u64 *out_gt) {
    init();
    if (!all_canon(g1, n * 2) || !all_canon(g2, n * 4)) return -1;
    fp12 acc, f, g;
    fp12_one(&acc);
    for (size_t i = 0; i < n; i++) {
        g1a p; g2a q;
        load_g1(&p, g1 + 12 * i, g1inf ? g1inf[i] : 0);
        load_g2(&q, g2 + 24 * i, g2inf ? g2inf[i] : 0);
        multi_miller(&f, &p, &q, 1);
        fp12_mul(&acc, &acc, &f);
    }
    if (out_miller_prod) store_fp12(out_miller_prod, &acc);
    if (out_gt) { final_exp(&g, &acc); store_fp12(out_gt, &g); }
    return 0;
}

/* [k]P for affine points, 256-bit scalars (4 LE u64 each); base==NULL means the generator. */
typedef struct { const u64 *base; const uint8_t *binf; const u64 *k; u64 *out; uint8_t *oinf; } mjob;
static void g1mul_range(void *vc, size_t lo, size_t hi) {
    mjob *c = (mjob *)vc;
    for (size_t i = lo; i < hi; i++) {
        g1a p, r;
        if (c->base) load_g1(&p, c->base + 12 * i, c->binf ? c->binf[i] : 0); else { p.x = G1X; p.y = G1Y; p.inf = 0; }
        g1_mul(&r, &p, c->k + 4 * i);
        store_g1(c->out + 12 * i, &c->oinf[i], &r);
    }
}
static void g2mul_range(void *vc, size_t lo, size_t hi) {
    mjob *c = (mjob *)vc;
    for (size_t i = lo; i < hi; i++) {
        g2a p, r;
        if (c->base) load_g2(&p, c->base + 24 * i, c->binf ? c->binf[i] : 0); else { p.x = G2X; p.y = G2Y; p.inf = 0; }
        g2_mul(&r, &p, c->k + 4 * i);
        store_g2(c->out + 24 * i, &c->oinf[i], &r);
    }
}
EXPORT int zo_g1_mul_batch(const u64 *base, const uint8_t *binf, const u64 *scalars, size_t n, u64 *out, uint8_t *oinf, int nthreads) {
    init();
    if (base && !all_canon(base, n * 2)) return -1;
    mjob c = {base, binf, scalars, out, oinf};
    parallel_for(g1mul_range, &c, n, nthreads);
    return 0;
}
EXPORT int zo_g2_mul_batch(const u64 *base, const uint8_t *binf, const u64 *scalars, size_t n, u64 *out, uint8_t *oinf, int nthreads) {
    init();
    if (base && !all_canon(base, n * 4)) return -1;
    mjob c = {base, binf, scalars, out, oinf};
    parallel_for(g2mul_range, &c, n, nthreads);
    return 0;
}
/* group ops for pinning against the reference KATs: op 0 double, 1 add(a,b), 2 on_curve,
 * 3 torsion_free (result in *flag)                                                           */
EXPORT int zo_g1_op(int op, const u64 *a, uint8_t ainf, const u64 *b, uint8_t binf, u64 *out, uint8_t *flag) {
    init();
    g1a p, q, r;
    load_g1(&p, a, ainf);
    if (op == 0) { g1_double(&r, &p); store_g1(out, flag, &r); }
    else if (op == 1) { load_g1(&q, b, binf); g1_add(&r, &p, &q); store_g1(out, flag, &r); }
    else if (op == 2) *flag = (uint8_t)g1_on_curve(&p);
    else if (op == 3) *flag = (uint8_t)g1_torsion_free(&p);
    else return -2;
    return 0;
}
EXPORT int zo_g2_op(int op, const u64 *a, uint8_t ainf, const u64 *b, uint8_t binf, u64 *out, uint8_t *flag) {
    init();
    g2a p, q, r;
    load_g2(&p, a, ainf);
    if (op == 0) { g2_double(&r, &p); store_g2(out, flag, &r); }
    else if (op == 1) { load_g2(&q, b, binf); g2_add(&r, &p, &q); store_g2(out, flag, &r); }
    else if (op == 2) *flag = (uint8_t)g2_on_curve(&p);
    else if (op == 3) *flag = (uint8_t)g2_torsion_free(&p);
    else return -2;
    return 0;
}
/* constants the CUDA side also derives (for cross-checking the generated header) */
EXPORT void zo_constants(u64 *r2, u64 *frob6_c1, u64 *frob6_c2, u64 *frob12_c1) {
    init();
    memcpy(r2, &R2, 48);                 /* canonical integer 2^768 mod p */
    store_fp2(frob6_c1, &FROB6_C1); store_fp2(frob6_c2, &FROB6_C2); store_fp2(frob12_c1, &FROB12_C1);
}
