/*
 * zkpair.hpp -- C++17 host mirror of the reference crate's value-type API over the C ABI (zkpair.h).
 *
 * The reference (0xWOLAND/zkvm-pairings) is a Rust crate; this image has no Rust toolchain, so the
 * host side above the C ABI is written in C++ with the SAME names, argument meaning and error
 * behaviour as the crate's public surface, so that tests written against it read like the crate's
 * own tests (tests/cpp/test_reference_api.cpp follows src/fp.rs:474-610, src/fp12.rs:294-410,
 * src/g1.rs:215-301, src/g2.rs:264-443).  Header only; link with -lzkpair.
 *
 *   reference item                                      here
 *   Fp   (src/fp.rs:24, methods :145-456)               zkp::Fp      zero one is_zero from_raw_unchecked from_bytes
 *                                                                    to_bytes pow_vartime sqrt invert square + - * / neg
 *   Fp2  (src/fp2.rs:10, methods :118-314)              zkp::Fp2     new_ = ctor, frobenius_map conjugate
 *                                                                    mul_by_nonresidue square invert pow_vartime, * Fp
 *   Fp6  (src/fp6.rs:13, methods :70-310)               zkp::Fp6     mul_by_1 mul_by_01 mul_by_nonresidue frobenius_map ...
 *   Fp12 (src/fp12.rs:13, methods :75-210)              zkp::Fp12    mul_by_014 conjugate frobenius_map pow_vartime ...
 *   From<Fp> for Fp2/Fp6/Fp12 (replicating, src/fp2.rs:32-36, src/fp6.rs:19-27, src/fp12.rs:18-25)   T::from(...)
 *   G1Affine / G2Affine (src/g1.rs:7-62, src/g2.rs:8-69) zkp::G1Affine / G2Affine: identity generator is_identity
 *                                                                    is_valid is_on_curve is_torsion_free neg double_
 *                                                                    + - (affine law), * Fr limbs
 *   pairings::* (src/pairings.rs is EMPTY)    zkp::pairings::{pairing, miller_loop, multi_miller_loop,
 *                                                                    final_exponentiation, pairing_batch, multi_pairing_batch}
 *
 * Error behaviour follows the crate: invert()/sqrt() return an empty optional where the crate returns
 * None / Err(()); operator/ on a zero divisor throws zkp::Panic where the crate panics through
 * unwrap() (src/fp.rs:448-450, src/fp2.rs:211-213, src/fp6.rs:269-271, src/fp12.rs:113-115); is_valid()
 * returns the crate's error strings.  Every arithmetic call goes to the GPU through libzkpair.so (one
 * call per operation, exactly like the crate's per-operation zkVM precompile syscalls src/fp.rs:126,443);
 * there is no CPU arithmetic in this header and no fallback: without a CUDA device the first operation
 * throws zkp::Error(ZKP_ERR_NO_DEVICE).  Bulk work belongs in the *_batch functions.
 *
 * Deliberate deviations: frobenius_map is the true a -> a^p (the crate's Fp6 constants are
 * wrong, src/fp6.rs:147-173); point * scalar uses all 256 bits (the crate's G1 loop drops bit 0,
 * src/g1.rs:138-142).
 */
#ifndef ZKPAIR_HPP
#define ZKPAIR_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkpair.h"

namespace zkp {

/* a failed C-ABI call (code = ZKP_ERR_*) */
struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};
/* where the crate panics (division by zero through unwrap()) */
struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};

/* Process-wide engine context (the crate's types are plain values; the context is an implementation
 * detail of this mirror).  Engine::use(devices) may be called once before the first operation. */
class Engine {
  public:
    static zkp_ctx *ctx() { return instance().ctx_; }
    static void use(const std::vector<int> &devices) {
        Engine &e = instance(&devices);
        (void)e;
    }
    static void check(int32_t rc) {
        if (rc != ZKP_OK) throw Error(rc, zkp_last_error());
    }

  private:
    zkp_ctx *ctx_ = nullptr;
    explicit Engine(const std::vector<int> *devices) {
        int32_t rc = devices && !devices->empty() ? zkp_ctx_create(devices->data(), (int)devices->size(), &ctx_)
                                                  : zkp_ctx_create(nullptr, 0, &ctx_);
        check(rc);
    }
    ~Engine() { zkp_ctx_destroy(ctx_); }
    static Engine &instance(const std::vector<int> *devices = nullptr) {
        static Engine e(devices);
        return e;
    }
};

namespace detail {
/* one tower operation on one element; T/U are standard-layout runs of u64 limbs */
template <class R, class A, class B>
inline R op(int32_t code, const A &a, const B *b, uint8_t *status = nullptr) {
    R out;
    uint8_t st = 0;
    Engine::check(zkp_tower_op_batch(Engine::ctx(), code, reinterpret_cast<const uint64_t *>(&a),
                                     reinterpret_cast<const uint64_t *>(b), reinterpret_cast<uint64_t *>(&out), &st, 1));
    if (status) *status = st;
    return out;
}
template <class A> inline A op1(int32_t code, const A &a, uint8_t *status = nullptr) {
    return op<A, A, A>(code, a, nullptr, status);
}
template <class A> inline A op2(int32_t code, const A &a, const A &b) { return op<A, A, A>(code, a, &b); }
}  // namespace detail

/* ---- Fp (src/fp.rs) ------------------------------------------------------------------------ */
struct Fp {
    uint64_t v[6];  /* `Fp.0`: canonical little-endian limbs (src/fp.rs:24) */

    static Fp zero() { return Fp{{0, 0, 0, 0, 0, 0}}; }                          /* src/fp.rs:148 */
    static Fp one() { return Fp{{1, 0, 0, 0, 0, 0}}; }                           /* src/fp.rs:154 */
    static Fp from(uint64_t x) { return Fp{{x, 0, 0, 0, 0, 0}}; }                /* From<u64> src/fp.rs:43 */
    static Fp from_raw_unchecked(const std::array<uint64_t, 6> &l) {              /* src/fp.rs:257 */
        Fp r;
        std::memcpy(r.v, l.data(), 48);
        return r;
    }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5]) == 0; } /* src/fp.rs:159 */
    /* Result<Fp, ()>: empty when the 48 big-endian bytes encode a value >= p (src/fp.rs:165-191) */
    static std::optional<Fp> from_bytes(const std::array<uint8_t, 48> &bytes) {
        Fp r;
        uint8_t ok = 0;
        Engine::check(zkp_fp_from_bytes_batch(Engine::ctx(), bytes.data(), 1, r.v, &ok));
        if (!ok) return std::nullopt;
        return r;
    }
    std::array<uint8_t, 48> to_bytes() const {                                   /* src/fp.rs:195-207 */
        std::array<uint8_t, 48> out;
        Engine::check(zkp_fp_to_bytes_batch(Engine::ctx(), v, 1, out.data()));
        return out;
    }
    Fp pow_vartime(const std::array<uint64_t, 6> &by) const {                     /* src/fp.rs:264-276 */
        return detail::op<Fp, Fp, std::array<uint64_t, 6>>(ZKP_OP_FP_POW, *this, &by);
    }
    std::optional<Fp> sqrt() const {                                             /* src/fp.rs:280-300 */
        uint8_t st;
        Fp r = detail::op1(ZKP_OP_FP_SQRT, *this, &st);
        if (st & 2) return std::nullopt;
        return r;
    }
    std::optional<Fp> invert() const {                                           /* src/fp.rs:306-319 */
        uint8_t st;
        Fp r = detail::op1(ZKP_OP_FP_INV, *this, &st);
        if (st & 2) return std::nullopt;
        return r;
    }
    Fp square() const { return detail::op1(ZKP_OP_FP_SQR, *this); }             /* src/fp.rs:452-455 */
    Fp neg() const { return detail::op1(ZKP_OP_FP_NEG, *this); }                /* src/fp.rs:381-405 */
    Fp add(const Fp &r) const { return detail::op2(ZKP_OP_FP_ADD, *this, r); }  /* src/fp.rs:351-368 */
    Fp sub(const Fp &r) const { return detail::op2(ZKP_OP_FP_SUB, *this, r); }  /* src/fp.rs:407-411 */
    Fp mul(const Fp &r) const { return detail::op2(ZKP_OP_FP_MUL, *this, r); }  /* src/fp.rs:413-434 */
    Fp div(const Fp &r) const {                                                  /* src/fp.rs:448-450 */
        auto i = r.invert();
        if (!i) throw Panic("called `Option::unwrap()` on a `None` value (Fp division by zero)");
        return mul(*i);
    }
    std::string debug() const {                                                  /* fmt::Debug src/fp.rs:26-35 */
        static const char *hex = "0123456789abcdef";
        std::string s = "0x";
        for (uint8_t b : to_bytes()) {
            s.push_back(hex[b >> 4]);
            s.push_back(hex[b & 15]);
        }
        return s;
    }
};
inline bool operator==(const Fp &a, const Fp &b) { return std::memcmp(a.v, b.v, 48) == 0; } /* src/fp.rs:53-58 */
inline bool operator!=(const Fp &a, const Fp &b) { return !(a == b); }
inline Fp operator+(const Fp &a, const Fp &b) { return a.add(b); }
inline Fp operator-(const Fp &a, const Fp &b) { return a.sub(b); }
inline Fp operator*(const Fp &a, const Fp &b) { return a.mul(b); }
inline Fp operator/(const Fp &a, const Fp &b) { return a.div(b); }
inline Fp operator-(const Fp &a) { return a.neg(); }

namespace detail {
/* `T * Fp`: every Fp coefficient of T times rhs (src/fp2.rs:95-102, src/fp6.rs:369-380, src/fp12.rs:248-255) */
template <class T> inline T scale(const T &a, const Fp &rhs) {
    constexpr size_t n = sizeof(T) / sizeof(Fp);
    Fp b[n];
    for (size_t i = 0; i < n; ++i) b[i] = rhs;
    T out;
    Engine::check(zkp_tower_op_batch(Engine::ctx(), ZKP_OP_FP_MUL, reinterpret_cast<const uint64_t *>(&a), b[0].v,
                                     reinterpret_cast<uint64_t *>(&out), nullptr, n));
    return out;
}
template <class T> inline bool all_zero(const T &a) {
    const uint64_t *p = reinterpret_cast<const uint64_t *>(&a);
    uint64_t acc = 0;
    for (size_t i = 0; i < sizeof(T) / 8; ++i) acc |= p[i];
    return acc == 0;
}
template <class T> inline std::optional<T> invert(int32_t code, const T &a) {
    uint8_t st;
    T r = op1(code, a, &st);
    if (st & 2) return std::nullopt;
    return r;
}
template <class T> inline T div(int32_t inv_code, int32_t mul_code, const T &a, const T &b, const char *what) {
    auto i = invert(inv_code, b);
    if (!i) throw Panic(std::string("called `Option::unwrap()` on a `None` value (") + what + " division by zero)");
    return op2(mul_code, a, *i);
}
}  // namespace detail

/* ---- Fp2 (src/fp2.rs) ---------------------------------------------------------------------- */
struct Fp2 {
    Fp c0, c1;
    static Fp2 zero() { return Fp2{Fp::zero(), Fp::zero()}; }                    /* src/fp2.rs:121 */
    static Fp2 one() { return Fp2{Fp::one(), Fp::zero()}; }                      /* src/fp2.rs:127 */
    static Fp2 new_(const Fp &c0, const Fp &c1) { return Fp2{c0, c1}; }          /* src/fp2.rs:131 */
    static Fp2 from(const Fp &f) { return Fp2{f, f}; }                           /* src/fp2.rs:32-36 (replicates) */
    bool is_zero() const { return detail::all_zero(*this); }                     /* src/fp2.rs:136 */
    Fp2 frobenius_map() const { return conjugate(); }                            /* src/fp2.rs:147-153 */
    Fp2 conjugate() const { return detail::op1(ZKP_OP_FP2_CONJ, *this); }       /* src/fp2.rs:155-157 */
    Fp2 mul_by_nonresidue() const { return detail::op1(ZKP_OP_FP2_MUL_NR, *this); } /* src/fp2.rs:161-168 */
    Fp2 square() const { return detail::op1(ZKP_OP_FP2_SQR, *this); }           /* src/fp2.rs:171-189 */
    Fp2 mul(const Fp2 &r) const { return detail::op2(ZKP_OP_FP2_MUL, *this, r); } /* src/fp2.rs:192-209 */
    Fp2 add(const Fp2 &r) const { return detail::op2(ZKP_OP_FP2_ADD, *this, r); } /* src/fp2.rs:216 */
    Fp2 sub(const Fp2 &r) const { return detail::op2(ZKP_OP_FP2_SUB, *this, r); } /* src/fp2.rs:221 */
    Fp2 neg() const { return detail::op1(ZKP_OP_FP2_NEG, *this); }              /* src/fp2.rs:226 */
    std::optional<Fp2> invert() const { return detail::invert(ZKP_OP_FP2_INV, *this); } /* src/fp2.rs:278-296 */
    Fp2 div(const Fp2 &r) const { return detail::div(ZKP_OP_FP2_INV, ZKP_OP_FP2_MUL, *this, r, "Fp2"); } /* :211-213 */
    Fp2 pow_vartime(const std::array<uint64_t, 6> &by) const {                    /* src/fp2.rs:301-313 */
        return detail::op<Fp2, Fp2, std::array<uint64_t, 6>>(ZKP_OP_FP2_POW, *this, &by);
    }
};

/* ---- Fp6 (src/fp6.rs) ---------------------------------------------------------------------- */
struct Fp6 {
    Fp2 c0, c1, c2;
    static Fp6 new_(const Fp2 &c0, const Fp2 &c1, const Fp2 &c2) { return Fp6{c0, c1, c2}; } /* src/fp6.rs:72 */
    static Fp6 zero() { return Fp6{Fp2::zero(), Fp2::zero(), Fp2::zero()}; }     /* src/fp6.rs:77 */
    static Fp6 one() { return Fp6{Fp2::one(), Fp2::zero(), Fp2::zero()}; }       /* src/fp6.rs:86 */
    static Fp6 from(const Fp &f) { return Fp6{Fp2::from(f), Fp2::from(f), Fp2::from(f)}; } /* src/fp6.rs:19-27 */
    static Fp6 from(const Fp2 &f) { return Fp6{f, Fp2::zero(), Fp2::zero()}; }   /* src/fp6.rs:29-37 */
    bool is_zero() const { return detail::all_zero(*this); }                     /* src/fp6.rs:179 */
    Fp6 mul_by_1(const Fp2 &c1_) const {                                         /* src/fp6.rs:102-108 */
        return detail::op<Fp6, Fp6, Fp2>(ZKP_OP_FP6_MUL_BY_1, *this, &c1_);
    }
    Fp6 mul_by_01(const Fp2 &c0_, const Fp2 &c1_) const {                        /* src/fp6.rs:110-125 */
        Fp2 b[2] = {c0_, c1_};
        return detail::op<Fp6, Fp6, Fp2>(ZKP_OP_FP6_MUL_BY_01, *this, b);
    }
    Fp6 mul_by_nonresidue() const { return detail::op1(ZKP_OP_FP6_MUL_NR, *this); } /* src/fp6.rs:128-139 */
    Fp6 frobenius_map() const { return detail::op1(ZKP_OP_FP6_FROB, *this); }   /* TRUE a^p; cf. src/fp6.rs:142-176 */
    Fp6 mul_interleaved(const Fp6 &r) const { return mul(r); }                   /* src/fp6.rs:188-267 */
    Fp6 mul(const Fp6 &r) const { return detail::op2(ZKP_OP_FP6_MUL, *this, r); }
    Fp6 square() const { return detail::op1(ZKP_OP_FP6_SQR, *this); }           /* src/fp6.rs:274-288 */
    Fp6 add(const Fp6 &r) const { return detail::op2(ZKP_OP_FP6_ADD, *this, r); }
    Fp6 sub(const Fp6 &r) const { return detail::op2(ZKP_OP_FP6_SUB, *this, r); }
    Fp6 neg() const { return detail::op1(ZKP_OP_FP6_NEG, *this); }
    std::optional<Fp6> invert() const { return detail::invert(ZKP_OP_FP6_INV, *this); } /* src/fp6.rs:291-309 */
    Fp6 div(const Fp6 &r) const { return detail::div(ZKP_OP_FP6_INV, ZKP_OP_FP6_MUL, *this, r, "Fp6"); } /* :269-271 */
};

/* ---- Fp12 (src/fp12.rs) -------------------------------------------------------------------- */
struct Fp12 {
    Fp6 c0, c1;
    static Fp12 new_(const Fp6 &c0, const Fp6 &c1) { return Fp12{c0, c1}; }      /* src/fp12.rs:77 */
    static Fp12 zero() { return Fp12{Fp6::zero(), Fp6::zero()}; }                /* src/fp12.rs:82 */
    static Fp12 one() { return Fp12{Fp6::one(), Fp6::zero()}; }                  /* src/fp12.rs:87 */
    static Fp12 from(const Fp &f) { return Fp12{Fp6::from(f), Fp6::from(f)}; }   /* src/fp12.rs:18-25 */
    static Fp12 from(const Fp2 &f) { return Fp12{Fp6::from(f), Fp6::zero()}; }   /* src/fp12.rs:27-34 */
    static Fp12 from(const Fp6 &f) { return Fp12{f, Fp6::zero()}; }              /* src/fp12.rs:36-43 */
    bool is_zero() const { return detail::all_zero(*this); }                     /* src/fp12.rs:118 */
    Fp12 mul_by_014(const Fp2 &c0_, const Fp2 &c1_, const Fp2 &c4_) const {      /* src/fp12.rs:99-111 */
        Fp2 b[3] = {c0_, c1_, c4_};
        return detail::op<Fp12, Fp12, Fp2>(ZKP_OP_FP12_MUL_BY_014, *this, b);
    }
    Fp12 conjugate() const { return detail::op1(ZKP_OP_FP12_CONJ, *this); }     /* src/fp12.rs:123-125 */
    Fp12 pow_vartime(const std::array<uint64_t, 6> &by) const {                   /* src/fp12.rs:127-139 */
        return detail::op<Fp12, Fp12, std::array<uint64_t, 6>>(ZKP_OP_FP12_POW, *this, &by);
    }
    Fp12 frobenius_map() const { return detail::op1(ZKP_OP_FP12_FROB, *this); } /* TRUE a^p; cf. src/fp12.rs:143-170 */
    Fp12 square() const { return detail::op1(ZKP_OP_FP12_SQR, *this); }         /* src/fp12.rs:173-184 */
    Fp12 mul(const Fp12 &r) const { return detail::op2(ZKP_OP_FP12_MUL, *this, r); } /* src/fp12.rs:193-210 */
    Fp12 add(const Fp12 &r) const { return detail::op2(ZKP_OP_FP12_ADD, *this, r); }
    Fp12 sub(const Fp12 &r) const { return detail::op2(ZKP_OP_FP12_SUB, *this, r); }
    Fp12 neg() const { return detail::op1(ZKP_OP_FP12_NEG, *this); }
    std::optional<Fp12> invert() const { return detail::invert(ZKP_OP_FP12_INV, *this); } /* src/fp12.rs:186-190 */
    Fp12 div(const Fp12 &r) const { return detail::div(ZKP_OP_FP12_INV, ZKP_OP_FP12_MUL, *this, r, "Fp12"); } /* :113-115 */
};
using Gt = Fp12;

static_assert(sizeof(Fp) == 48 && sizeof(Fp2) == 96 && sizeof(Fp6) == 288 && sizeof(Fp12) == 576,
              "value types must be plain runs of limbs: they are passed to the C ABI as they are");

#define ZKP_HPP_OPERATORS(T)                                                                        \
    inline bool operator==(const T &a, const T &b) { return std::memcmp(&a, &b, sizeof(T)) == 0; }  \
    inline bool operator!=(const T &a, const T &b) { return !(a == b); }                            \
    inline T operator+(const T &a, const T &b) { return a.add(b); }                                 \
    inline T operator-(const T &a, const T &b) { return a.sub(b); }                                 \
    inline T operator*(const T &a, const T &b) { return a.mul(b); }                                 \
    inline T operator/(const T &a, const T &b) { return a.div(b); }                                 \
    inline T operator-(const T &a) { return a.neg(); }                                              \
    inline T operator*(const T &a, const Fp &b) { return detail::scale(a, b); }
ZKP_HPP_OPERATORS(Fp2)
ZKP_HPP_OPERATORS(Fp6)
ZKP_HPP_OPERATORS(Fp12)
#undef ZKP_HPP_OPERATORS

/* The limbs of an `Fr` (src/fr.rs): four little-endian u64, the scalar of `&G1Affine * &Fr`. */
using FrLimbs = std::array<uint64_t, 4>;
inline FrLimbs fr_from(uint64_t x) { return FrLimbs{x, 0, 0, 0}; }

/* `Result<(), String>` of is_valid (src/g1.rs:49-62, src/g2.rs:57-69) */
struct Validity {
    std::string err; /* empty = Ok(()) */
    bool is_ok() const { return err.empty(); }
    bool is_err() const { return !err.empty(); }
};

namespace detail {
inline Validity validity(uint8_t st) {
    if (st == ZKP_POINT_NOT_ON_CURVE) return Validity{"Point is not on curve"};
    if (st == ZKP_POINT_NOT_TORSION_FREE) return Validity{"Point is not torsion free"};
    return Validity{};
}
}  // namespace detail

/* ---- G1Affine (src/g1.rs) ------------------------------------------------------------------ */
struct G1Affine {
    Fp x, y;
    bool is_infinity;
    static G1Affine new_(const Fp &x, const Fp &y, bool inf) { return G1Affine{x, y, inf}; } /* src/g1.rs:22 */
    static G1Affine identity() { return G1Affine{Fp::zero(), Fp::one(), true}; }             /* src/g1.rs:25-31 */
    static G1Affine generator() {                                                            /* src/g1.rs:41-47, src/common.rs:92-109 */
        return G1Affine{
            Fp{{0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL, 0xc3688c4f9774b905ULL,
                0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL}},
            Fp{{0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL, 0xfcf5e095d5d00af6ULL,
                0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL}},
            false};
    }
    bool is_identity() const { return is_infinity; }                                          /* src/g1.rs:33 */
    bool is_zero() const { return x.is_zero() && y.is_zero(); }                               /* src/g1.rs:37 */
    uint8_t status() const {
        uint8_t inf = is_infinity, st = 0;
        Engine::check(zkp_g1_check_batch(Engine::ctx(), x.v, &inf, 1, &st));
        return st;
    }
    Validity is_valid() const { return detail::validity(status()); }                          /* src/g1.rs:49-62 */
    bool is_on_curve() const {                                                                /* src/g1.rs:95-101 */
        G1Affine t = *this;
        t.is_infinity = false;
        return t.status() != ZKP_POINT_NOT_ON_CURVE;
    }
    bool is_torsion_free() const {                                                            /* src/g1.rs:109-115 */
        G1Affine t = *this;
        t.is_infinity = false;
        return t.status() == ZKP_POINT_OK;
    }
    G1Affine neg() const { return G1Affine{x, y.neg(), is_infinity}; }                        /* src/g1.rs:118-128 */
    G1Affine mul(const FrLimbs &k) const {                                                    /* src/g1.rs:130-153 */
        G1Affine out;
        uint8_t inf = is_infinity, oinf = 0;
        Engine::check(zkp_g1_mul_batch(Engine::ctx(), x.v, &inf, k.data(), 1, out.x.v, &oinf));
        out.is_infinity = oinf != 0;
        return out;
    }
    /* `&G1Affine + &G1Affine` (src/g1.rs:155-187); P + (-P) divides by zero in the crate and panics (:177) */
    G1Affine add(const G1Affine &o) const {
        G1Affine out;
        uint8_t ia = is_infinity, ib = o.is_infinity, flag = 0;
        Engine::check(zkp_g1_add_batch(Engine::ctx(), x.v, &ia, o.x.v, &ib, 1, out.x.v, &flag));
        if (flag & 2) throw Panic("called `Option::unwrap()` on a `None` value (affine addition: division by zero)");
        out.is_infinity = (flag & 1) != 0;
        return out;
    }
    G1Affine double_() const { return is_infinity ? identity() : add(*this); }                /* `double`, src/g1.rs:74-91 */
};
/* points compare by coordinates only: is_infinity is ignored (src/g1.rs:13-17) */
inline bool operator==(const G1Affine &a, const G1Affine &b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const G1Affine &a, const G1Affine &b) { return !(a == b); }
inline G1Affine operator-(const G1Affine &a) { return a.neg(); }
inline G1Affine operator*(const G1Affine &a, const FrLimbs &k) { return a.mul(k); }
inline G1Affine operator+(const G1Affine &a, const G1Affine &b) { return a.add(b); }
inline G1Affine operator-(const G1Affine &a, const G1Affine &b) { return a.add(b.neg()); }            /* src/g1.rs:189-196 */

/* ---- G2Affine (src/g2.rs) ------------------------------------------------------------------ */
struct G2Affine {
    Fp2 x, y;
    bool is_infinity;
    static G2Affine new_(const Fp2 &x, const Fp2 &y, bool inf) { return G2Affine{x, y, inf}; } /* src/g2.rs:23 */
    static G2Affine identity() { return G2Affine{Fp2::zero(), Fp2::one(), true}; }            /* src/g2.rs:27-33 */
    static G2Affine generator() {                                                             /* src/g2.rs:43-55, src/common.rs:110-145 */
        return G2Affine{
            Fp2{Fp{{0xd48056c8c121bdb8ULL, 0x0bac0326a805bbefULL, 0xb4510b647ae3d177ULL, 0xc6e47ad4fa403b02ULL,
                    0x260805272dc51051ULL, 0x024aa2b2f08f0a91ULL}},
                Fp{{0xe5ac7d055d042b7eULL, 0x334cf11213945d57ULL, 0xb5da61bbdc7f5049ULL, 0x596bd0d09920b61aULL,
                    0x7dacd3a088274f65ULL, 0x13e02b6052719f60ULL}}},
            Fp2{Fp{{0xe193548608b82801ULL, 0x923ac9cc3baca289ULL, 0x6d429a695160d12cULL, 0xadfd9baa8cbdd3a7ULL,
                    0x8cc9cdc6da2e351aULL, 0x0ce5d527727d6e11ULL}},
                Fp{{0xaaa9075ff05f79beULL, 0x3f370d275cec1da1ULL, 0x267492ab572e99abULL, 0xcb3e287e85a763afULL,
                    0x32acd2b02bc28b99ULL, 0x0606c4a02ea734ccULL}}},
            false};
    }
    bool is_identity() const { return is_infinity; }                                          /* src/g2.rs:35 */
    bool is_zero() const { return x.is_zero() && y.is_zero(); }                               /* src/g2.rs:39 */
    uint8_t status() const {
        uint8_t inf = is_infinity, st = 0;
        Engine::check(zkp_g2_check_batch(Engine::ctx(), x.c0.v, &inf, 1, &st));
        return st;
    }
    Validity is_valid() const { return detail::validity(status()); }                          /* src/g2.rs:57-69 */
    bool is_on_curve() const {                                                                /* src/g2.rs:109-120 */
        G2Affine t = *this;
        t.is_infinity = false;
        return t.status() != ZKP_POINT_NOT_ON_CURVE;
    }
    /* false for a point that is not on the curve (the crate runs its affine formulas on such a point anyway) */
    bool is_torsion_free() const {                                                            /* src/g2.rs:166-170 */
        G2Affine t = *this;
        t.is_infinity = false;
        return t.status() == ZKP_POINT_OK;
    }
    G2Affine neg() const { return G2Affine{x, y.neg(), is_infinity}; }                        /* src/g2.rs:173-183 */
    G2Affine mul(const FrLimbs &k) const {                                                    /* src/g2.rs:185-208 */
        G2Affine out;
        uint8_t inf = is_infinity, oinf = 0;
        Engine::check(zkp_g2_mul_batch(Engine::ctx(), x.c0.v, &inf, k.data(), 1, out.x.c0.v, &oinf));
        out.is_infinity = oinf != 0;
        return out;
    }
    /* `&G2Affine + &G2Affine` (src/g2.rs:210-242); P + (-P) panics in the crate (:232) */
    G2Affine add(const G2Affine &o) const {
        G2Affine out;
        uint8_t ia = is_infinity, ib = o.is_infinity, flag = 0;
        Engine::check(zkp_g2_add_batch(Engine::ctx(), x.c0.v, &ia, o.x.c0.v, &ib, 1, out.x.c0.v, &flag));
        if (flag & 2) throw Panic("called `Option::unwrap()` on a `None` value (affine addition: division by zero)");
        out.is_infinity = (flag & 1) != 0;
        return out;
    }
    G2Affine double_() const { return is_infinity ? identity() : add(*this); }                /* `double`, src/g2.rs:81-105 */
};
inline bool operator==(const G2Affine &a, const G2Affine &b) { return a.x == b.x && a.y == b.y; } /* src/g2.rs:14-18 */
inline bool operator!=(const G2Affine &a, const G2Affine &b) { return !(a == b); }
inline G2Affine operator-(const G2Affine &a) { return a.neg(); }
inline G2Affine operator*(const G2Affine &a, const FrLimbs &k) { return a.mul(k); }
inline G2Affine operator+(const G2Affine &a, const G2Affine &b) { return a.add(b); }
inline G2Affine operator-(const G2Affine &a, const G2Affine &b) { return a.add(b.neg()); }            /* src/g2.rs:244-251 */

static_assert(offsetof(G1Affine, y) == 48 && offsetof(G2Affine, y) == 96, "x|y must be contiguous for the C ABI");

/* ---- pairings (the module the reference declares but leaves empty: src/lib.rs:12, src/pairings.rs) ---- */
namespace pairings {

namespace detail {
struct Flat {
    std::vector<uint64_t> g1, g2;
    std::vector<uint8_t> i1, i2;
};
inline Flat flatten(const G1Affine *ps, const G2Affine *qs, size_t n) {
    Flat f;
    f.g1.resize(12 * n);
    f.g2.resize(24 * n);
    f.i1.resize(n);
    f.i2.resize(n);
    for (size_t i = 0; i < n; ++i) {
        std::memcpy(&f.g1[12 * i], ps[i].x.v, 96);
        std::memcpy(&f.g2[24 * i], qs[i].x.c0.v, 192);
        f.i1[i] = ps[i].is_infinity;
        f.i2[i] = qs[i].is_infinity;
    }
    return f;
}
}  // namespace detail

/* n independent pairings e(P_i, Q_i) -- the batch entry point the north star adds */
inline std::vector<Gt> pairing_batch(const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs) {
    if (ps.size() != qs.size()) throw Error(ZKP_ERR_INVALID_ARG, "pairing_batch: length mismatch");
    auto f = detail::flatten(ps.data(), qs.data(), ps.size());
    std::vector<Gt> out(ps.size());
    Engine::check(zkp_pairing_batch(Engine::ctx(), f.g1.data(), f.i1.data(), f.g2.data(), f.i2.data(), ps.size(),
                                    reinterpret_cast<uint64_t *>(out.data())));
    return out;
}
inline std::vector<Fp12> miller_loop_batch(const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs) {
    if (ps.size() != qs.size()) throw Error(ZKP_ERR_INVALID_ARG, "miller_loop_batch: length mismatch");
    auto f = detail::flatten(ps.data(), qs.data(), ps.size());
    std::vector<Fp12> out(ps.size());
    Engine::check(zkp_miller_loop_batch(Engine::ctx(), f.g1.data(), f.i1.data(), f.g2.data(), f.i2.data(), ps.size(),
                                        reinterpret_cast<uint64_t *>(out.data())));
    return out;
}
inline std::vector<Gt> final_exponentiation_batch(const std::vector<Fp12> &fs) {
    std::vector<Gt> out(fs.size());
    Engine::check(zkp_final_exp_batch(Engine::ctx(), reinterpret_cast<const uint64_t *>(fs.data()), fs.size(),
                                      reinterpret_cast<uint64_t *>(out.data())));
    return out;
}
/* n_checks products of k pairs (check-major) with one shared final exponentiation each; second = "is one" */
inline std::pair<std::vector<Gt>, std::vector<bool>> multi_pairing_batch(const std::vector<G1Affine> &ps,
                                                                         const std::vector<G2Affine> &qs, size_t k) {
    if (k == 0 || ps.size() != qs.size() || ps.size() % k) throw Error(ZKP_ERR_INVALID_ARG, "multi_pairing_batch: bad shape");
    size_t n = ps.size() / k;
    auto f = detail::flatten(ps.data(), qs.data(), ps.size());
    std::vector<Gt> out(n);
    std::vector<uint8_t> one(n);
    Engine::check(zkp_multi_pairing_batch(Engine::ctx(), f.g1.data(), f.i1.data(), f.g2.data(), f.i2.data(), n, (int32_t)k,
                                          reinterpret_cast<uint64_t *>(out.data()), one.data()));
    return {out, std::vector<bool>(one.begin(), one.end())};
}
/* zkcrypto-lineage single-value functions */
inline Gt pairing(const G1Affine &p, const G2Affine &q) { return pairing_batch({p}, {q})[0]; }
inline Fp12 miller_loop(const G1Affine &p, const G2Affine &q) { return miller_loop_batch({p}, {q})[0]; }
inline Gt final_exponentiation(const Fp12 &f) { return final_exponentiation_batch({f})[0]; }
/* one Miller loop over several pairs sharing the accumulator (<= ZKP_MAX_PAIRS_PER_CHECK pairs) */
inline Fp12 multi_miller_loop(const std::vector<std::pair<G1Affine, G2Affine>> &terms) {
    std::vector<G1Affine> ps;
    std::vector<G2Affine> qs;
    for (auto &t : terms) {
        ps.push_back(t.first);
        qs.push_back(t.second);
    }
    if (terms.empty()) return Fp12::one();
    auto f = detail::flatten(ps.data(), qs.data(), ps.size());
    Fp12 out;
    Engine::check(zkp_multi_miller_loop_batch(Engine::ctx(), f.g1.data(), f.i1.data(), f.g2.data(), f.i2.data(), 1,
                                              (int32_t)ps.size(), reinterpret_cast<uint64_t *>(&out)));
    return out;
}
/* prod_i e(P_i, Q_i) over ANY number of pairs: per-GPU partial products, 576-byte gather, one final exponentiation */
inline Gt multi_miller_product(const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs) {
    if (ps.size() != qs.size()) throw Error(ZKP_ERR_INVALID_ARG, "multi_miller_product: length mismatch");
    auto f = detail::flatten(ps.data(), qs.data(), ps.size());
    Gt out;
    Engine::check(zkp_multi_miller_product(Engine::ctx(), f.g1.data(), f.i1.data(), f.g2.data(), f.i2.data(), ps.size(),
                                           nullptr, reinterpret_cast<uint64_t *>(&out)));
    return out;
}

}  // namespace pairings
}  // namespace zkp

#endif /* ZKPAIR_HPP */
