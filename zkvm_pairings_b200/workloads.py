"""Synthetic workloads of BASELINE.json's configs, built with the engine itself (GPU) and Python integers.

Nothing here touches the oracle: tests and bench.py use these generators for the INPUTS and the expected
verdicts, and compare Gt values with the oracle separately.

Config 3 ("Groth16-style 4-pair multi-pairing product checks, shared final exponentiation per check"):
every check is the verification equation of a Groth16-shaped verifier with a fixed verifying key
(beta, gamma, delta in G2) and per-proof points A, C, L, X in G1 and B in G2,

    e(A, B) * e(X, beta G2) * e(C, -delta G2) * e(L, -gamma G2) == 1,

with A = a G1, B = b G2, C = c G1, L = l G1 from 64-bit seeded scalars and X = s G1 where
s = (c delta + l gamma - a b) / beta mod r makes the product one.  A seeded 1 % of the checks is corrupted
(s + 1), so the expected verdict vector is known without running a pairing.
"""
from __future__ import annotations

import numpy as np

# BLS12-381 group order and generators (the curve's standard constants; the reference carries them in
# src/common.rs:69-157)
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1_GEN = (
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
)
G2_GEN = (
    (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
     0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
    (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
     0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
)
VK_SEED = 0x6716_0B20_0


def _limbs(v: int, n: int):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def _fp_row(vals):
    return np.array([x for v in vals for x in _limbs(v, 6)], dtype=np.uint64)


def splitmix64_at(seed: int, idx: np.ndarray) -> np.ndarray:
    """Output number idx+1 of the SplitMix64 stream seeded with `seed` (vectorised; the scalars
    zkp_gen_points uses, csrc/ops.cuh splitmix64_at)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def gen_scalars(seed: int, first: int, n: int):
    """(a_i, b_i) of zkp_gen_points(seed, first, n): zero mapped to one."""
    i = np.arange(first, first + n, dtype=np.uint64)
    a = splitmix64_at(seed, np.uint64(2) * i)
    b = splitmix64_at(seed, np.uint64(2) * i + np.uint64(1))
    a[a == 0] = 1
    b[b == 0] = 1
    return a, b


def scalar_matrix(ks) -> np.ndarray:
    """Python ints -> (n, 4) little-endian u64 limbs (the limbs of an Fr)."""
    return np.array([_limbs(int(k), 4) for k in ks], dtype=np.uint64).reshape(-1, 4)


def groth16_checks(engine, n_checks: int, seed: int = 0x6716, corrupt_every: int = 100, corrupt_at: int = 7):
    """BASELINE config 3 inputs.  Returns a dict:
         g1 (n*4, 12)   A, X, C, L per check            g2 (n*4, 24)  B, beta, -delta, -gamma per check
         g2_var (n, 24) B only (prepared variant)        fixed (3, 24) beta G2, -delta G2, -gamma G2
         expect_one (n,) bool                            pairs_per_check = 4, prepared_pairs = 3
    """
    n = int(n_checks)
    vk = splitmix64_at(VK_SEED, np.arange(12, dtype=np.uint64))
    beta, gamma, delta = (int.from_bytes(vk[4 * j:4 * j + 4].tobytes(), "little") % R_ORDER or 1 for j in range(3))
    g2gen = np.tile(_fp_row([G2_GEN[0][0], G2_GEN[0][1], G2_GEN[1][0], G2_GEN[1][1]]), (3, 1))
    fixed, finf = engine.g2_mul_batch(g2gen, scalar_matrix([beta, R_ORDER - delta, R_ORDER - gamma]))
    assert not finf.any()
    # per-proof points from the seeded generator: (A, B), C, L
    gA, _, gB, _ = engine.gen_points(seed, 0, n)
    gC, _, _, _ = engine.gen_points(seed + 1, 0, n)
    gL, _, _, _ = engine.gen_points(seed + 2, 0, n)
    a, b = gen_scalars(seed, 0, n)
    c, _ = gen_scalars(seed + 1, 0, n)
    l, _ = gen_scalars(seed + 2, 0, n)
    beta_inv = pow(beta, -1, R_ORDER)
    corrupt = np.zeros(n, dtype=bool)
    if corrupt_every:
        corrupt[corrupt_at % corrupt_every::corrupt_every] = True
    s = np.empty((n, 4), dtype=np.uint64)
    for i in range(n):
        v = ((int(c[i]) * delta + int(l[i]) * gamma - int(a[i]) * int(b[i])) * beta_inv + (1 if corrupt[i] else 0)) % R_ORDER
        s[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        s[i, 1] = (v >> 64) & 0xFFFFFFFFFFFFFFFF
        s[i, 2] = (v >> 128) & 0xFFFFFFFFFFFFFFFF
        s[i, 3] = v >> 192
    g1gen = np.broadcast_to(_fp_row([G1_GEN[0], G1_GEN[1]]), (n, 12))
    gX, xinf = engine.g1_mul_batch(np.ascontiguousarray(g1gen), s)
    assert not xinf.any()
    g1 = np.empty((n, 4, 12), dtype=np.uint64)
    g1[:, 0], g1[:, 1], g1[:, 2], g1[:, 3] = gA, gX, gC, gL
    g2 = np.empty((n, 4, 24), dtype=np.uint64)
    g2[:, 0] = gB
    g2[:, 1:] = fixed[None]
    return {"g1": g1.reshape(-1, 12), "g2": g2.reshape(-1, 24), "g2_var": np.ascontiguousarray(gB), "fixed": fixed,
            "expect_one": ~corrupt, "pairs_per_check": 4, "prepared_pairs": 3}
