#!/usr/bin/env python3
"""Regenerates tests/golden/*.json.  Run in the dev container only (needs /root/reference).

reference_kats.json : every known-answer vector / fixed operand the REFERENCE's own tests hold
                      for the hot path, extracted textually from /root/reference/src/*.rs (the
                      crate cannot be compiled here: no rustc, un-vendored sp1 git dependency).
pairing_vectors.json: vectors for the path the reference leaves empty (src/pairings.rs, 0 bytes),
                      produced by oracle/pyref.py and pinned to the published BLS12-381
                      e(G1,G2) value.
"""
import json, os, re, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as o  # noqa: E402

ARR = re.compile(r"from_raw_unchecked\(\[\s*((?:0x[0-9a-fA-F_]+\s*,\s*){5}0x[0-9a-fA-F_]+\s*,?\s*)\]\)")


def arrays(fname, lo, hi):
    """All 6-limb from_raw_unchecked([...]) literals in lines lo..hi (1-based, inclusive)."""
    with open(os.path.join(REF, fname)) as f:
        text = "".join(f.readlines()[lo - 1:hi])
    out = []
    for m in ARR.finditer(text):
        limbs = [int(t.replace("_", ""), 16) for t in re.findall(r"0x[0-9a-fA-F_]+", m.group(1))]
        assert len(limbs) == 6
        out.append(["0x%016x" % l for l in limbs])
    return out


def main():
    kats = {"_source": "extracted from /root/reference/src by tests/golden/make_golden.py",
            "_format": "each Fp = six little-endian u64 limbs, canonical (src/fp.rs:24)"}
    # src/fp.rs:577-588
    kats["fp_sqrt"] = {"cite": "src/fp.rs:577-588", "input_u64": 300855555557,
                       "sqrt_be_hex": "025e51146a92917731d9d66d63f8c24ed8cae114e7c9d188e3eaa1e79bb19769f5877f9443e03723d9ed1eebbf92df98",
                       "non_residue_u64": 72057594037927816}
    # src/g1.rs:263-301 : generator and 2*generator ; :303-341 a second point and its double
    g = arrays("g1.rs", 262, 342)
    assert len(g) == 8
    kats["g1_double"] = {"cite": "src/g1.rs:263-341",
                         "cases": [{"p": g[0:2], "p2": g[2:4], "asserted_by_reference": True},
                                   {"p": g[4:6], "p2": g[6:8], "asserted_by_reference": False}]}
    # src/g2.rs:349-397 : 2*G2 generator
    g = arrays("g2.rs", 348, 398)
    assert len(g) == 4
    kats["g2_double_generator"] = {"cite": "src/g2.rs:349-397", "p2": g}
    # src/g2.rs:401-443 : a curve point that is NOT torsion free
    g = arrays("g2.rs", 400, 444)
    assert len(g) == 4
    kats["g2_not_torsion_free"] = {"cite": "src/g2.rs:401-443", "p": g}
    # src/fp6.rs:562-757 fixed operands a,b,c (6 Fp each)
    g = arrays("fp6.rs", 561, 735)
    assert len(g) == 18, len(g)
    kats["fp6_abc"] = {"cite": "src/fp6.rs:562-757", "a": g[0:6], "b": g[6:12], "c": g[12:18]}
    # src/fp12.rs:414-799 fixed operands a,b,c (12 Fp each)
    g = arrays("fp12.rs", 413, 765)
    assert len(g) == 36, len(g)
    kats["fp12_abc"] = {"cite": "src/fp12.rs:414-799", "a": g[0:12], "b": g[12:24], "c": g[24:36]}
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    # ---- pairing vectors (oracle-generated; reference has none) ----
    def h(a):
        return ["%096x" % c for c in o.fp12_flatten(a)]

    def g1h(p):
        return {"x": "%096x" % p[0], "y": "%096x" % p[1], "inf": p[2]}

    def g2h(q):
        return {"x": ["%096x" % q[0][0], "%096x" % q[0][1]], "y": ["%096x" % q[1][0], "%096x" % q[1][1]], "inf": q[2]}

    G1, G2 = o.G1_GENERATOR, o.G2_GENERATOR
    vec = {"_source": "oracle/pyref.py via tests/golden/make_golden.py; e(G1,G2) equals the published BLS12-381 Gt generator (SURVEY 9.4)",
           "_format": "Fp = 96-hex-digit big-endian canonical; Fp12 = 12 Fp in order c0.c0.c0, c0.c0.c1, ..., c1.c2.c1"}
    ml = o.miller_loop(G1, G2)
    e = o.final_exponentiation(ml)
    assert o.fp12_sha256(e) == "06fa588b89fdfb034dbc1c163ecb3dfac228f552b643c7294cc5f2c4dc170b84"
    assert o.fp12_sha256(ml) == "eceb6467936a62ed011881c3efceb3b9f05b6017afd264fa0caeebf4f8437115"
    vec["generators"] = {"g1": g1h(G1), "g2": g2h(G2), "miller_loop": h(ml), "pairing": h(e),
                         "miller_sha256": o.fp12_sha256(ml), "pairing_sha256": o.fp12_sha256(e)}
    cases = []
    for (a, b) in [(1, 2), (2, 1), (6, 11), (5, 7), (0xdeadbeef, 0xfeedface12345),
                   (o.R_ORDER - 1, 3), (123456789123456789, o.R_ORDER - 2)]:
        p, q = o.g1_mul(G1, a), o.g2_mul(G2, b)
        m = o.miller_loop(p, q)
        cases.append({"a": hex(a), "b": hex(b), "g1": g1h(p), "g2": g2h(q), "miller_loop": h(m),
                      "pairing": h(o.final_exponentiation(m))})
    # infinity handling
    cases.append({"a": "0x0", "b": "0x1", "g1": g1h(o.G1_IDENTITY), "g2": g2h(G2),
                  "miller_loop": h(o.miller_loop(o.G1_IDENTITY, G2)), "pairing": h(o.pairing(o.G1_IDENTITY, G2))})
    cases.append({"a": "0x1", "b": "0x0", "g1": g1h(G1), "g2": g2h(o.G2_IDENTITY),
                  "miller_loop": h(o.miller_loop(G1, o.G2_IDENTITY)), "pairing": h(o.pairing(G1, o.G2_IDENTITY))})
    vec["pairings"] = cases
    # 4-pair product checks (Groth16 shape): valid => one, corrupted => not one
    A, B = o.g1_mul(G1, 35), o.g2_mul(G2, 6)              # e(A,B) = e^(210)
    al, be = o.g1_mul(G1, 10), o.g2_mul(G2, 7)            # 70
    C, de = o.g1_mul(G1, 20), o.g2_mul(G2, 5)             # 100
    L, ga = o.g1_mul(G1, 8), o.g2_mul(G2, 5)              # 40
    good = [(A, B), (o.g1_neg(al), be), (o.g1_neg(C), de), (o.g1_neg(L), ga)]
    bad = [(A, B), (o.g1_neg(al), be), (o.g1_neg(C), de), (o.g1_neg(o.g1_mul(G1, 9)), ga)]
    mg, mb = o.multi_miller_loop(good), o.multi_miller_loop(bad)
    assert o.final_exponentiation(mg) == o.FP12_ONE and o.final_exponentiation(mb) != o.FP12_ONE
    vec["multi"] = [{"pairs": [{"g1": g1h(p), "g2": g2h(q)} for p, q in prs], "multi_miller": h(m),
                     "gt": h(o.final_exponentiation(m)), "is_one": one}
                    for prs, m, one in ((good, mg, True), (bad, mb, False))]
    with open(os.path.join(HERE, "pairing_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)
    print("wrote reference_kats.json, pairing_vectors.json")


if __name__ == "__main__":
    main()
