"""ctypes binding of the C oracle (oracle/zkp_oracle.c -> oracle/libzkp_oracle.so).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.

All arrays are numpy uint64, canonical little-endian limbs, array-of-structs:
Fp (n,6)  Fp2 (n,12)  Fp6 (n,36)  Fp12 (n,72)  G1 (n,12)=x|y  G2 (n,24)=x.c0|x.c1|y.c0|y.c1,
infinity flags uint8 (n,).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzkp_oracle.so")

OPS = {
    "fp_add": 0, "fp_sub": 1, "fp_neg": 2, "fp_mul": 3, "fp_sqr": 4, "fp_inv": 5, "fp_sqrt": 6,
    "fp2_add": 16, "fp2_sub": 17, "fp2_neg": 18, "fp2_mul": 19, "fp2_sqr": 20, "fp2_inv": 21,
    "fp2_mul_nr": 22, "fp2_conj": 23,
    "fp6_add": 32, "fp6_sub": 33, "fp6_neg": 34, "fp6_mul": 35, "fp6_sqr": 36, "fp6_inv": 37,
    "fp6_mul_nr": 38, "fp6_frob": 39, "fp6_mul_by_1": 40, "fp6_mul_by_01": 41,
    "fp12_add": 48, "fp12_sub": 49, "fp12_neg": 50, "fp12_mul": 51, "fp12_sqr": 52, "fp12_inv": 53,
    "fp12_conj": 54, "fp12_frob": 55, "fp12_mul_by_014": 56, "fp12_cyc_sqr": 57, "fp12_cyc_exp": 58,
}
_WIDTH = {"fp_": 6, "fp2_": 12, "fp6_": 36, "fp12_": 72}
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "zkp_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libzkp_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def ncores():
    return len(os.sched_getaffinity(0))


def tower_op(name, a, b=None, want_ok=False):
    w = next(v for k, v in _WIDTH.items() if name.startswith(k))
    a = _u64(a).reshape(-1, w)
    n = a.shape[0]
    b = None if b is None else _u64(b).reshape(n, -1)
    out = np.zeros_like(a)
    ok = np.ones(n, dtype=np.uint8)
    rc = lib().zo_tower_op(OPS[name], _p(a), _p(b), _p(out), _p(ok), ctypes.c_size_t(n))
    assert rc == 0, rc
    return (out, ok) if want_ok else out


def _pairs(fn, g1, g1inf, g2, g2inf, n, k, out, extra, nthreads):
    g1, g2 = _u64(g1).reshape(-1, 12), _u64(g2).reshape(-1, 24)
    assert g1.shape[0] == g2.shape[0] == n * k
    g1inf, g2inf = _u8(g1inf), _u8(g2inf)
    args = [_p(g1), _p(g1inf), _p(g2), _p(g2inf), ctypes.c_size_t(n)]
    if k is not None and fn.__name__.startswith("zo_multi"):
        args.append(ctypes.c_int(k))
    args.append(_p(out))
    args.extend(extra)
    args.append(ctypes.c_int(nthreads or ncores()))
    rc = fn(*args)
    if rc != 0:
        raise ValueError("oracle: non-canonical input (rc=%d)" % rc)
    return out


def miller_loop_batch(g1, g1inf, g2, g2inf, nthreads=None):
    n = _u64(g1).reshape(-1, 12).shape[0]
    return _pairs(lib().zo_miller_loop_batch, g1, g1inf, g2, g2inf, n, 1, np.zeros((n, 72), np.uint64), [], nthreads)


def pairing_batch(g1, g1inf, g2, g2inf, nthreads=None):
    n = _u64(g1).reshape(-1, 12).shape[0]
    return _pairs(lib().zo_pairing_batch, g1, g1inf, g2, g2inf, n, 1, np.zeros((n, 72), np.uint64), [], nthreads)


def multi_miller_batch(g1, g1inf, g2, g2inf, k, nthreads=None):
    n = _u64(g1).reshape(-1, 12).shape[0] // k
    return _pairs(lib().zo_multi_miller_batch, g1, g1inf, g2, g2inf, n, k, np.zeros((n, 72), np.uint64), [], nthreads)


def multi_pairing_batch(g1, g1inf, g2, g2inf, k, nthreads=None):
    n = _u64(g1).reshape(-1, 12).shape[0] // k
    is_one = np.zeros(n, np.uint8)
    out = _pairs(lib().zo_multi_pairing_batch, g1, g1inf, g2, g2inf, n, k, np.zeros((n, 72), np.uint64), [_p(is_one)], nthreads)
    return out, is_one


def final_exp_batch(f, nthreads=None):
    f = _u64(f).reshape(-1, 72)
    out = np.zeros_like(f)
    rc = lib().zo_final_exp_batch(_p(f), ctypes.c_size_t(f.shape[0]), _p(out), ctypes.c_int(nthreads or ncores()))
    if rc != 0:
        raise ValueError("oracle: non-canonical input")
    return out


def miller_product(g1, g1inf, g2, g2inf):
    g1, g2 = _u64(g1).reshape(-1, 12), _u64(g2).reshape(-1, 24)
    ml, gt = np.zeros(72, np.uint64), np.zeros(72, np.uint64)
    rc = lib().zo_miller_product(_p(g1), _p(_u8(g1inf)), _p(g2), _p(_u8(g2inf)), ctypes.c_size_t(g1.shape[0]), _p(ml), _p(gt))
    if rc != 0:
        raise ValueError("oracle: non-canonical input")
    return ml, gt


def _mul(fn, w, base, binf, scalars, nthreads):
    k = _u64(scalars).reshape(-1, 4)
    n = k.shape[0]
    base = None if base is None else _u64(base).reshape(n, w)
    out, oinf = np.zeros((n, w), np.uint64), np.zeros(n, np.uint8)
    rc = fn(_p(base), _p(_u8(binf)), _p(k), ctypes.c_size_t(n), _p(out), _p(oinf), ctypes.c_int(nthreads or ncores()))
    if rc != 0:
        raise ValueError("oracle: non-canonical input")
    return out, oinf


def g1_mul_batch(scalars, base=None, binf=None, nthreads=None):
    """[k_i]P_i (P_i = generator when base is None); scalars (n,4) little-endian u64."""
    return _mul(lib().zo_g1_mul_batch, 12, base, binf, scalars, nthreads)


def g2_mul_batch(scalars, base=None, binf=None, nthreads=None):
    return _mul(lib().zo_g2_mul_batch, 24, base, binf, scalars, nthreads)


def group_op(group, op, a, ainf=0, b=None, binf=0):
    """op: 'double' | 'add' | 'on_curve' | 'torsion_free'."""
    code = {"double": 0, "add": 1, "on_curve": 2, "torsion_free": 3}[op]
    w = 12 if group == "g1" else 24
    a = _u64(a).reshape(w)
    b = None if b is None else _u64(b).reshape(w)
    out, flag = np.zeros(w, np.uint64), ctypes.c_uint8(0)
    fn = lib().zo_g1_op if group == "g1" else lib().zo_g2_op
    rc = fn(code, _p(a), ctypes.c_uint8(ainf), _p(b), ctypes.c_uint8(binf), _p(out), ctypes.byref(flag))
    assert rc == 0
    return (out, flag.value) if code < 2 else bool(flag.value)


def constants():
    r2 = np.zeros(6, np.uint64)
    c = [np.zeros(12, np.uint64) for _ in range(3)]
    lib().zo_constants(_p(r2), _p(c[0]), _p(c[1]), _p(c[2]))
    return r2, c[0], c[1], c[2]
