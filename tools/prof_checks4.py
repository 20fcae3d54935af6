#!/usr/bin/env python3
"""Profiling target for BASELINE config 3: ONE k_pairing<4> launch over 2^LOG2 Groth16-shaped 4-pair checks, plain and
with the three verifying-key G2 points prepared (after small warm-up launches of both).
Usage: python tools/prof_checks4.py [LOG2=16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nc, k, kf = 1 << log2, 4, 3
n = nc * k
eng = z.PairingEngine([0])
dev = torch.device("cuda", 0)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
st = s.cuda_stream
g1 = torch.empty((n, 12), dtype=torch.int64, device=dev)
g2 = torch.empty((n, 24), dtype=torch.int64, device=dev)
i1 = torch.empty(n, dtype=torch.uint8, device=dev)
i2 = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty((nc, 72), dtype=torch.int64, device=dev)
one = torch.empty(nc, dtype=torch.uint8, device=dev)
eng.gen_points_dev(11, 0, n, g1, i1, g2, i2, stream=st)
fixed = g2[:kf].contiguous()
tab = torch.empty((kf, eng.G2_PREPARED_U64), dtype=torch.int64, device=dev)
eng.g2_prepare_dev(fixed, kf, tab, stream=st)
var = g2.view(nc, k, 24)[:, 0, :].contiguous()
torch.cuda.synchronize()


def plain(m):
    eng.pairing_dev(z.MODE_PAIRING, out, g1=g1, g2=g2, n_checks=m, pairs_per_check=k, is_one=one, stream=st)


def prepared(m):
    eng.multi_pairing_prepared_dev(out, g1, var, m, k, tab, kf, is_one=one, stream=st)


warm = 256 if os.environ.get("ZKP_PROF_SMALL_WARMUP") else nc   # ncu captures skip the warm-up launches: keep them cheap there
plain(warm)
prepared(warm)
torch.cuda.synchronize()
for name, fn in (("plain", plain), ("prepared", prepared)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(nc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("checks4 %-8s n=2^%d checks  %.3f ms  %.3f M checks/s" % (name, log2, ms, nc / ms / 1e3))
eng.close()
