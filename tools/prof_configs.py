#!/usr/bin/env python3
"""Times the BASELINE.json configs that are not the bench line, on device-resident synthetic data:
  config 2: final exponentiation only on 2^20 Miller-loop outputs
  config 3: 2^18 Groth16-style product checks of 4 pairs (shared final exponentiation)
Usage: python tools/prof_configs.py [LOG2_CHECKS=18]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 18
nc, k = 1 << log2, 4
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
ml = torch.empty((n, 72), dtype=torch.int64, device=dev)
out = torch.empty((n, 72), dtype=torch.int64, device=dev)
one = torch.empty(nc, dtype=torch.uint8, device=dev)
eng.gen_points_dev(11, 0, n, g1, i1, g2, i2, stream=st)
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


eng.pairing_dev(z.MODE_MILLER, ml, g1=g1, g2=g2, stream=st)
ms = timed(lambda: eng.pairing_dev(z.MODE_FINAL_EXP, out, in_fp12=ml, stream=st))
print("config 2  final exponentiation only   n=2^%d   %.2f ms   %.3f M/s" % (log2 + 2, ms, n / ms / 1e3))
ms = timed(lambda: eng.pairing_dev(z.MODE_MILLER, ml, g1=g1, g2=g2, stream=st))
print("          Miller loop only            n=2^%d   %.2f ms   %.3f M/s" % (log2 + 2, ms, n / ms / 1e3))
ms = timed(lambda: eng.pairing_dev(z.MODE_PAIRING, out, g1=g1, g2=g2, n_checks=nc, pairs_per_check=k, is_one=one, stream=st))
print("config 3  4-pair product checks       n=2^%d checks   %.2f ms   %.3f M checks/s  (%.3f M pairs/s)" % (log2, ms, nc / ms / 1e3, n / ms / 1e3))
# config 3 with the three "verifying-key" G2 points of every check prepared once (G2Prepared tables)
kf = 3
fixed = g2[:kf].contiguous()
tab = torch.empty((kf, eng.G2_PREPARED_U64), dtype=torch.int64, device=dev)
eng.g2_prepare_dev(fixed, kf, tab, stream=st)
var = g2.view(nc, k, 24)[:, 0, :].contiguous()
ms = timed(lambda: eng.multi_pairing_prepared_dev(out, g1, var, nc, k, tab, kf, is_one=one, stream=st))
print("config 3  same, 3 of 4 G2 prepared     n=2^%d checks   %.2f ms   %.3f M checks/s" % (log2, ms, nc / ms / 1e3))
eng.close()
