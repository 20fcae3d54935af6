#!/usr/bin/env python3
"""Experiment: Miller-loop kernel and final-exponentiation kernel CONCURRENTLY on two streams (different
data), against running them back to back.  ZKP_SMEM_MILLER / ZKP_SMEM_FE (bytes of dynamic shared memory
per block) steer the block mix per SM.  Usage: python tools/prof_overlap.py [LOG2=18]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << log2
eng = z.PairingEngine([0])
dev = torch.device("cuda", 0)
sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
g1 = torch.empty((n, 12), dtype=torch.int64, device=dev)
g2 = torch.empty((n, 24), dtype=torch.int64, device=dev)
i1 = torch.empty(n, dtype=torch.uint8, device=dev)
i2 = torch.empty(n, dtype=torch.uint8, device=dev)
ml = torch.empty((n, 72), dtype=torch.int64, device=dev)
ml2 = torch.empty((n, 72), dtype=torch.int64, device=dev)
out = torch.empty((n, 72), dtype=torch.int64, device=dev)
eng.gen_points_dev(7, 0, n, g1, i1, g2, i2, stream=sa.cuda_stream)
eng.pairing_dev(1, ml, g1=g1, g2=g2, stream=sa.cuda_stream)
eng.pairing_dev(2, out, in_fp12=ml, stream=sa.cuda_stream)
torch.cuda.synchronize()


def run(miller, fe):
    ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ea0.record(sa); eb0.record(sb)
    if miller:
        eng.pairing_dev(1, ml2, g1=g1, g2=g2, stream=sa.cuda_stream)
    if fe:
        eng.pairing_dev(2, out, in_fp12=ml, stream=sb.cuda_stream)
    ea1.record(sa); eb1.record(sb)
    torch.cuda.synchronize()
    return ea0.elapsed_time(ea1), eb0.elapsed_time(eb1), max(ea0.elapsed_time(ea1), ea0.elapsed_time(eb1))


for _ in range(2):
    a = run(True, False)[0]
    b = run(False, True)[1]
    ta, tb, both = run(True, True)
    print("n=2^%d  Miller alone %.1f ms, final exp alone %.1f ms, sum %.1f ms | concurrent: Miller %.1f, final exp %.1f, wall %.1f ms  (%.2fx)"
          % (log2, a, b, a + b, ta, tb, both, (a + b) / both))
eng.close()
