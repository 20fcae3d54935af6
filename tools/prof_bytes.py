#!/usr/bin/env python3
"""HBM roofline of the byte (de)serialisation kernel k_fp_bytes on device-resident data:
python tools/prof_bytes.py [LOG2]   (96 algorithmic bytes per element: 48 read + 48 written)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log2
eng = z.PairingEngine([0])
dev = torch.device("cuda", 0)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
src = torch.randint(0, 256, (n, 48), dtype=torch.uint8, device=dev)
dst = torch.empty((n, 6), dtype=torch.int64, device=dev)
ok = torch.empty(n, dtype=torch.uint8, device=dev)
peak = 6650.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for direction, name in ((0, "from_bytes"), (1, "to_bytes")):
    a, b = (src, dst) if direction == 0 else (dst, src)
    for _ in range(3):
        eng.fp_bytes_dev(direction, a, b, n, ok=ok if direction == 0 else None, stream=s.cuda_stream)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.fp_bytes_dev(direction, a, b, n, ok=ok if direction == 0 else None, stream=s.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    gbs = n * 96 / (best * 1e-3) / 1e9
    print("k_fp_bytes %-10s n=2^%d  %.3f ms  %.0f GB/s  = %.1f%% of the measured copy peak %.0f GB/s  (%.1f G elements/s)"
          % (name, log2, best, gbs, 100 * gbs / peak, peak, n / best / 1e6))
eng.close()
