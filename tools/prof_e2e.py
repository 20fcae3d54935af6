#!/usr/bin/env python3
"""End-to-end timing of zkp_pairing_batch from PINNED host buffers (the bench's `e2e` leg alone):
python tools/prof_e2e.py [LOG2] [STEPS]   (env ZKP_TAPER=0/1 selects the chunk schedule, ZKPAIR_LIB the build)"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zkvm_pairings_b200 as z

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1 << log2
eng = z.PairingEngine([0])
dev = torch.device("cuda", 0)
g1 = torch.empty((n, 12), dtype=torch.int64, device=dev)
g2 = torch.empty((n, 24), dtype=torch.int64, device=dev)
i1 = torch.empty(n, dtype=torch.uint8, device=dev)
i2 = torch.empty(n, dtype=torch.uint8, device=dev)
eng.gen_points_dev(7, 0, n, g1, i1, g2, i2)
torch.cuda.synchronize()
h_g1 = torch.empty((n, 12), dtype=torch.int64).pin_memory()
h_g2 = torch.empty((n, 24), dtype=torch.int64).pin_memory()
h_out = torch.empty((n, 72), dtype=torch.int64).pin_memory()
h_g1.copy_(g1)
h_g2.copy_(g2)
lib = z._lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())


def step():
    rc = lib.zkp_pairing_batch(eng._ctx, p(h_g1), None, p(h_g2), None, n, p(h_out))
    assert rc == 0, rc


step()
step()
t0 = time.perf_counter()
for _ in range(steps):
    step()
dt = (time.perf_counter() - t0) / steps
chk = int(h_out.view(-1)[:: 4099].sum().item()) & 0xFFFFFFFF
print("e2e n=2^%d taper=%s  %.3f ms  %.0f /s  checksum %08x" % (log2, os.environ.get("ZKP_TAPER", "default"), dt * 1e3, n / dt, chk))
eng.close()
