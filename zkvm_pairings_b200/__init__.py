"""zkvm_pairings_b200 -- B200-native batched BLS12-381 pairing engine.

Drop-in for the pairing hot path of 0xWOLAND/zkvm-pairings: hand-written sm_100a CUDA behind the
C ABI of include/zkpair.h (libzkpair.so), driven from Python through ctypes.  No CPU fallback.
"""
from .engine import (MODE_FINAL_EXP, MODE_MILLER, MODE_MILLER_FOR_FINAL_EXP, MODE_PAIRING, TOWER_OPS, NonCanonicalError, PairingEngine,
                     ZkpError, device_count, op_widths)

__all__ = ["PairingEngine", "ZkpError", "NonCanonicalError", "TOWER_OPS", "MODE_MILLER", "MODE_FINAL_EXP",
           "MODE_PAIRING", "MODE_MILLER_FOR_FINAL_EXP", "device_count", "op_widths"]
