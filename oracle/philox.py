"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) in numpy -- TEST INFRASTRUCTURE: the checker for the device RNG (integer, bit-exact).
Known-answer vectors are the Random123 kat_vectors entries for philox4x32 10 rounds."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(ctr: np.ndarray, key0: int, key1: int) -> np.ndarray:
    """ctr: uint32 [n,4] -> uint32 [n,4]."""
    c = np.asarray(ctr, dtype=np.uint32).reshape(-1, 4).copy()
    k0, k1 = np.uint32(key0), np.uint32(key1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[:, 0].astype(np.uint64)
            p1 = M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c


def u01(x: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in (0,1), same map as csrc/vbn_device.cuh:u01."""
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-24)
