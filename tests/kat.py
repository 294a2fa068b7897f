"""Known-answer vectors for Philox4x32-10 (Random123 v1.14, examples/kat_vectors, philox4x32 10)."""
import numpy as np

PHILOX4X32_10_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def kat_inputs():
    return np.array([list(c) + list(k) for c, k, _ in PHILOX4X32_10_KAT], dtype=np.uint32)


def kat_outputs():
    return np.array([list(o) for _, _, o in PHILOX4X32_10_KAT], dtype=np.uint32)
