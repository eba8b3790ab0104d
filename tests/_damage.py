"""Seeded damage generators shared by the host and GPU parity tests."""
import numpy as np


def header_offsets(data: bytes):
    """Byte offsets that belong to marker segments (not to entropy-coded data)."""
    hdr, i = [], 2
    while i + 4 <= len(data):
        assert data[i] == 0xFF
        m = data[i + 1]
        if m == 0xD9:
            break
        seglen = (data[i + 2] << 8) | data[i + 3]
        hdr += list(range(i, i + 2 + seglen))
        i += 2 + seglen
        if m == 0xDA:  # skip the entropy-coded data up to the next real marker
            while i + 1 < len(data) and not (data[i] == 0xFF and data[i + 1] != 0 and not 0xD0 <= data[i + 1] <= 0xD7):
                i += 1
    return hdr


def header_damage(data: bytes, rng: np.random.Generator, count: int):
    """`count` copies of data with 1-3 random bytes of its marker segments overwritten."""
    hdr = header_offsets(data)
    out = []
    for _ in range(count):
        d = bytearray(data)
        for k in rng.choice(hdr, size=int(rng.integers(1, 4)), replace=False):
            d[int(k)] = int(rng.integers(0, 256))
        out.append(bytes(d))
    return out
