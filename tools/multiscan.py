"""Re-code a baseline interleaved JPEG as a multi-scan SEQUENTIAL one (test data generator).

Pillow / libjpeg-turbo only write single-scan baseline files; the reference (like Go's image/jpeg) also
decodes sequential frames whose components come in separate scans (`processSos` with fewer scan components
than frame components, non-interleaved block order, src/jpeg/decoder.zig:1294-1336).  This tool takes the
quantised coefficients of a single-scan file (read with the oracle's tap) and writes them again with the
file's own Huffman tables (which must be complete: encode the source with optimize=False, Annex K tables)
in a scan script of your choice, optionally with restart intervals.

    scans = [[0], [1], [2]]            # one scan per component
    scans = [[0], [1, 2]]              # luma alone, chroma interleaved
    data = recode(src_bytes, scans, restart_interval=7)
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
          62, 63]


def segments(data: bytes):
    """[(marker, payload offset, payload length)] up to and including the first SOS header."""
    out, i = [], 2
    while i + 4 <= len(data):
        assert data[i] == 0xFF, hex(data[i])
        m = data[i + 1]
        n = (data[i + 2] << 8) | data[i + 3]
        out.append((m, i + 4, n - 2))
        i += 2 + n
        if m == 0xDA:
            break
    return out


def huff_codes(counts, vals):
    """symbol -> (code, length) of a canonical table"""
    codes, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            codes[vals[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return codes


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value: int, length: int):
        self.acc = (self.acc << length) | (value & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1 if self.n else 0

    def flush(self):  # pad with ones to a byte boundary
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def magnitude(v: int):
    if v == 0:
        return 0, 0
    size = int(abs(v)).bit_length()
    return size, (v if v > 0 else v + (1 << size) - 1)


def recode(src: bytes, scans, restart_interval: int = 0) -> bytes:
    from oracle import oracle as O

    img, recs = O.decode(src, tap=True)
    segs = segments(src)
    sof = next(s for s in segs if s[0] in (0xC0, 0xC1))
    p = sof[1]
    height, width, nc = (src[p + 1] << 8) | src[p + 2], (src[p + 3] << 8) | src[p + 4], src[p + 5]
    comp = []
    for c in range(nc):
        cid, hv, tq = src[p + 6 + 3 * c: p + 9 + 3 * c]
        comp.append(dict(id=cid, h=hv >> 4, v=hv & 15))
    hmax, vmax = comp[0]["h"], comp[0]["v"]
    mxx, myy = (width + 8 * hmax - 1) // (8 * hmax), (height + 8 * vmax - 1) // (8 * vmax)
    # tables
    dc_tab, ac_tab = {}, {}
    for m, off, n in segs:
        if m != 0xC4:
            continue
        j = off
        while j < off + n:
            tc, th = src[j] >> 4, src[j] & 15
            counts = list(src[j + 1: j + 17])
            vals = list(src[j + 17: j + 17 + sum(counts)])
            (ac_tab if tc else dc_tab)[th] = huff_codes(counts, vals)
            j += 17 + sum(counts)
    sos = segs[-1]
    q = sos[1]
    sel = {}
    for k in range(src[q]):
        sel[src[q + 1 + 2 * k]] = (src[q + 2 + 2 * k] >> 4, src[q + 2 + 2 * k] & 15)
    # coefficient blocks by (component, bx, by)
    blocks = {(int(r[0]), int(r[1]), int(r[2])): r[3:] for r in recs}

    out = bytearray(src[: sos[1] - 4])  # everything before the SOS marker
    # drop an existing DRI, add ours
    if restart_interval:
        out += bytes([0xFF, 0xDD, 0, 4, restart_interval >> 8, restart_interval & 255])

    def encode_block(bw: BitWriter, c: int, coef, pred):
        td, ta = sel[comp[c]["id"]]
        dct, act = dc_tab[td], ac_tab[ta]
        diff = int(coef[0]) - pred[c]
        pred[c] = int(coef[0])
        size, bits = magnitude(diff)
        bw.put(*dct[size])
        if size:
            bw.put(bits, size)
        run = 0
        for z in range(1, 64):
            v = int(coef[ZIGZAG[z]])
            if v == 0:
                run += 1
                continue
            while run > 15:
                bw.put(*act[0xF0])
                run -= 16
            size, bits = magnitude(v)
            bw.put(*act[(run << 4) | size])
            bw.put(bits, size)
            run = 0
        if run:
            bw.put(*act[0x00])

    for scan in scans:
        out += bytes([0xFF, 0xDA, 0, 6 + 2 * len(scan), len(scan)])
        for c in scan:
            td, ta = sel[comp[c]["id"]]
            out += bytes([comp[c]["id"], td << 4 | ta])
        out += bytes([0, 63, 0])
        # the units ("MCUs") of this scan, in the order processSos visits them
        units = []
        if len(scan) == 1:
            c = scan[0]
            h, v = comp[c]["h"], comp[c]["v"]
            cw = (width * h + 8 * hmax - 1) // (8 * hmax)   # blocks that intersect the image
            ch = (height * v + 8 * vmax - 1) // (8 * vmax)
            # decoder.zig:1331-1336: row-major over the padded grid, only blocks with 8*bx < width and 8*by < height
            # (the reference tests against the IMAGE size)
            for by in range(myy * v):
                for bx in range(mxx * h):
                    if bx * 8 < width and by * 8 < height:
                        units.append([(c, bx, by)])
            del cw, ch
        else:
            for my in range(myy):
                for mx in range(mxx):
                    u = []
                    for c in scan:
                        h, v = comp[c]["h"], comp[c]["v"]
                        for j in range(h * v):
                            u.append((c, h * mx + j % h, v * my + j // h))
                    units.append(u)
        bw = BitWriter()
        pred = [0] * nc
        rst = 0
        # restart intervals count MCU ITERATIONS of the scan; for a one-component scan of a sub-sampled frame the
        # reference iterates mxx*myy times with h*v blocks each (some skipped): group accordingly
        if len(scan) == 1:
            c = scan[0]
            h, v = comp[c]["h"], comp[c]["v"]
            per_iter, it = h * v, []
            # rebuild iteration structure: linear block counter over the padded grid
            allb = [(bx, by) for by in range(myy * v) for bx in range(mxx * h)]
            iters = [allb[k: k + per_iter] for k in range(0, len(allb), per_iter)]
            units = [[(c, bx, by) for (bx, by) in itr if bx * 8 < width and by * 8 < height] for itr in iters]
        for k, u in enumerate(units):
            for (c, bx, by) in u:
                encode_block(bw, c, blocks.get((c, bx, by), np.zeros(64, np.int32)), pred)
            if restart_interval and (k + 1) % restart_interval == 0 and k + 1 < len(units):
                bw.flush()
                bw.out += bytes([0xFF, 0xD0 + rst])
                rst = (rst + 1) & 7
                pred = [0] * nc
        bw.flush()
        out += bw.out
    out += b"\xff\xd9"
    return bytes(out)


if __name__ == "__main__":
    from oracle import oracle as O
    from tools import synth_jpeg as S

    base = S.encode(60000, 97, 75, subsampling="4:2:0")
    ref = O.decode(base).rgbaPixels()
    for scans, ri in [([[0], [1], [2]], 0), ([[0], [1, 2]], 0), ([[2], [0], [1]], 3), ([[0, 1, 2]], 0), ([[0, 1], [2]], 5)]:
        d = recode(base, scans, ri)
        got = O.decode(d).rgbaPixels()
        print(scans, ri, len(d), "same pixels as the single-scan file:", bool(np.array_equal(got, ref)))
