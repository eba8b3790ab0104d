"""Lane-by-lane Python model of k0_unstuff (zpix_b200/csrc/zpx_k0.cu): the kernel's index logic (not its speed) against a
plain FF 00 -> FF replacement on random pieces.  `python tools/k0_model.py N` runs N seeds; tests/test_host.py runs 300."""
import random, sys
M32 = 0xffffffff
def popc(x): return bin(x & M32).count("1")

def run_piece(blob, ublob, src, dst, ln, flags):
    a0 = src & ~15
    m = dst & 15
    gptr = dst - m
    tot = m
    head_shared = m != 0
    s_in = bytearray(16 + 512)
    out = bytearray(b"\xAA" * 544)
    carry31 = 0
    rel0 = -(src - a0)
    gin = a0
    def word(buf, off): return int.from_bytes(buf[off:off + 4], "little")
    while rel0 < max(ln, 1):
        vw31 = 0
        for lane in range(32):
            v = bytes(16)
            if rel0 + 16 * lane < ln:
                v = bytes(blob[gin + 16 * lane: gin + 16 * lane + 16]); assert len(v) == 16, "read past blob"
            s_in[16 + 16 * lane: 32 + 16 * lane] = v
            if lane == 31: vw31 = int.from_bytes(v[12:16], "little")
        s_in[12:16] = carry31.to_bytes(4, "little")
        carry31 = vw31
        full = rel0 >= 1 and rel0 + 512 <= ln
        for k in range(4):
            rs = rel0 + 128 * k
            xs, drops = [], []
            for lane in range(32):
                x = word(s_in, 16 + 4 * (32 * k + lane)); pw = word(s_in, 16 + 4 * (32 * k + lane) - 4)
                f = ((x << 8) | (pw >> 24)) & M32
                t = (x | (~f & M32)) & M32
                drop = ~((((t & 0x7f7f7f7f) + 0x7f7f7f7f) | t) | 0x7f7f7f7f) & M32
                xs.append(x); drops.append(drop)
            if full or (rs >= 1 and rs + 128 <= ln):
                if not any(drops):
                    for lane in range(32):
                        o = 4 * lane + tot
                        out[o:o + 4] = xs[lane].to_bytes(4, "little")
                    tot += 128
                else:
                    ng = [popc(d) for d in drops]
                    b1 = sum((1 << l) for l in range(32) if ng[l] >= 1); b2 = sum((1 << l) for l in range(32) if ng[l] >= 2)
                    assert all(n <= 2 for n in ng)
                    for lane in range(32):
                        lt = (1 << lane) - 1
                        o = 4 * lane + tot - (popc(b1 & lt) + popc(b2 & lt))
                        for i in range(4):
                            if not (drops[lane] & (0x80 << (8 * i))):
                                out[o] = (xs[lane] >> (8 * i)) & 255; o += 1
                    tot += 128 - (popc(b1) + popc(b2))
            elif rs < ln:
                gone = []
                for lane in range(32):
                    rel = rs + 4 * lane; g = drops[lane]
                    for i in range(4):
                        if rel + i == 0: g &= ~(0x80 << (8 * i))
                        if rel + i < 0 or rel + i >= ln: g |= 0x80 << (8 * i)
                    gone.append(g & M32)
                ng = [popc(g) for g in gone]
                bs = [sum((1 << l) for l in range(32) if ng[l] >= j) for j in (1, 2, 3, 4)]
                for lane in range(32):
                    lt = (1 << lane) - 1
                    o = 4 * lane + tot - sum(popc(b & lt) for b in bs)
                    for i in range(4):
                        if not (gone[lane] & (0x80 << (8 * i))):
                            out[o] = (xs[lane] >> (8 * i)) & 255; o += 1
                tot += 128 - sum(popc(b) for b in bs)
        if rel0 + 512 >= ln and (flags & 1):
            padn = (16 - (tot & 15)) & 15
            for lane in range(32):
                if lane < padn: out[tot + lane] = 0
            tot += padn
        assert tot <= 544
        nvec, rest = tot >> 4, tot & 15
        for lane in range(32):
            if lane < nvec and not (head_shared and lane == 0):
                ublob[gptr + 16 * lane: gptr + 16 * lane + 16] = out[16 * lane: 16 * lane + 16]
        if nvec > 32: ublob[gptr + 512: gptr + 528] = out[512:528]
        if head_shared and nvec > 0:
            for lane in range(32):
                if m <= lane < 16: ublob[gptr + lane] = out[lane]
            head_shared = False
        keep = [out[16 * nvec + lane] if lane < rest else 0 for lane in range(32)]
        for lane in range(32):
            if lane < rest and nvec > 0: out[lane] = keep[lane]
        gptr += 16 * nvec
        tot = rest
        rel0 += 512; gin += 512
    for lane in range(32):
        if lane < tot and not (head_shared and lane < m): ublob[gptr + lane] = out[lane]

def make_interval(rng, n):
    """random entropy-coded bytes with FF 00 pairs (no other FF xx), about n bytes"""
    b = bytearray()
    while len(b) < n:
        r = rng.random()
        if r < rng.choice([0.003, 0.02, 0.3]): b += b"\xff\x00"
        elif r < 0.4: b.append(0)
        else: b.append(rng.randrange(0, 255))
    return bytes(b)

def test(seed, big=False):
    rng = random.Random(seed)
    pre = rng.randrange(0, 40)
    sizes = [1, 2, 5, 15, 16, 17, 100, 127, 128, 129, 500, 511, 512, 513, 640, 1024, 3000] if not big else [6400, 16383, 16385, 40000]
    ivs = [make_interval(rng, rng.choice(sizes)) for _ in range(rng.randrange(1, 5))]
    blob = bytearray(rng.randbytes(pre))
    pieces, dstpos, exp = [], 0, bytearray()
    seg = rng.choice([64, 200, 512, 1000, 4000]) if not big else 16384  # (ZPX_SEG_BYTES)
    for iv in ivs:
        src0 = len(blob); blob += iv; blob += rng.randbytes(rng.randrange(0, 7))
        ustart = dstpos
        # cut into pieces of about seg raw bytes that never split an FF 00 pair
        cuts = [0]
        while len(iv) - cuts[-1] > seg:
            c = cuts[-1] + seg
            if iv[c - 1] == 0xff: c -= 1
            if c <= cuts[-1]: c = cuts[-1] + seg + 1
            cuts.append(c)
        cuts.append(len(iv))
        un = iv.replace(b"\xff\x00", b"\xff")
        for q in range(len(cuts) - 1):
            a, b_ = cuts[q], cuts[q + 1]
            uoff = len(iv[:a].replace(b"\xff\x00", b"\xff"))
            pieces.append((src0 + a, ustart + uoff, b_ - a, 1 if q == len(cuts) - 2 else 0))
        padded = (len(un) + 15) & ~15
        exp += un + bytes(padded - len(un))
        dstpos += padded
    blob += bytes(600)  # the kernel may read up to the 16-byte boundary past a piece; (device blobs are padded)
    ublob = bytearray(b"\x55" * (dstpos + 16))
    order = list(range(len(pieces))); rng.shuffle(order)
    for i in order: run_piece(blob, ublob, *pieces[i])
    assert ublob[:dstpos] == exp, f"seed {seed}: mismatch"
    assert ublob[dstpos:] == b"\x55" * 16, f"seed {seed}: wrote past the end"

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for s in range(n): test(s)
    for s in range(max(1, n // 50)): test(s, big=True)
    print("ok", n)
