"""Bug hunt, not a test: many seeds of header / entropy damage over the reference's fixtures, GPU result against the
oracle under the same rules as tests/test_gpu_parity.py::_assert_same.  Prints every disagreement and saves the file.

  python tools/fuzz_hunt.py [--seeds 20] [--out gpurun_out/fuzz]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _damage import header_damage  # noqa: E402
from oracle import oracle as O  # noqa: E402
from zpix_b200 import jpeg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=20)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--out", default="gpurun_out/fuzz")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--native", type=int, default=0, help="1: also compare the native variant (planes / interleave) byte for byte; 2: the same with ZPX_OPT_FORCE_GENERIC")
ap.add_argument("--structural", type=int, default=0, help="N more files per base with marker-level damage")
ap.add_argument("--synth", type=int, default=0, help="1: synthetic base files (other shapes, DRI, CMYK, YCbCrK) instead of the fixtures")
ap.add_argument("--prog-only", type=int, default=0, help="1: progressive base files only")
ap.add_argument("--prog-mode", type=int, default=0, help="ZPX_OPT_PROGRESSIVE_MODE: 0 lane per scan where the script allows, 1 warp per scan")
a = ap.parse_args()
os.makedirs(a.out, exist_ok=True)
FX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")


def entropy_damage(data, rng, n_trunc, n_flip):
    first_sos = data.index(b"\xff\xda")
    out = [data[: int(k)] for k in rng.integers(first_sos + 14, len(data) - 2, n_trunc)]
    for k in rng.integers(first_sos + 14, len(data) - 2, n_flip):
        d = bytearray(data)
        d[int(k)] = int(rng.integers(0, 256))
        out.append(bytes(d))
    return out


def bases():
    if not a.synth:
        return [open(os.path.join(FX, n), "rb").read() for n in sorted(os.listdir(FX)) if n != "iceberg.jpg"]
    from tools import synth_jpeg as S
    if a.synth == 2:
        # larger files: segments of several hundred sub-sequences (many warps per segment in the self-synchronising
        # decoder, its boundary kernel and later sweeps), long restart intervals
        kw2 = [dict(subsampling="4:2:0"), dict(subsampling="4:4:4"), dict(mode="L"), dict(mode="CMYK"),
               dict(subsampling="4:2:0", restart_rows=8), dict(subsampling="4:2:2", restart_rows=3),
               dict(subsampling="4:2:0", quality=97), dict(subsampling="4:2:0", progressive=True)]
        out = [S.encode(71000 + i, 640, 480, **k) for i, k in enumerate(kw2)]
        # gray rows cut into several fused-kernel tiles, progressive gray / 4:2:2, multi-scan frames of several tiles per row
        from tools.multiscan import recode
        out += [S.encode(71100, 1100, 40, mode="L"), S.encode(71101, 1100, 40, mode="L", progressive=True),
                S.encode(71102, 1363, 48, subsampling="4:2:2", progressive=True),
                recode(S.encode(71104, 1363, 48, subsampling="4:2:2"), [[0], [1], [2]], 0),
                recode(S.encode(71106, 1920, 32, subsampling="4:2:0"), [[1], [0, 2]], 0)]
        return out
    kw = [dict(subsampling="4:2:0"), dict(subsampling="4:2:0", restart_rows=1), dict(subsampling="4:2:0", restart_blocks=3),
          dict(subsampling="4:2:2"), dict(subsampling="4:2:2", restart_blocks=5), dict(subsampling="4:4:4"),
          dict(subsampling="4:4:4", restart_rows=1), dict(mode="L"), dict(mode="L", restart_blocks=4), dict(mode="CMYK"),
          dict(mode="CMYK", ycck=True), dict(mode="CMYK", restart_rows=1), dict(subsampling="4:2:0", progressive=True),
          dict(subsampling="4:2:0", progressive=True, restart_rows=1), dict(subsampling="4:4:4", progressive=True),
          dict(mode="L", progressive=True), dict(mode="CMYK", progressive=True), dict(subsampling="4:2:0", quality=30),
          dict(subsampling="4:2:0", quality=98), dict(subsampling="4:2:2", progressive=True, restart_blocks=7)]
    sizes = [(211, 157), (97, 64), (320, 96), (33, 250)]
    out = [S.encode(70000 + i, *sizes[i % 4], **k) for i, k in enumerate(kw)]
    # sequential frames re-coded into several scans (tools/multiscan.py)
    from tools.multiscan import recode
    b420, b444 = S.encode(70100, 97, 75, subsampling="4:2:0"), S.encode(70102, 64, 48, subsampling="4:4:4")
    bcmyk = S.encode(70104, 83, 41, mode="CMYK")
    out += [recode(b420, [[0], [1], [2]], 0), recode(b420, [[0], [1, 2]], 4), recode(b444, [[2], [0], [1]], 3),
            recode(b420, [[0, 1], [2]], 5), recode(bcmyk, [[0], [1], [2], [3]], 0), recode(bcmyk, [[0, 1], [2, 3]], 2)]
    return out


BASES = [b for b in bases() if not a.prog_only or b"\xff\xc2" in b[:4096]]
def structural_damage(data, rng, count):
    """Marker-level damage: stray / missing / renumbered RSTn, inserted and deleted bytes, a scan played twice,
    a segment moved, a changed restart interval."""
    first_sos = data.index(b"\xff\xda")
    out = []
    for _ in range(count):
        d = bytearray(data)
        op = int(rng.integers(0, 7))
        k = int(rng.integers(first_sos + 14, max(first_sos + 15, len(d) - 2)))
        if op == 0:      # stray restart marker
            d[k:k] = bytes([0xFF, 0xD0 + int(rng.integers(0, 8))])
        elif op == 1:    # delete a byte
            del d[k]
        elif op == 2:    # insert a byte
            d.insert(k, int(rng.integers(0, 256)))
        elif op == 3:    # renumber / remove an existing restart marker
            rst = [i for i in range(first_sos, len(d) - 1) if d[i] == 0xFF and 0xD0 <= d[i + 1] <= 0xD7]
            if rst:
                i = rst[int(rng.integers(0, len(rst)))]
                if rng.integers(0, 2):
                    d[i + 1] = 0xD0 + int(rng.integers(0, 8))
                else:
                    del d[i:i + 2]
        elif op == 4:    # the tail of the file (from a random scan on) played twice
            sos = [i for i in range(len(d) - 1) if d[i] == 0xFF and d[i + 1] == 0xDA]
            i = sos[int(rng.integers(0, len(sos)))]
            d = d[:-2] + d[i:]
        elif op == 5:    # junk before a marker
            mk = [i for i in range(2, len(d) - 1) if d[i] == 0xFF and d[i + 1] in (0xC4, 0xDA, 0xDB, 0xD9)]
            i = mk[int(rng.integers(0, len(mk)))]
            d[i:i] = bytes(rng.integers(0, 255, int(rng.integers(1, 4))).astype("uint8"))
        else:            # a DRI segment before the first scan
            d[first_sos:first_sos] = bytes([0xFF, 0xDD, 0, 4, 0, int(rng.integers(0, 9))])
        out.append(bytes(d))
    return out


ctx = jpeg.Context([0])
ctx.set_option(1, a.mode)
ctx.set_option(10, a.prog_mode)
if a.native == 2:
    ctx.set_option(2, 1)  # 2: everything on the unfused kernels (1: the fused kernel's images get their planes on demand)
bad = total = 0
for seed in range(a.first, a.first + a.seeds):
    rng = np.random.default_rng(900000 + seed)
    datas = []
    for base in BASES:
        datas += header_damage(base, rng, 20) + entropy_damage(base, rng, 4, 16)
        if a.structural:
            datas += structural_damage(base, rng, a.structural)
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        b.decode()
        outs, st = b.fetch_rgba()
        nats = b.fetch_native()[0] if a.native else None
    for i, d in enumerate(datas):
        total += 1
        try:
            img = O.decode(d)
            want, err = img.rgbaPixels(), "ok"
        except O.OracleError as e:
            want, err = None, e.name
        ovf = O.last_coef_overflow()
        got = jpeg.lib.zpx_error_name(st[i]).decode() if st[i] else "ok"
        if err == "ReferencePanics" or (st[i] == 104 and ovf):
            continue
        ok = (got == err) if want is None else (st[i] == 0 and np.array_equal(outs[i], want))
        if ok and want is not None and a.native:
            ok = nats[i] is not None and np.array_equal(np.asarray(nats[i]).reshape(-1), np.asarray(img.pixels).reshape(-1))
            if not ok:
                got = "native variant differs"
        if not ok:
            bad += 1
            p = os.path.join(a.out, f"seed{seed}_case{i}.jpg")
            open(p, "wb").write(d)
            print(f"MISMATCH seed {seed} case {i}: oracle {err} gpu {got} overflow {ovf} -> {p}", flush=True)
print(f"{total} cases, {bad} mismatches")
