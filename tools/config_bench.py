"""Device-resident throughput of the other BASELINE.json configurations (cfg3, cfg4, cfg5) on one GPU.

bench.py measures the headline configuration (cfg2); this tool records the others for DESIGN.md /
profiles/.  One JSON line per workload: Mpixel/s with inputs resident in HBM (CUDA events around the
whole decode), the entropy and IDCT/colour stage times, K2's algorithmic GB/s.

  python tools/config_bench.py [--only cfg3,cfg4,...] [--scale 1.0] [--iters 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth_jpeg as S  # noqa: E402
from zpix_b200 import jpeg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--scale", type=float, default=1.0, help="multiplies every batch size")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--distinct", type=int, default=32)
a = ap.parse_args()
CACHE = "/tmp/zpx_synth"


def rep(base, n):
    return [base[i % len(base)] for i in range(n)]


def cfg3(n):
    d = min(a.distinct, n // 2)
    g = S.make_batch(3, d, 512, 512, cache_dir=CACHE, mode="L")
    c = S.make_batch(3, d, 512, 512, cache_dir=CACHE, first=5000, mode="YCbCr", subsampling="4:4:4")
    return rep(g, n // 2) + rep(c, n - n // 2)


def cfg4(n, dri):
    d = min(a.distinct // 2, n)
    kw = dict(mode="YCbCr", subsampling="4:2:2")
    if dri:
        kw["restart_rows"] = 1
    return rep(S.make_batch(4, d, 3840, 2160, cache_dir=CACHE, **kw), n)


def cfg5(n):
    d = min(a.distinct // 2, max(1, n // 4))
    cmyk = S.make_batch(5, d, 1920, 1080, cache_dir=CACHE, mode="CMYK")
    ycck = S.make_batch(5, d, 1920, 1080, cache_dir=CACHE, first=2000, mode="CMYK", ycck=True)
    prog = S.make_batch(5, d, 1920, 1080, cache_dir=CACHE, first=4000, mode="YCbCr", subsampling="4:2:0", progressive=True)
    return rep(cmyk, n // 4) + rep(ycck, n // 4) + rep(prog, n - 2 * (n // 4))


def cfg5_prog(n):
    d = min(a.distinct, n)
    return rep(S.make_batch(5, d, 1920, 1080, cache_dir=CACHE, first=4000, mode="YCbCr", subsampling="4:2:0", progressive=True), n)


WORK = {
    "cfg3": ("4096 x 512x512, half gray + half 4:4:4, baseline, no DRI", lambda s: cfg3(int(4096 * s))),
    "cfg4_dri": ("512 x 3840x2160 4:2:2, baseline, DRI = one MCU row", lambda s: cfg4(int(512 * s), True)),
    "cfg4_nodri": ("512 x 3840x2160 4:2:2, baseline, no DRI", lambda s: cfg4(int(512 * s), False)),
    "cfg5_mixed": ("512 x 1920x1080: 1/4 Adobe CMYK, 1/4 YCbCrK, 1/2 progressive 4:2:0", lambda s: cfg5(int(512 * s))),
    "cfg5_progressive_2048": ("2048 x 1920x1080 progressive 4:2:0 (10 scans)", lambda s: cfg5_prog(int(2048 * s))),
}

only = [x for x in a.only.split(",") if x]
ctx = jpeg.Context([0])
for name, (desc, make) in WORK.items():
    if only and name not in only:
        continue
    t0 = time.time()
    datas = make(a.scale)
    t_synth = time.time() - t0
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        best = None
        for _ in range(a.iters):
            b.decode()
            tm = b.timing(0)
            if best is None or tm["total_ms"] < best["total_ms"]:
                best = tm
        st = b.status()
    bad = sum(1 for s in st if s)
    line = {
        "workload": name, "desc": desc, "images": len(datas), "failed": bad,
        "Mpixels_per_s": round(best["pixels"] / 1e6 / (best["total_ms"] / 1e3), 1),
        "total_ms": round(best["total_ms"], 3), "entropy_ms": round(best["entropy_ms"], 3), "idct_ms": round(best["idct_ms"], 3),
        "k2_fused_ms": round(best["idct_fused_ms"], 3),
        "k2_fused_GBps": round(best["idct_fused_bytes"] / 1e9 / (max(best["idct_fused_ms"], 1e-6) / 1e3), 1) if best["idct_fused_bytes"] else None,
        "entropy_in_GBps": round(best["entropy_bytes_in"] / 1e9 / (best["entropy_ms"] / 1e3), 2),
        "compressed_MB": round(sum(map(len, datas)) / 1e6, 1), "synth_s": round(t_synth, 1),
    }
    print(json.dumps(line), flush=True)
ctx.close()
