"""K2 timing probe for kernel experiments (not the benchmark): fused-kernel time of a cfg2 batch (256 x 1080p 4:2:0),
a 4:4:4 batch (1024 x 512x512) and a 4:2:2 batch (64 x 2160p), best of 5, with a parity check of the first images."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools import synth_jpeg as S  # noqa: E402
from zpix_b200 import jpeg  # noqa: E402

CACHE = os.environ.get("ZPX_SYNTH_CACHE", "/tmp/zpx_synth")
ctx = jpeg.Context([0])
work = [("cfg2 420", S.make_batch(2, 32, 1920, 1080, cache_dir=CACHE, subsampling="4:2:0", restart_rows=1), 512),
        ("444 512", S.make_batch(3, 32, 512, 512, cache_dir=CACHE, first=5000, mode="YCbCr", subsampling="4:4:4"), 2048),
        ("gray 512", S.make_batch(3, 32, 512, 512, cache_dir=CACHE, mode="L"), 2048),
        ("422 2160", S.make_batch(4, 16, 3840, 2160, cache_dir=CACHE, mode="YCbCr", subsampling="4:2:2", restart_rows=1), 256)]
for name, base, n in work:
    datas = [base[i % len(base)] for i in range(n)]
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        best = None
        for _ in range(6):
            b.decode()
            t = b.timing(0)
            if best is None or t["idct_fused_ms"] < best["idct_fused_ms"]:
                best = t
        outs, st = b.fetch_rgba()
    bad = sum(0 if np.array_equal(outs[i].reshape(-1), O.decode(datas[i]).rgbaPixels().reshape(-1)) else 1 for i in range(4))
    print(f"{name:10s} n={n:5d} k2 {best['idct_fused_ms']:.3f} ms  {best['idct_fused_bytes'] / 1e9 / (best['idct_fused_ms'] / 1e3):7.0f} GB/s  "
          f"failed {sum(1 for s in st if s)} mismatch(4) {bad}", flush=True)
ctx.close()
