"""K1 timing probe for kernel experiments (not the benchmark): entropy-stage time (k0 + k1) of a cfg2 batch (512 x 1080p
4:2:0, DRI = one MCU row) and of a 2160p 4:2:2 batch, best of 6, with a parity check of the first images."""
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
work = [("cfg2 420", S.make_batch(2, 64, 1920, 1080, cache_dir=CACHE, subsampling="4:2:0", restart_rows=1), 1024),
        ("422 2160", S.make_batch(4, 16, 3840, 2160, cache_dir=CACHE, mode="YCbCr", subsampling="4:2:2", restart_rows=1), 256)]
for name, base, n in work:
    datas = [base[i % len(base)] for i in range(n)]
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        best = None
        for _ in range(6):
            b.decode()
            t = b.timing(0)
            if best is None or t["entropy_ms"] < best["entropy_ms"]:
                best = t
        outs, st = b.fetch_rgba()
    bad = sum(0 if np.array_equal(outs[i].reshape(-1), O.decode(datas[i]).rgbaPixels().reshape(-1)) else 1 for i in range(4))
    print(f"{name:10s} n={n:5d} entropy {best['entropy_ms']:.3f} ms  {best['entropy_bytes_in'] / 1e9 / (best['entropy_ms'] / 1e3):6.1f} GB/s in  "
          f"failed {sum(1 for s in st if s)} mismatch(4) {bad}", flush=True)
ctx.close()
