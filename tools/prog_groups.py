import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from tools import synth_jpeg as S
from zpix_b200 import jpeg
base = S.make_batch(5, 16, 1920, 1080, cache_dir=bench.CACHE, first=4000, mode="YCbCr", subsampling="4:2:0", progressive=True)
for n in (256, 2048):
    datas = [base[i % 16] for i in range(n)]
    for k in (1, 2, 4, 8):
        try:
            ctx = jpeg.Context([0] * k)
        except Exception as e:
            print("ctx", k, "failed", e); continue
        with jpeg.Batch(ctx, datas) as b:
            b.upload()
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                b.decode()
                st = b.status()
                wall = 1e3 * (time.perf_counter() - t0)
                tms = [b.timing(d)["total_ms"] for d in range(k)]
                best = wall if best is None else min(best, wall)
            failed = sum(1 for s in st if s)
        print(json.dumps({"images": n, "virtual_devices": k, "wall_ms": round(best, 2), "per_dev_ms": [round(t, 1) for t in tms], "failed": failed,
                          "Gpix_s": round(n * 1920 * 1080 / 1e6 / best, 1)}), flush=True)
        ctx.close()
