"""Progressive (SOF2) throughput on one GPU: n x 1920x1080 4:2:0 progressive files (16 distinct, libjpeg's 10-scan
script), device-resident, both progressive kernels (ZPX_OPT_PROGRESSIVE_MODE 0 = one lane per scan, 1 = one warp per
scan), outputs compared with each other and (distinct files) with the CPU oracle.
    python tools/prog_bench.py [--n 256 2048] [--skip-warp]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[256, 2048])
    ap.add_argument("--skip-warp", action="store_true")
    ap.add_argument("--size", type=int, nargs=2, default=[1920, 1080])
    a = ap.parse_args()
    import bench
    from tools import synth_jpeg as S
    from zpix_b200 import jpeg

    w, h = a.size
    base = S.make_batch(5, 16, w, h, cache_dir=bench.CACHE, first=4000, mode="YCbCr", subsampling="4:2:0", progressive=True)
    print(json.dumps({"distinct": len(base), "bytes_per_file": int(np.mean([len(x) for x in base]))}), flush=True)
    for n in a.n:
        datas = [base[i % len(base)] for i in range(n)]
        res = {}
        for mode in ([0] if a.skip_warp else [0, 1]):
            ctx = jpeg.Context([0])
            ctx.set_option(10, mode)
            with jpeg.Batch(ctx, datas) as b:
                b.upload()
                best = None
                for _ in range(3):
                    b.decode()
                    tm = b.timing(0)
                    if best is None or tm["total_ms"] < best["total_ms"]:
                        best = tm
                st = list(b.status())
                failed = sum(1 for s in st if s)
                if failed:
                    from collections import Counter
                    print(json.dumps({"status_histogram": Counter(st).most_common(6), "first_failed": [i for i, s in enumerate(st) if s][:8]}), flush=True)
                par = bench.parity_device(b, datas, len(base))
            ctx.close()
            res[mode] = {"mode": "lane per scan" if mode == 0 else "warp per scan", "images": n, "failed": failed,
                         "entropy_ms": round(best["entropy_ms"], 3), "idct_ms": round(best["idct_ms"], 3),
                         "total_ms": round(best["total_ms"], 3), "launches": best["entropy_launches"],
                         "Mpixels_s": round(best["pixels"] / 1e6 / (best["total_ms"] / 1e3), 1), "parity": par}
            print(json.dumps(res[mode]), flush=True)


if __name__ == "__main__":
    main()
