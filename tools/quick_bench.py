"""Developer timing probe (not the benchmark contract; see bench.py)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from tools import synth_jpeg as S
from zpix_b200 import jpeg

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--distinct", type=int, default=64)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--sub", default="4:2:0")
ap.add_argument("--mode", default="YCbCr")
ap.add_argument("--dri", type=int, default=1)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--generic", type=int, default=0)
ap.add_argument("--emode", type=int, default=0)
ap.add_argument("--subb", type=int, default=0)
ap.add_argument("--prog", type=int, default=0)
a = ap.parse_args()

print("cpus", os.cpu_count())
t = time.time()
kw = dict(mode=a.mode)
if a.mode == "YCbCr":
    kw["subsampling"] = a.sub
if a.dri:
    kw["restart_rows"] = a.dri
if a.prog:
    kw["progressive"] = True
base = S.make_batch(2, a.distinct, a.w, a.h, cache_dir="/tmp/zpx_synth", **kw)
datas = [base[i % a.distinct] for i in range(a.n)]
print(f"synth {a.distinct} images in {time.time()-t:.1f}s, avg {sum(map(len, base))/len(base):.0f} B")
ctx = jpeg.Context([0])
if a.generic:
    ctx.set_option(2, 1)
ctx.set_option(1, a.emode)
ctx.set_option(3, a.subb)
t = time.time()
b = jpeg.Batch(ctx, datas)
t1 = time.time()
b.upload()
t2 = time.time()
print(f"open {1e3*(t1-t):.1f} ms  upload {1e3*(t2-t1):.1f} ms")
for it in range(a.iters):
    t = time.time()
    b.decode()
    tm = b.timing(0)
    wall = time.time() - t
    mp = tm["pixels"] / 1e6
    print(f"iter {it}: wall {1e3*wall:.2f} ms entropy {tm['entropy_ms']:.3f} ms idct {tm['idct_ms']:.3f} ms (fused {tm['idct_fused_ms']:.3f}) total {tm['total_ms']:.3f} ms "
          f"-> {mp/ (tm['total_ms']/1e3)/1e3:.2f} Gpix/s ; K2 {tm['idct_fused_bytes']/1e9/(max(tm['idct_fused_ms'],1e-6)/1e3):.0f} GB/s ; K1 in {tm['entropy_bytes_in']/1e9/(tm['entropy_ms']/1e3):.1f} GB/s")
t = time.time()
outs, st = b.fetch_rgba()
print(f"fetch {1e3*(time.time()-t):.1f} ms, failed {sum(1 for s in st if s)}")
b.close()
