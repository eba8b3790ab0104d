"""End-to-end (host buffers in, pinned host RGBA out) time of zpx_decode_batch_rgba on cfg2 for several
pipeline settings, on one box in one process (A/B without box-to-box variance).

  python tools/e2e_bench.py [--n 1024] [--steps 3] [--chunks 64,128,256] [--ramp 0,1]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth_jpeg as S  # noqa: E402
from zpix_b200 import jpeg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--distinct", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--chunks", default="64,128,256")
ap.add_argument("--ramp", default="0,1")
ap.add_argument("--workers", default="2,3,4")
a = ap.parse_args()
W, H = 1920, 1080
base = S.make_batch(2, a.distinct, W, H, cache_dir="/tmp/zpx_synth", mode="YCbCr", subsampling="4:2:0", restart_rows=1)
datas = [base[i % a.distinct] for i in range(a.n)]
lib = jpeg.lib
out_bytes = 4 * W * H
pinned = lib.zpx_host_alloc(out_bytes * a.n)
assert pinned
outs = (C.c_void_p * a.n)(*[pinned + i * out_bytes for i in range(a.n)])
keep = [np.frombuffer(d, np.uint8) for d in datas]
ptrs = (C.c_void_p * a.n)(*[x.ctypes.data for x in keep])
lens = (C.c_size_t * a.n)(*[x.size for x in keep])
st = (C.c_int32 * a.n)()
for chunk in [int(x) for x in a.chunks.split(",")]:
  for workers in [int(x) for x in a.workers.split(",")]:
    for ramp in [int(x) for x in a.ramp.split(",")]:
        ctx = jpeg.Context([0])
        ctx.set_option(4, chunk)
        ctx.set_option(6, ramp)
        ctx.set_option(7, workers)
        for _ in range(2):
            assert lib.zpx_decode_batch_rgba(ctx.handle, ptrs, lens, a.n, outs, None, st) == 0
        ts = []
        for _ in range(a.steps):
            t0 = time.perf_counter()
            assert lib.zpx_decode_batch_rgba(ctx.handle, ptrs, lens, a.n, outs, None, st) == 0
            ts.append(1e3 * (time.perf_counter() - t0))
        print(json.dumps({"chunk": chunk, "workers": workers, "ramp": ramp, "ms_min": round(min(ts), 2), "ms_mean": round(sum(ts) / len(ts), 2),
                          "Gpix_s": round(a.n * W * H / 1e9 / (min(ts) / 1e3), 2)}), flush=True)
        ctx.close()
lib.zpx_host_free(pinned)
