"""cfg1: one image (the reference's iceberg.jpg, 2048x2048 4:4:4, no DRI, optimised tables) through the drop-in
`jpeg.loadFromBuffer` mirror and through the batch API, next to the CPU restatement.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from zpix_b200 import jpeg  # noqa: E402

d = open(os.path.join(ROOT, "tests", "golden", "ref_fixtures", "iceberg.jpg"), "rb").read()
ctx = jpeg.Context([0])
best_dev, best_call = 1e9, 1e9
with jpeg.Batch(ctx, [d]) as b:
    b.upload()
    for _ in range(8):
        b.decode()
        best_dev = min(best_dev, b.timing(0)["total_ms"])
    tm = b.timing(0)
for _ in range(5):
    t0 = time.perf_counter()
    img = jpeg.loadFromBuffer(d, ctx)
    px = img.rgbaPixels()
    best_call = min(best_call, 1e3 * (time.perf_counter() - t0))
t0 = time.perf_counter()
want = O.decode(d).rgbaPixels()
cpu_ms = 1e3 * (time.perf_counter() - t0)
assert np.array_equal(np.asarray(px).reshape(want.shape), want)
print(json.dumps({"workload": "cfg1 iceberg.jpg 2048x2048 4:4:4 no DRI", "device_ms": round(best_dev, 3),
                  "entropy_ms": round(tm["entropy_ms"], 3), "idct_ms": round(tm["idct_ms"], 3),
                  "loadFromBuffer_ms": round(best_call, 2), "cpu_restatement_ms": round(cpu_ms, 1), "bit_exact": True}))
