"""Smallest end-to-end batch for compute-sanitizer memcheck: every kernel family once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zpix_b200 import jpeg
fx = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_fixtures")
names = ["video-001.q50.420.jpeg", "video-001.restart2.jpeg", "video-005.gray.jpeg", "video-001.cmyk.jpeg",
         "video-001.q50.420.progressive.jpeg", "video-001.rgb.jpeg", "video-001.q50.411.jpeg"]
datas = [open(os.path.join(fx, n), "rb").read() for n in names]
for mode in (1, 2):
    ctx = jpeg.Context([0])
    ctx.set_option(1, mode)
    ctx.set_option(3, 32)
    res = jpeg.decodeBatch(datas, ctx, raise_on_error=True)
    print("mode", mode, "ok", len(res))
    ctx.close()
