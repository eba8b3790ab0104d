"""Pinned host<->device copy bandwidth of this box (the floor of the end-to-end number).

  python tools/pcie_bw.py [--mb 1024] [--iters 5]
"""
import argparse
import json

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1024)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
n = a.mb << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn):
    best = 0.0
    for _ in range(a.iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


out = {
    "h2d_GBps": timed(lambda: d.copy_(h, non_blocking=True)),
    "d2h_GBps": timed(lambda: h.copy_(d, non_blocking=True)),
    "duplex_each_GBps": timed(both),
    "bytes": n,
}
print(json.dumps(out))
