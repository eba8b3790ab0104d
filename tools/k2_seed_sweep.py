"""Seed sweep of the block-level K2 test (tests/test_gpu_blocks.py) on one GPU: every sampling x sizes x seeds, with the
sparse-block IDCT on and off; prints every mismatch with its first coordinates.
    python tools/k2_seed_sweep.py [--seeds 20] [--names 411 420 ...]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--names", nargs="*", default=None)
    ap.add_argument("--shuffle", type=int, default=0, help="non-zero: run the cases in a random order (context reuse across shapes)")
    ap.add_argument("--kinds", nargs="*", default=["sparse", "dense", "lo4_mixed"])
    a = ap.parse_args()
    import test_gpu_blocks as T
    from zpix_b200 import jpeg

    ctx = jpeg.Context([0])
    names = a.names or [n for n in sorted(T.SAMPLINGS)]
    bad = runs = 0
    jobs = [(name, seed, wh, kind) for name in names for seed in range(a.seeds)
            for wh in ((256, 64), (640, 32), (253, 61), (36, 130)) for kind in a.kinds]
    if a.shuffle:
        np.random.default_rng(a.shuffle).shuffle(jobs)
    for name, seed, (width, height), kind in jobs:
        mode, comp_hv = T.SAMPLINGS[name]
        rng = np.random.default_rng([seed, width, len(kind)])
        if True:
            if True:
                if True:
                    mxx, myy = T._geometry(width, height, comp_hv)
                    n = mxx * myy * sum(h * v for h, v in comp_hv)
                    blocks = T._random_blocks(rng, n, kind)
                    quant = rng.integers(1, 64, (len(comp_hv), 64))
                    _, rgba = T._expected(width, height, comp_hv, quant, mode, blocks)
                    for dense in (0, 1):
                        ctx.set_option(11, dense)
                        got, _, _ = T._run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, False)
                        ctx.set_option(11, 0)
                        runs += 1
                        if not np.array_equal(got, rgba):
                            bad += 1
                            w = np.argwhere(got.reshape(rgba.shape) != rgba)
                            print("MISMATCH", name, "seed", seed, (width, height), kind, "dense_only", dense, "count", len(w),
                                  "first", w[:4].tolist(), "last", w[-1].tolist(), flush=True)
    print(f"k2 seed sweep: {runs} runs, {bad} mismatches")
    ctx.close()


if __name__ == "__main__":
    main()
