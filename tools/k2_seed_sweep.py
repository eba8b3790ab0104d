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
    ap.add_argument("--like-test", type=float, default=0, help="seconds to spend replaying the edge-tile test with fresh seeds")
    ap.add_argument("--shuffle", type=int, default=0, help="non-zero: run the cases in a random order (context reuse across shapes)")
    ap.add_argument("--kinds", nargs="*", default=["sparse", "dense", "lo4_mixed", "r6_mixed"])
    a = ap.parse_args()
    import test_gpu_blocks as T
    from zpix_b200 import jpeg

    ctx = jpeg.Context([0])
    bad = runs = 0
    names = a.names or [n for n in sorted(T.SAMPLINGS)]
    if a.like_test:
        # the sequence of tests/test_gpu_blocks.py::test_every_sampling_interior_and_edge_tiles, seed after seed: every
        # sampling in sorted order, four sizes from ONE generator, planes fetched too, one shared context
        import time
        t_end = time.time() + a.like_test
        seed = 0
        while time.time() < t_end:
            seed += 1
            for name in sorted(T.SAMPLINGS):
                rng = np.random.default_rng([seed, len(name)])
                mode, comp_hv = T.SAMPLINGS[name]
                for width, height in ((256, 64), (640, 32), (253, 61), (36, 130)):
                    mxx, myy = T._geometry(width, height, comp_hv)
                    n = mxx * myy * sum(h * v for h, v in comp_hv)
                    blocks = T._random_blocks(rng, n, "dense" if width == 256 else "sparse")
                    quant = rng.integers(1, 64, (len(comp_hv), 64))
                    planes, rgba = T._expected(width, height, comp_hv, quant, mode, blocks)
                    got, nat, _ = T._run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, mode in (T.MODE_GRAY, T.MODE_YCBCR))
                    runs += 1
                    ok = np.array_equal(got, rgba)
                    if nat is not None:
                        ref = np.concatenate([p.reshape(-1) for p in planes[:1 if mode == T.MODE_GRAY else 3]])
                        ok = ok and np.array_equal(nat[:ref.size], ref)
                    if not ok:
                        bad += 1
                        w = np.argwhere(got.reshape(rgba.shape) != rgba)
                        print("MISMATCH like-test", name, "seed", seed, (width, height), "rgba diffs", len(w), w[:4].tolist(), flush=True)
                        np.savez(f"gpurun_out/k2_mismatch_{name}_{seed}_{width}.npz", blocks=blocks, quant=quant, got=got, want=rgba)
        print(f"k2 seed sweep (like the test): {seed} seeds, {runs} runs, {bad} mismatches")
        ctx.close()
        return
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
