"""Block-level property tests of the reconstruction kernels (K2), fed directly through the test hooks of the C ABI
(zpx_batch_open_synthetic / zpx_batch_set_coefficients / zpx_test_colour) instead of through encoded files:

  * random int16 coefficient blocks x random 8- and 16-bit quantisers -> reconstructBlock (decoder.zig:1553-1634,
    idct.zig:77-201) of the oracle, for gray and every chroma sampling the fused kernel takes, fused and unfused,
    inside and outside the range [-4096, 4095] that selects the kernels' exact all-AC-zero-row variant;
  * every (Y, Cb, Cr) triple (2^24) and 2^20 CMYK / YCbCrK samples -> Color.toRGBA >> 8 (color.zig:90-121);
  * every (Cb, Cr) pair through the real fused kernels as DC-only blocks.
Bit-exact, no tolerance."""
import os
import zlib

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

MODE_GRAY, MODE_YCBCR, MODE_RGB, MODE_CMYK, MODE_YCCK = range(5)
UNZIG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                  21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60,
                  61, 54, 47, 55, 62, 63])


@pytest.fixture(scope="module")
def jpeg():
    from zpix_b200 import jpeg as J

    return J


@pytest.fixture(scope="module")
def ctx(jpeg):
    c = jpeg.Context([0])
    yield c
    c.close()


def _geometry(width, height, comp_hv):
    h0, v0 = comp_hv[0]
    mxx = -(-width // (8 * h0))
    myy = -(-height // (8 * v0))
    return mxx, myy


def _block_order(mxx, myy, comp_hv):
    """(component, bx, by) of every block in the order zpx_batch_fetch_coefficients uses (MCU-interleaved)."""
    out = []
    for my in range(myy):
        for mx in range(mxx):
            for c, (h, v) in enumerate(comp_hv):
                for j in range(h * v):
                    out.append((c, h * mx + j % h, v * my + j // h))
    return out


def _expected(width, height, comp_hv, quant_zz, mode, blocks):
    """planes (makeImg geometry) and RGBA from the oracle's reconstructBlock / toRGBA, assembled in numpy"""
    mxx, myy = _geometry(width, height, comp_hv)
    order = _block_order(mxx, myy, comp_hv)
    comp = np.array([o[0] for o in order])
    planes = []
    for c, (h, v) in enumerate(comp_hv):
        idx = np.nonzero(comp == c)[0]
        px = O.reconstruct_blocks(blocks[idx].astype(np.int32), quant_zz[c])
        pl = np.zeros((8 * v * myy, 8 * h * mxx), np.uint8)
        for k, i in enumerate(idx):
            _, bx, by = order[i]
            pl[8 * by:8 * by + 8, 8 * bx:8 * bx + 8] = px[k]
        planes.append(pl)
    ys, xs = np.mgrid[0:height, 0:width]
    if mode == MODE_GRAY:
        g = planes[0][:height, :width]
        rgba = np.stack([g, g, g, np.full_like(g, 255)], axis=-1)
    else:
        h0, v0 = comp_hv[0]
        hr, vr = h0 // comp_hv[1][0], v0 // comp_hv[1][1]
        Y = planes[0][ys, xs]
        C1 = planes[1][ys // vr, xs // hr]
        C2 = planes[2][ys // vr, xs // hr]
        if mode == MODE_RGB:
            rgba = np.stack([Y, C1, C2, np.full_like(Y, 255)], axis=-1)
        elif mode == MODE_YCBCR:
            rgba = O.ycbcr_to_rgba8_batch(np.stack([Y, C1, C2], axis=-1).reshape(-1, 3)).reshape(height, width, 4)
        elif mode == MODE_CMYK:
            s = []
            for t in range(4):
                sub = comp_hv[t] != comp_hv[0]
                s.append(255 - planes[t][ys >> 1, xs >> 1] if sub else 255 - planes[t][ys, xs])
            rgba = O.cmyk_to_rgba8_batch(np.stack(s, axis=-1).reshape(-1, 4)).reshape(height, width, 4)
        else:  # YCbCrK by intent: RGB of the YCbCr planes in C, M, Y; K = 255 - black plane
            rgb = O.ycbcr_to_rgba8_batch(np.stack([Y, C1, C2], axis=-1).reshape(-1, 3)).reshape(height, width, 4)
            K = 255 - planes[3][ys, xs]
            rgba = O.cmyk_to_rgba8_batch(np.concatenate([rgb[..., :3], K[..., None]], axis=-1).reshape(-1, 4)).reshape(height, width, 4)
    return planes, rgba


def _random_blocks(rng, n, kind):
    b = np.zeros((n, 64), np.int64)
    if kind == "sparse":      # what real streams look like: a few small coefficients at low frequencies
        nz = rng.integers(0, 12, n)
        for i in range(n):
            pos = UNZIG[rng.integers(0, 20, nz[i])]
            b[i, pos] = rng.integers(-60, 61, nz[i])
        b[:, 0] = rng.integers(-1024, 1024, n)
    elif kind == "dense":     # every coefficient set, full conforming range
        b = rng.integers(-1023, 1024, (n, 64))
        b[:, 0] = rng.integers(-2047, 2048, n)
    elif kind == "edge":      # the corners of the range the fast IDCT variant takes
        b = rng.choice(np.array([-4096, -4095, -1, 0, 1, 4095]), (n, 64), p=[.1, .1, .1, .5, .1, .1])
    elif kind == "wide":      # outside [-4096, 4095]: the kernels' exact all-AC-zero-row variant
        b = rng.integers(-32768, 32768, (n, 64))
        b[rng.random((n, 64)) < 0.6] = 0
    elif kind == "dc_rows":   # rows whose only non-zero value sits in column 0 (the reference's row shortcut), huge
        b[:, ::8] = rng.integers(-32768, 32768, (n, 8))
        b[rng.random(n) < 0.5, 8:] = 0
    elif kind in ("lo4", "lo4_mixed"):
        # nothing outside the top-left 4x4 corner: warps made of such blocks take the fused kernel's sparse IDCT;
        # "lo4_mixed" puts one coefficient just outside the corner into a few blocks, so that warps of both kinds
        # (and warps that fall back because of a single block) occur in one frame
        corner = np.array([8 * r + c for r in range(4) for c in range(4)])
        b[:, corner] = rng.choice(np.array([-4096, -1023, -60, -1, 0, 1, 60, 1023, 4095]), (n, 16), p=[.05, .1, .15, .1, .2, .1, .15, .1, .05])
        if kind == "lo4_mixed":
            hit = np.nonzero(rng.random(n) < 0.02)[0]
            b[hit, rng.choice(np.array([4, 32, 39, 60, 63, 12, 33]), len(hit))] = rng.choice(np.array([-1, 1, 300]), len(hit))
    elif kind == "r7":        # only coefficient row 7 zero (the gray kernel's variant), a few blocks with a value there
        b[:, :56] = rng.integers(-1023, 1024, (n, 56))
        hit = np.nonzero(rng.random(n) < 0.01)[0]
        b[hit, rng.choice(np.array([56, 60, 63]), len(hit))] = rng.choice(np.array([-1, 1, 300]), len(hit))
    elif kind in ("r6", "r6_mixed"):
        # coefficient rows 6 and 7 zero (warps of such blocks skip those rows in the fused kernel); "r6_mixed": a few blocks
        # with a value in row 6 or 7, and a third of the blocks with nothing outside the 4x4 corner
        b[:, :48] = rng.choice(np.array([-4096, -1023, -60, -1, 0, 1, 60, 1023, 4095]), (n, 48), p=[.03, .07, .15, .1, .3, .1, .15, .07, .03])
        if kind == "r6_mixed":
            hit = np.nonzero(rng.random(n) < 0.02)[0]
            b[hit, rng.choice(np.array([48, 55, 56, 63, 59]), len(hit))] = rng.choice(np.array([-1, 1, 300]), len(hit))
            lo = np.nonzero(rng.random(n) < 0.33)[0]
            keep = np.zeros(64, bool)
            keep[[8 * r + c for r in range(4) for c in range(4)]] = True
            b[np.ix_(lo, np.nonzero(~keep)[0])] = 0
    return b.astype(np.int16)


SAMPLINGS = {
    "gray": (MODE_GRAY, [(1, 1)]),
    "444": (MODE_YCBCR, [(1, 1), (1, 1), (1, 1)]),
    "422": (MODE_YCBCR, [(2, 1), (1, 1), (1, 1)]),
    "420": (MODE_YCBCR, [(2, 2), (1, 1), (1, 1)]),
    "440": (MODE_YCBCR, [(1, 2), (1, 1), (1, 1)]),
    "411": (MODE_YCBCR, [(4, 1), (1, 1), (1, 1)]),
    "410": (MODE_YCBCR, [(4, 2), (1, 1), (1, 1)]),
    "221212": (MODE_YCBCR, [(2, 2), (1, 2), (1, 2)]),   # unfused kernels only
    "rgb": (MODE_RGB, [(1, 1), (1, 1), (1, 1)]),
    "rgb_2x2": (MODE_RGB, [(2, 2), (1, 1), (1, 1)]),
    "cmyk": (MODE_CMYK, [(1, 1), (1, 1), (1, 1), (1, 1)]),
    "cmyk_sub": (MODE_CMYK, [(2, 2), (1, 1), (1, 1), (2, 2)]),
    "ycck": (MODE_YCCK, [(1, 1), (1, 1), (1, 1), (1, 1)]),
}


def _run(jpeg, ctx, width, height, comp_hv, quant_zz, mode, blocks, generic, want_planes):
    ctx.set_option(2, 1 if generic else 0)
    ctx.set_option(9, 1)  # native planes beside the RGBA, from the fused kernel too
    try:
        with jpeg.SyntheticBatch(ctx, width, height, comp_hv, quant_zz, mode) as b:
            inf = b.info(0)
            b.upload()
            b.set_coefficients(blocks)
            b.decode()
            outs, stt = b.fetch_rgba()
            assert stt == [0]
            nat = b.fetch_native()[0][0] if want_planes else None
            back = b.coefficients(0)
        assert np.array_equal(back, blocks)   # the injection hook and the fetch hook are inverses
        return outs[0], nat, inf
    finally:
        ctx.set_option(2, 0)
        ctx.set_option(9, 0)


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 2**31 - 1), name=st.sampled_from(sorted(SAMPLINGS)),
       kind=st.sampled_from(["sparse", "dense", "edge", "wide", "dc_rows", "lo4", "lo4_mixed", "r6", "r6_mixed", "r7"]), q16=st.booleans(), dims=st.sampled_from([(64, 48), (150, 103), (257, 33), (33, 130)]))
def test_blocks_match_reconstruct_block(jpeg, ctx, seed, name, kind, q16, dims):
    """random blocks x random quantisers, every sampling, fused and unfused kernels, planes and RGBA"""
    rng = np.random.default_rng(seed)
    mode, comp_hv = SAMPLINGS[name]
    width, height = dims
    mxx, myy = _geometry(width, height, comp_hv)
    n = mxx * myy * sum(h * v for h, v in comp_hv)
    blocks = _random_blocks(rng, n, kind)
    hi = 65535 if q16 else 255
    quant = rng.integers(1, hi + 1, (len(comp_hv), 64))
    if rng.random() < 0.3:
        quant[:, :] = rng.integers(1, 4, (len(comp_hv), 64))  # near-lossless tables: large dequantised values stay exact
    planes, rgba = _expected(width, height, comp_hv, quant, mode, blocks)
    for generic in (False, True):
        got, nat, inf = _run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, generic, want_planes=mode in (MODE_GRAY, MODE_YCBCR))
        assert np.array_equal(got, rgba), (name, kind, q16, dims, generic, np.argwhere(got != rgba)[:3])
        if nat is not None:
            if mode == MODE_GRAY:
                assert np.array_equal(nat.reshape(planes[0].shape), planes[0])
            else:
                ylen = planes[0].size
                clen = planes[1].size
                assert np.array_equal(nat[:ylen].reshape(planes[0].shape), planes[0])
                assert np.array_equal(nat[ylen:ylen + clen].reshape(planes[1].shape), planes[1])
                assert np.array_equal(nat[ylen + clen:].reshape(planes[2].shape), planes[2])


def test_exact_row_variant_equals_fast_variant_in_range(jpeg, ctx):
    """blocks inside [-4096, 4095]: the exact all-AC-zero-row IDCT (taken by flagged images) and the fast one agree"""
    rng = np.random.default_rng(11)
    for name in ("gray", "420", "444", "422"):
        mode, comp_hv = SAMPLINGS[name]
        width, height = 200, 120
        mxx, myy = _geometry(width, height, comp_hv)
        n = mxx * myy * sum(h * v for h, v in comp_hv)
        blocks = _random_blocks(rng, n, "edge")
        quant = rng.integers(1, 256, (len(comp_hv), 64))
        _, rgba = _expected(width, height, comp_hv, quant, mode, blocks)
        for wide in (0, 1):
            ctx.set_option(8, wide)
            try:
                got, _, _ = _run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, False)
            finally:
                ctx.set_option(8, 0)
            assert np.array_equal(got, rgba), (name, wide)


@pytest.mark.parametrize("name", ["gray", "444", "422", "420", "440", "411", "410"])
def test_sparse_block_idct_equals_the_general_one(jpeg, ctx, name):
    """the fused kernel's sparse-block IDCTs (warps whose 32 blocks have zero coefficient rows 6 and 7 -- 4:2:0 --, a
    zero row 7 -- gray --, or nothing outside the top-left 4x4 corner -- the other samplings) against
    the oracle's reconstructBlock and against the same kernel with the variant switched off (ZPX_OPT_K2_DENSE)"""
    rng = np.random.default_rng(5 + len(name))
    mode, comp_hv = SAMPLINGS[name]
    for kind, (width, height) in (("lo4", (640, 64)), ("lo4_mixed", (1280, 48)), ("lo4_mixed", (333, 77)), ("r6", (640, 64)),
                                  ("r6_mixed", (1280, 48)), ("r6_mixed", (333, 77)), ("r7", (1280, 48))):
        mxx, myy = _geometry(width, height, comp_hv)
        n = mxx * myy * sum(h * v for h, v in comp_hv)
        blocks = _random_blocks(rng, n, kind)
        quant = rng.integers(1, 256, (len(comp_hv), 64))
        _, rgba = _expected(width, height, comp_hv, quant, mode, blocks)
        for dense in (0, 1):
            ctx.set_option(11, dense)
            try:
                got, _, _ = _run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, False)
            finally:
                ctx.set_option(11, 0)
            assert np.array_equal(got, rgba), (name, kind, width, dense, np.argwhere(got != rgba)[:3])


def test_every_ycbcr_triple(jpeg, ctx):
    """exhaustive: 2^24 (Y, Cb, Cr) -> RGBA8 == Color.toRGBA(.ycbcr) >> 8 (color.zig:90-114)"""
    v = np.arange(1 << 24, dtype=np.uint32)
    ycc = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8)
    got = jpeg.test_colour(ctx, MODE_YCBCR, ycc)
    want = O.ycbcr_to_rgba8_batch(ycc)
    assert np.array_equal(got, want)


def test_cmyk_and_ycck_sweeps(jpeg, ctx):
    """2^20 random CMYK samples plus the faces of the cube == Color.toRGBA(.cmyk) >> 8 (color.zig:115-121);
    YCbCrK by intent: RGB of (Y, Cb, Cr), K = 255 - plane, then the same formula"""
    rng = np.random.default_rng(3)
    s = rng.integers(0, 256, (1 << 20, 4)).astype(np.uint8)
    edge = np.array([[a, b, c, d] for a in (0, 1, 127, 128, 254, 255) for b in (0, 255, 128) for c in (0, 255, 77) for d in range(256)], np.uint8)
    s = np.concatenate([s, edge])
    assert np.array_equal(jpeg.test_colour(ctx, MODE_CMYK, s), O.cmyk_to_rgba8_batch(s))
    rgb = O.ycbcr_to_rgba8_batch(s[:, :3])
    want = O.cmyk_to_rgba8_batch(np.concatenate([rgb[:, :3], (255 - s[:, 3:4])], axis=-1))
    assert np.array_equal(jpeg.test_colour(ctx, MODE_YCCK, s), want)


@pytest.mark.parametrize("name", ["444", "420", "422"])
def test_every_chroma_pair_through_the_fused_kernel(jpeg, ctx, name):
    """DC-only blocks (quantiser 1): an 8x8 block of plane value P needs DC = 8 (P - 128).  One MCU per (Cb, Cr)
    pair, luma varying with the block: all 65536 chroma pairs go through the real fused kernel's colour phase."""
    mode, comp_hv = SAMPLINGS[name]
    h0, v0 = comp_hv[0]
    width, height = 256 * 8 * h0, 256 * 8 * v0
    mxx, myy = _geometry(width, height, comp_hv)
    assert (mxx, myy) == (256, 256)
    order = _block_order(mxx, myy, comp_hv)
    comp = np.array([o[0] for o in order])
    bx = np.array([o[1] for o in order])
    by = np.array([o[2] for o in order])
    plane_val = np.where(comp == 0, (bx * 37 + by * 101) & 255, np.where(comp == 1, bx, by))
    blocks = np.zeros((len(order), 64), np.int16)
    blocks[:, 0] = 8 * (plane_val - 128)
    quant = np.ones((3, 64), np.int64)
    _, rgba = _expected(width, height, comp_hv, quant, mode, blocks)
    got, _, _ = _run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, False)
    assert np.array_equal(got, rgba)
    # and the planes really hold the intended values: spot check through the colour formula
    assert tuple(got[0, 0]) == tuple(O.ycbcr_to_rgba8_batch(np.array([[0, 0, 0]], np.uint8))[0])


@pytest.mark.parametrize("name", sorted(SAMPLINGS))
def test_every_sampling_interior_and_edge_tiles(jpeg, ctx, name):
    """deterministic sweep (the hypothesis test above draws samplings at random): every sampling at a size whose tiles
    all lie inside the image (the fused kernel's bounds-test-free path) and at sizes with partial MCUs on the right and
    bottom edges, widths that are and are not multiples of four"""
    # (str hashes change from run to run; ZPX_TEST_SEED varies the draw on purpose)
    rng = np.random.default_rng([zlib.crc32(name.encode()) & 0xffff, int(os.environ.get("ZPX_TEST_SEED", "0"))])
    mode, comp_hv = SAMPLINGS[name]
    for width, height in ((256, 64), (640, 32), (253, 61), (36, 130)):
        mxx, myy = _geometry(width, height, comp_hv)
        n = mxx * myy * sum(h * v for h, v in comp_hv)
        blocks = _random_blocks(rng, n, "dense" if width == 256 else "sparse")
        quant = rng.integers(1, 64, (len(comp_hv), 64))
        planes, rgba = _expected(width, height, comp_hv, quant, mode, blocks)
        got, nat, _ = _run(jpeg, ctx, width, height, comp_hv, quant, mode, blocks, False, mode in (MODE_GRAY, MODE_YCBCR))
        assert np.array_equal(got, rgba), (name, width, height, np.argwhere(got != rgba)[:3])
        if nat is not None:
            ylen = planes[0].size
            assert np.array_equal(nat[:ylen].reshape(planes[0].shape), planes[0]), (name, width, height)
            if mode == MODE_YCBCR:
                clen = planes[1].size
                assert np.array_equal(nat[ylen:ylen + clen].reshape(planes[1].shape), planes[1])
                assert np.array_equal(nat[ylen + clen:].reshape(planes[2].shape), planes[2])
