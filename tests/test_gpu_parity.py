"""GPU parity tests: the CUDA path (through the C ABI, libzpixcuda.so) against the CPU oracle on the
same bytes.  Bit-exact: integer/byte work, no tolerance anywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from tools import synth_jpeg as S  # noqa: E402
from _damage import header_damage  # noqa: E402


@pytest.fixture(scope="module")
def jpeg():
    from zpix_b200 import jpeg as J  # fails loudly if libzpixcuda.so is missing

    return J


# entropy-decoder variants every parity test runs under:
#   auto   lane-per-interval when the batch has many restart intervals, else self-synchronising
#   lane   one lane per restart interval (serial inside the interval)
#   sub    self-synchronising sub-sequence decoder, default sub-sequence size
#   sub32  the same with 32-byte sub-sequences: many warps per segment, stresses cross-warp sweeps
ENTROPY_MODES = {"auto": (0, 0), "lane": (1, 0), "sub": (2, 0), "sub32": (2, 32)}


@pytest.fixture(scope="module", params=list(ENTROPY_MODES))
def ctx(jpeg, request):
    c = jpeg.Context()
    mode, sub = ENTROPY_MODES[request.param]
    c.set_option(1, mode)
    c.set_option(3, sub)
    yield c
    c.close()


def _read(d, name):
    with open(os.path.join(d, name), "rb") as f:
        return f.read()


def _oracle_rgba(data):
    """what the GPU path has to return: the oracle's pixels or error name (frames in which an End-Of-Band run crosses a
    scan boundary included: the library decodes those a second time, scan by scan with the run carried)"""
    try:
        img = O.decode(data)
    except O.OracleError as e:
        return None, e.name
    return img.rgbaPixels(), "ok"


def _gpu_batch(jpeg, ctx, datas):
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        b.decode()
        outs, st = b.fetch_rgba()
        t = b.timing(0)
    return outs, st, t


def _assert_same(jpeg, ctx, datas, names=None):
    outs, st, _ = _gpu_batch(jpeg, ctx, datas)
    for i, d in enumerate(datas):
        want, err = _oracle_rgba(d)
        overflow = O.last_coef_overflow()
        tag = names[i] if names else i
        if err == "ReferencePanics":
            continue  # the Zig code hits a panic / out-of-bounds index on this input: nothing defined to match
        if st[i] == 104:
            # documented deviation: coefficients are int16 in HBM; a (non-conforming) stream whose coefficients leave
            # that range is refused with CoefficientOutOfRange where the reference, keeping int32, decodes on
            assert overflow, f"{tag}: CoefficientOutOfRange but the oracle's coefficients fit int16"
            continue
        if want is None:
            assert st[i] != 0, f"{tag}: oracle fails with {err}, GPU path succeeded"
            assert jpeg.lib.zpx_error_name(st[i]).decode() == err, f"{tag}: {err} vs {st[i]}"
        else:
            assert st[i] == 0, f"{tag}: status {st[i]} {jpeg.lib.zpx_error_name(st[i]).decode()}"
            assert outs[i].shape == want.shape, tag
            if not np.array_equal(outs[i], want):
                bad = np.argwhere(outs[i] != want)
                raise AssertionError(f"{tag}: {len(bad)} bytes differ, first at {bad[0]}: gpu {outs[i][tuple(bad[0])]} want {want[tuple(bad[0])]}")


BASELINE_FIXTURES = [
    "video-001.jpeg", "video-001.q50.410.jpeg", "video-001.q50.411.jpeg", "video-001.q50.420.jpeg",
    "video-001.q50.422.jpeg", "video-001.q50.440.jpeg", "video-001.q50.444.jpeg", "video-005.gray.jpeg",
    "video-005.gray.q50.jpeg", "video-005.gray.q50.2x2.jpeg", "video-001.221212.jpeg", "video-001.restart2.jpeg",
    "video-001.rgb.jpeg", "video-001.cmyk.jpeg",
]


def test_reference_fixtures_one_batch(jpeg, ctx, fixtures_dir, golden_dir):
    """Every baseline fixture of the reference, in one mixed batch; RGBA == oracle == committed sha256."""
    datas = [_read(fixtures_dir, n) for n in BASELINE_FIXTURES]
    _assert_same(jpeg, ctx, datas, BASELINE_FIXTURES)
    outs, st, _ = _gpu_batch(jpeg, ctx, datas)
    gold = json.load(open(os.path.join(golden_dir, "rgba_sha256.json")))
    for n, o in zip(BASELINE_FIXTURES, outs):
        assert hashlib.sha256(o.tobytes()).hexdigest() == gold[n], n


def test_iceberg_cfg1(jpeg, ctx, fixtures_dir, golden_dir):
    """BASELINE.json configs[0]: iceberg.jpg, 2048x2048 4:4:4, no DRI, optimised Huffman tables."""
    data = _read(fixtures_dir, "iceberg.jpg")
    outs, st, _ = _gpu_batch(jpeg, ctx, [data])
    assert st == [0]
    gold = json.load(open(os.path.join(golden_dir, "rgba_sha256.json")))
    assert hashlib.sha256(outs[0].tobytes()).hexdigest() == gold["iceberg.jpg"]


def test_padded_rst_and_error_kinds(jpeg, ctx, fixtures_dir, golden_dir):
    """decoder.zig:2029-2279 through the GPU path: padded RST decodes; junk before RST1 is skipped
    (7 PASS splices) or raises BadRSTMarker (3 FAIL splices); truncations raise UnexpectedEof."""
    datas = [_read(golden_dir, "padded_rst_issue28717.jpg"), _read(golden_dir, "fuzz_issue10413.bin")]
    base = _read(fixtures_dir, "video-001.restart2.jpeg")
    for infix in [b"", b"\x00", b"\x61", b"\x61\x62\x63\xff\x00\x64", b"\xff", b"\xff\x00",
                  b"\xff\xff\xff\x00\xff\x00\x00\xff\xff\xff", b"\xff\x03", b"\xff\xd5", b"\xff\xff\xd5"]:
        datas.append(base[:2816] + infix + base[2816:])
    g = _read(fixtures_dir, "video-005.gray.q50.jpeg")
    i = g.index(b"\xff\xda") + 2
    datas += [g[:k] for k in range(i, i + 10)]
    datas.append(_read(fixtures_dir, "video-001.jpeg")[:24])
    # truncation in the middle of the entropy-coded data, and a file without EOI
    full = _read(fixtures_dir, "video-001.q50.420.jpeg")
    datas += [full[: len(full) // 2], full[:-2], full[:-1], b"", b"\xff", b"\x89PNG"]
    _assert_same(jpeg, ctx, datas)


def test_corrupt_image_inside_good_batch(jpeg, ctx, fixtures_dir):
    good = _read(fixtures_dir, "video-001.q50.420.jpeg")
    bad = bytearray(good)
    k = bad.index(b"\xff\xda") + 40
    bad[k:k + 8] = b"\xff\xd9\x00\x00\x00\x00\x00\x00"  # marker in the middle of the scan
    datas = [good, bytes(bad), good]
    _assert_same(jpeg, ctx, datas)


def test_coefficients_match_oracle_tap(jpeg, ctx, fixtures_dir):
    """K1 alone: int16 coefficient blocks == the oracle's pre-dequantisation blocks, scan order."""
    for name in ["video-001.q50.420.jpeg", "video-001.restart2.jpeg", "video-005.gray.jpeg", "video-001.jpeg"]:
        data = _read(fixtures_dir, name)
        _, recs = O.decode(data, tap=True)
        with jpeg.Batch(ctx, [data]) as b:
            b.upload()
            b.decode()
            co = b.coefficients(0)
        assert co.shape[0] == recs.shape[0], name
        assert np.array_equal(co.astype(np.int32), recs[:, 3:]), name


SYNTH_CASES = [
    # (w, h, kwargs)
    (1, 1, dict(subsampling="4:2:0")),
    (7, 9, dict(subsampling="4:4:4")),
    (17, 33, dict(subsampling="4:2:2")),
    (33, 17, dict(subsampling="4:2:0", restart_blocks=1)),
    (250, 131, dict(subsampling="4:2:0", restart_blocks=7)),
    (250, 131, dict(subsampling="4:2:0", restart_rows=1)),
    (250, 131, dict(subsampling="4:2:0", restart_blocks=5000)),
    (640, 480, dict(subsampling="4:2:0", restart_rows=1)),
    (640, 480, dict(subsampling="4:2:2", restart_rows=1)),
    (640, 480, dict(subsampling="4:4:4", restart_rows=2)),
    (641, 479, dict(subsampling="4:4:4")),
    (642, 478, dict(subsampling="4:2:0")),
    (643, 477, dict(subsampling="4:2:2")),
    (512, 512, dict(mode="L")),
    (513, 511, dict(mode="L", restart_rows=1)),
    (1920, 1080, dict(subsampling="4:2:0", restart_rows=1)),
    (3840, 2160, dict(subsampling="4:2:2", restart_rows=1)),
    (300, 200, dict(mode="CMYK")),
    (300, 200, dict(mode="CMYK", ycck=True)),
    (300, 200, dict(mode="CMYK", restart_rows=1)),
    (320, 240, dict(subsampling="4:2:0", quality=100)),
    (320, 240, dict(subsampling="4:4:4", quality=5)),
]


def test_synthetic_shapes(jpeg, ctx):
    """Odd sizes, every sampling Pillow emits, DRI in {1 block, 7, one row, > MCU count}, both table kinds."""
    datas, names = [], []
    for k, (w, h, kw) in enumerate(SYNTH_CASES):
        for seed in (k * 2, k * 2 + 1):  # even: Annex-K tables, odd: optimised tables
            datas.append(S.encode(90000 + seed, w, h, **kw))
            names.append(f"{w}x{h} {kw} seed{seed}")
    _assert_same(jpeg, ctx, datas, names)


def test_wide_gray_rows_cut_into_several_tiles(jpeg, ctx):
    """gray images wider than one tile of the fused kernel (128 blocks): the tiles of a row start at block columns
    that are not multiples of eight (the coefficient rows' swizzle key is the ABSOLUTE column) -- found by the
    planar-layout test of round 2: 1100 px = 138 blocks = 2 tiles of 69"""
    datas = [S.encode(62000 + i, w, h, mode="L", restart_rows=rr) for i, (w, h, rr) in
             enumerate([(1100, 33, 0), (1100, 40, 1), (2056, 24, 0), (3000, 17, 1), (1032, 64, 0), (4099, 9, 0)])]
    _assert_same(jpeg, ctx, datas)


def test_cfg2_sample(jpeg, ctx):
    """BASELINE.json configs[1] at reduced count: 1920x1080 4:2:0, DRI = one MCU row."""
    datas = S.make_batch(2, 16, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    _assert_same(jpeg, ctx, datas)


def test_cfg3_sample(jpeg, ctx):
    """configs[2] at reduced count: 512x512 gray + 4:4:4, no DRI."""
    datas = S.make_batch(3, 8, 512, 512, mode="L") + S.make_batch(3, 8, 512, 512, first=8, subsampling="4:4:4")
    _assert_same(jpeg, ctx, datas)


def test_cfg4_sample(jpeg, ctx):
    """configs[3] at reduced count: 3840x2160 4:2:2, with and without DRI."""
    datas = S.make_batch(4, 2, 3840, 2160, subsampling="4:2:2", restart_rows=1)
    datas += S.make_batch(4, 1, 3840, 2160, first=2, subsampling="4:2:2")
    _assert_same(jpeg, ctx, datas)


PROGRESSIVE_FIXTURES = [
    "video-001.progressive.jpeg", "video-001.q50.410.progressive.jpeg", "video-001.q50.411.progressive.jpeg",
    "video-001.q50.420.progressive.jpeg", "video-001.q50.422.progressive.jpeg", "video-001.q50.440.progressive.jpeg",
    "video-001.q50.444.progressive.jpeg", "video-005.gray.q50.progressive.jpeg", "video-005.gray.q50.2x2.progressive.jpeg",
    "video-001.separate.dc.progression.jpeg", "video-001.separate.dc.progression.progressive.jpeg",
]


def test_progressive_fixtures(jpeg, ctx, fixtures_dir, golden_dir):
    """SOF2 frames (spectral selection + successive approximation, 6-14 scans) on the GPU: RGBA == oracle
    == committed sha256, and == the baseline twin (the reference's own `decode + progressive` test)."""
    datas = [_read(fixtures_dir, n) for n in PROGRESSIVE_FIXTURES]
    _assert_same(jpeg, ctx, datas, PROGRESSIVE_FIXTURES)
    outs, st, _ = _gpu_batch(jpeg, ctx, datas)
    gold = json.load(open(os.path.join(golden_dir, "rgba_sha256.json")))
    for n, o in zip(PROGRESSIVE_FIXTURES, outs):
        assert hashlib.sha256(o.tobytes()).hexdigest() == gold[n], n
    pairs = [n for n in PROGRESSIVE_FIXTURES if n.endswith(".progressive.jpeg")]
    base, _, _ = _gpu_batch(jpeg, ctx, [_read(fixtures_dir, n.replace(".progressive", "")) for n in pairs])
    for n, b in zip(pairs, base):
        assert np.array_equal(b, outs[PROGRESSIVE_FIXTURES.index(n)]), n


def test_progressive_coefficients_match_oracle(jpeg, ctx, fixtures_dir):
    for name in ["video-001.q50.420.progressive.jpeg", "video-005.gray.q50.progressive.jpeg", "video-001.progressive.jpeg"]:
        data = _read(fixtures_dir, name)
        _, recs = O.decode(data, tap=True)
        with jpeg.Batch(ctx, [data]) as b:
            b.upload()
            b.decode()
            co = b.coefficients(0)
        assert co.shape[0] == recs.shape[0], name
        assert np.array_equal(co.astype(np.int32), recs[:, 3:]), name


def test_cfg5_sample(jpeg, ctx):
    """BASELINE.json configs[4] at reduced count: mixed batch of Adobe CMYK, YCbCrK (intent) and
    progressive 4:2:0 at 1920x1080, plus progressive gray / 4:4:4 / DRI variants at small sizes."""
    datas = S.make_batch(5, 2, 1920, 1080, mode="CMYK")
    datas += S.make_batch(5, 1, 1920, 1080, first=2, mode="CMYK", ycck=True)
    datas += S.make_batch(5, 3, 1920, 1080, first=3, subsampling="4:2:0", progressive=True)
    datas += [S.encode(50010, 333, 211, subsampling="4:4:4", progressive=True),
              S.encode(50011, 333, 211, mode="L", progressive=True),
              S.encode(50012, 333, 211, subsampling="4:2:2", progressive=True, restart_rows=1),
              S.encode(50013, 97, 64, subsampling="4:2:0", progressive=True, restart_blocks=3),
              S.encode(50014, 300, 200, mode="CMYK", progressive=True)]
    _assert_same(jpeg, ctx, datas)


def test_fused_equals_generic(jpeg, fixtures_dir):
    """The fused kernel and the unfused IDCT->planes->colour path give identical bytes."""
    datas = [_read(fixtures_dir, n) for n in BASELINE_FIXTURES] + S.make_batch(2, 2, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    c1, c2 = jpeg.Context(), jpeg.Context()
    c2.set_option(2, 1)
    a, sa, _ = _gpu_batch(jpeg, c1, datas)
    b, sb, _ = _gpu_batch(jpeg, c2, datas)
    assert sa == sb
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    c1.close()
    c2.close()


def _damaged(data: bytes, seed: int, n_trunc: int, n_flip: int):
    """Truncations and single-byte damage inside the entropy-coded part of `data`."""
    rng = np.random.default_rng(seed)
    first_sos = data.index(b"\xff\xda")
    out = []
    for k in rng.integers(first_sos + 14, len(data) - 2, n_trunc):
        out.append(data[: int(k)])
    for k in rng.integers(first_sos + 14, len(data) - 2, n_flip):
        d = bytearray(data)
        v = int(rng.integers(0, 255))
        if v == 0xFF or d[int(k)] == 0xFF or d[int(k) - 1] == 0xFF:
            continue  # keep the marker structure: those cases are the host parser's (test_host)
        d[int(k)] = v
        out.append(bytes(d))
    return out


def test_damaged_progressive_streams_match_oracle(jpeg, ctx, fixtures_dir):
    """Truncated / bit-damaged progressive files: same error name as the oracle (UnexpectedEof,
    BadHuffmanCode, UnexpectedHuffmanCode, TooManyCoefficients, ...) or, when the damage still decodes,
    the same pixels.  Exercises the refinement pass's skipped-by-count correction bits on bad input."""
    datas = []
    for name, seed in [("video-001.q50.420.progressive.jpeg", 1), ("video-005.gray.q50.progressive.jpeg", 2),
                       ("video-001.separate.dc.progression.progressive.jpeg", 3)]:
        datas += _damaged(_read(fixtures_dir, name), seed, 12, 40)
    datas += _damaged(S.encode(50021, 160, 120, subsampling="4:2:0", progressive=True, restart_rows=1), 4, 6, 20)
    _assert_same(jpeg, ctx, datas)


def test_end_of_band_run_carried_into_the_next_scan(jpeg, ctx, fixtures_dir):
    """corrupt progressive streams in which a scan ends inside an End-Of-Band run: the reference keeps the run
    (decoder.zig:144) and the next scan starts by skipping blocks; the GPU path flags such frames in the batch decode and
    decodes them again scan by scan with the run handed on (zpx_api.cu: rescue_eob_carry): same pixels / same error."""
    rng = np.random.default_rng(77)
    datas = []
    for name in PROGRESSIVE_FIXTURES[:7]:
        base = _read(fixtures_dir, name)
        first_sos = base.index(b"\xff\xda")
        found = 0
        for _ in range(300):  # (about 6 % of single-byte damages leave a run open at a scan boundary)
            d = bytearray(base)
            d[int(rng.integers(first_sos + 14, len(d) - 2))] = int(rng.integers(0, 256))
            d = bytes(d)
            try:
                carry = O.decode(d).eob_carry
            except O.OracleError:
                carry = O.last_eob_carry()
            if carry:
                datas.append(d)
                found += 1
                if found == 8:
                    break
    assert len(datas) >= 20
    # among good frames, in both progressive modes, natively too
    good = [_read(fixtures_dir, n) for n in PROGRESSIVE_FIXTURES[:3]]
    mixed = [x for pair in zip(datas, (good * len(datas))[:len(datas)]) for x in pair]
    _assert_same(jpeg, ctx, mixed)
    c2 = jpeg.Context()
    c2.set_option(10, 1)
    c2.set_option(9, 1)
    try:
        _assert_same(jpeg, c2, mixed)
        with jpeg.Batch(c2, mixed) as b:
            b.upload()
            b.decode()
            st = b.status()
            nat, st2 = b.fetch_native()
        for d, s_, n_ in zip(mixed, st, nat):
            try:
                ref = O.decode(d)
            except O.OracleError:
                continue
            if s_ == 0:
                assert np.array_equal(n_, ref.pixels)
    finally:
        c2.close()


def test_end_of_band_run_carried_across_sequential_scans(jpeg, ctx, fixtures_dir):
    """the same state in SEQUENTIAL multi-scan frames (End-Of-Band-run symbols do not belong there at all, SURVEY B6, but
    the reference honours them and keeps the run across scans): re-decoded scan by scan on the lane kernel"""
    from tools.multiscan import recode
    datas = []
    for name in ["video-001.q50.420.jpeg", "video-001.q50.444.jpeg", "video-001.q50.422.jpeg"]:
        base = _read(fixtures_dir, name)
        for script, ri in (([[0], [1], [2]], 0), ([[0], [1, 2]], 0), ([[2], [0], [1]], 3)):
            multi = recode(base, script, ri)
            for swaps in ({0x04: 0x30}, {0x03: 0x10, 0x12: 0x20}, {0x05: 0xE0}, {0x02: 0x10}, {0x11: 0x40, 0x21: 0x50}, {0x01: 0x90}):
                datas.append(_with_eob_run_symbols(multi, swaps))
    carried = 0
    for d in datas:
        try:
            carried += bool(O.decode(d).eob_carry)
        except O.OracleError:
            carried += bool(O.last_eob_carry())
    assert carried >= 5, carried
    _assert_same(jpeg, ctx, datas)


def test_progressive_warp_per_scan_kernel(jpeg, fixtures_dir):
    """Progressive frames take one lane per scan by default (zpx_k3l.cu); frames whose scan script is not an ordinary
    successive approximation fall back to one warp per scan (zpx_k3.cu).  ZPX_OPT_PROGRESSIVE_MODE = 1 sends every
    frame there: same fixtures, same damaged files, same oracle."""
    c = jpeg.Context()
    c.set_option(10, 1)
    try:
        datas = [_read(fixtures_dir, n) for n in PROGRESSIVE_FIXTURES]
        _assert_same(jpeg, c, datas, PROGRESSIVE_FIXTURES)
        datas = []
        for name, seed in [("video-001.q50.420.progressive.jpeg", 1), ("video-005.gray.q50.progressive.jpeg", 2)]:
            datas += _damaged(_read(fixtures_dir, name), seed, 12, 40)
        datas += [S.encode(50012, 333, 211, subsampling="4:2:2", progressive=True, restart_rows=1),
                  S.encode(50013, 97, 64, subsampling="4:2:0", progressive=True, restart_blocks=3)]
        _assert_same(jpeg, c, datas)
    finally:
        c.close()


def _scan_script_variants(data: bytes):
    """Rewrites of the SOS headers of a progressive file that keep it decodable by the reference but leave the ordinary
    successive-approximation order: a refinement pass turned into a second first pass of its band, a refinement that
    repeats the previous Al, a first pass whose Al is raised."""
    out = []
    pos, sos = 2, []
    while pos + 4 <= len(data):
        assert data[pos] == 0xFF
        m = data[pos + 1]
        ln = int.from_bytes(data[pos + 2:pos + 4], "big")
        if m == 0xDA:
            sos.append(pos)
            nxt = pos + 2 + ln
            while not (data[nxt] == 0xFF and data[nxt + 1] not in (0, 0xD0, 0xD1, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7)):
                nxt += 1
            pos = nxt
            continue
        if m == 0xD9:
            break
        pos += 2 + ln
    for p in sos:
        ln = int.from_bytes(data[p + 2:p + 4], "big")
        q = p + 2 + ln - 1  # the Ah/Al byte
        ah, al = data[q] >> 4, data[q] & 15
        ss = data[q - 2]
        for nah, nal in ((0, al), (al + 2, al + 1), (0, al + 1), (ah, al)):
            if (nah, nal) == (ah, al) or (nah != 0 and nah != nal + 1) or nal > 13:
                continue
            d = bytearray(data)
            d[q] = nah << 4 | nal
            out.append((bytes(d), f"sos@{p} ss={ss} ah/al {ah}/{al} -> {nah}/{nal}"))
    return out


def test_unusual_scan_scripts(jpeg, ctx, fixtures_dir):
    """Progressive files whose scan scripts are legal for the reference's parser but not an ordinary successive
    approximation (repeated first passes, refinements that do not lower Al): the lane-per-scan kernels must hand
    them to the kernel that works on the coefficients themselves; results as the oracle's either way."""
    datas, names = [], []
    for name in ["video-001.q50.420.progressive.jpeg", "video-005.gray.q50.progressive.jpeg"]:
        for d, tag in _scan_script_variants(_read(fixtures_dir, name)):
            datas.append(d)
            names.append(f"{name} {tag}")
    assert len(datas) > 20
    _assert_same(jpeg, ctx, datas, names)


def test_header_fuzz_decodes_like_oracle(jpeg, ctx, fixtures_dir):
    """Seeded damage in the marker segments: odd-but-legal sampling factors, table assignments, scan scripts,
    restart intervals ... must decode to the oracle's pixels, everything else must fail with its error."""
    rng = np.random.default_rng(77)
    datas = []
    for fname in ["video-001.jpeg", "video-001.q50.420.progressive.jpeg", "video-001.cmyk.jpeg", "video-001.restart2.jpeg",
                  "video-005.gray.q50.2x2.jpeg", "video-001.separate.dc.progression.jpeg", "video-001.rgb.jpeg",
                  "video-001.q50.411.jpeg"]:
        datas += header_damage(_read(fixtures_dir, fname), rng, 60)
    _assert_same(jpeg, ctx, datas)


def _with_eob_run_symbols(data: bytes, swaps):
    """Rewrite symbols of the AC Huffman tables (class 1 DHT segments) so that ordinary run/size symbols become
    End-Of-Band-run symbols (r, 0), 0 < r < 15 -- legal only in progressive scans, honoured by the reference in
    sequential ones too."""
    d = bytearray(data)
    i = 2
    while i + 4 <= len(d) and not (d[i] == 0xFF and d[i + 1] == 0xDA):
        seglen = (d[i + 2] << 8) | d[i + 3]
        if d[i + 1] == 0xC4:
            j = i + 4
            while j < i + 2 + seglen:
                tc, n = d[j] >> 4, sum(d[j + 1:j + 17])
                if tc == 1:
                    for k in range(j + 17, j + 17 + n):
                        if d[k] in swaps:
                            d[k] = swaps[d[k]]
                j += 17 + n
        i += 2 + seglen
    return bytes(d)


def test_eob_runs_in_sequential_scans(jpeg, ctx, fixtures_dir):
    """the reference counts End-Of-Band runs down in baseline scans as well (decoder.zig:1379-1407),
    resetting them only at RSTn.  Such scans always take the lane-per-interval kernel, which models the run."""
    datas = []
    for name in ["video-001.restart2.jpeg", "video-001.jpeg", "video-005.gray.jpeg", "video-001.q50.420.jpeg",
                 "video-001.cmyk.jpeg"]:
        base = _read(fixtures_dir, name)
        for swaps in ({0x04: 0x30}, {0x03: 0x10, 0x12: 0x20}, {0x05: 0xE0}, {0x02: 0x10}, {0x11: 0x40, 0x21: 0x50}):
            datas.append(_with_eob_run_symbols(base, swaps))
    oks = sum(1 for d in datas if _oracle_rgba(d)[0] is not None)
    assert oks >= 5, oks  # enough of them decode (garbage, but the same garbage) rather than fail
    _assert_same(jpeg, ctx, datas)


def _multiscan_cases():
    from tools.multiscan import recode
    out = []
    for seed, (w, h), kw in [(60000, (97, 75), dict(subsampling="4:2:0")), (60002, (130, 50), dict(subsampling="4:2:2")),
                             (60004, (64, 64), dict(subsampling="4:4:4")), (60006, (83, 41), dict(mode="CMYK")),
                             (60008, (640, 360), dict(subsampling="4:2:0"))]:
        base = S.encode(seed, w, h, **kw)  # even seed: Annex K tables, complete
        nc = 4 if kw.get("mode") == "CMYK" else 3
        scripts = [([[c] for c in range(nc)], 0), ([[0], list(range(1, nc))], 0), ([[nc - 1], [0], *[[c] for c in range(1, nc - 1)]], 3),
                   ([[0, 1], *[[c] for c in range(2, nc)]], 5), ([[c] for c in range(nc)], 1)]
        out += [recode(base, sc, ri) for sc, ri in scripts]
    return out


def test_multi_scan_sequential_frames(jpeg, ctx):
    """Sequential frames whose components arrive in separate scans (non-interleaved block order, restart
    intervals counted in the scan's own MCU iterations; decoder.zig:1294-1336).  Pillow cannot write
    them: tools/multiscan.py re-codes single-scan files.  These take the planar coefficient layout (one block grid per
    component), which the fused kernel gathers tile by tile."""
    _assert_same(jpeg, ctx, _multiscan_cases())


def test_planar_layout_frames_fused_equal_generic_equal_oracle(jpeg, fixtures_dir):
    """Frames that keep one block grid per component (multi-scan sequential, progressive) through the fused kernel
    (default, RGBA output) and through the unfused kernels (ZPX_OPT_FORCE_GENERIC): both == oracle.  Every sampling
    the fused kernel has, sizes with several tiles per MCU row and partial MCUs at the edges."""
    import ctypes as C
    from tools.multiscan import recode
    from zpix_b200 import _lib as zl
    datas = []
    for n in ["video-001.q50.410.jpeg", "video-001.q50.411.jpeg", "video-001.q50.440.jpeg", "video-001.q50.444.jpeg",
              "video-001.q50.422.jpeg", "video-001.q50.420.jpeg"]:
        datas.append(recode(_read(fixtures_dir, n), [[0], [1], [2]], 0))
        datas.append(recode(_read(fixtures_dir, n), [[2], [0, 1]], 0))
    datas.append(recode(_read(fixtures_dir, "video-005.gray.q50.jpeg"), [[0]], 2))
    for seed, (w, h), kw in [(61000, (1920, 136), dict(subsampling="4:2:0")), (61001, (1363, 77), dict(subsampling="4:2:2")),
                             (61002, (1000, 40), dict(subsampling="4:4:4")), (61003, (1100, 33), dict(mode="L"))]:
        datas.append(S.encode(seed, w, h, progressive=True, **kw))
        datas.append(recode(S.encode(seed + 10, w, h, **kw), [[0]] if kw.get("mode") == "L" else [[0], [1], [2]], 0))
    c1, c2 = jpeg.Context(), jpeg.Context()
    c2.set_option(2, 1)
    try:
        for d in datas:  # the host parser's report: these frames are the fused kernel's
            a8 = np.frombuffer(d, np.uint8)
            inf, rep = zl.ZpxImageInfo(), zl.ZpxParseReport()
            assert zl.lib.zpx_parse_report_of(a8.ctypes.data, a8.size, C.byref(inf), C.byref(rep)) == 0
            assert rep.status == 0 and rep.fused == 1
        _assert_same(jpeg, c1, datas)
        a, sa, ta = _gpu_batch(jpeg, c1, datas)
        b, sb, tb = _gpu_batch(jpeg, c2, datas)
        assert sa == sb == [0] * len(datas)
        assert ta["idct_fused_bytes"] > 0 and tb["idct_fused_bytes"] == 0
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    finally:
        c1.close()
        c2.close()


def test_files_found_by_fuzzing(jpeg, ctx, golden_dir):
    """Regression files from tools/fuzz_hunt.py: garbage streams the reference still decodes.
    idct_dc_row_wrap_*: first-column values beyond 2^20, where the reference's all-AC-zero row shortcut (s0 << 3,
    idct.zig:84-97) and the general row formula part ways; prog_run_past_band_end: a run/size symbol whose run
    leaves the band (value bits stay unread, decoding goes on); eob_run_in_baseline_dri: SURVEY B6."""
    d = os.path.join(golden_dir, "fuzz_found")
    names = sorted(os.listdir(d))
    assert len(names) >= 5
    _assert_same(jpeg, ctx, [open(os.path.join(d, n), "rb").read() for n in names], names)


def test_fuzz_every_fixture(jpeg, fixtures_dir):
    """A wider net than the tests above, default entropy mode only: every fixture of the reference, damaged in its
    marker segments, in its entropy-coded data and by truncation (~1000 files in one batch)."""
    rng = np.random.default_rng(4242)
    datas = []
    for k, name in enumerate(sorted(os.listdir(fixtures_dir))):
        if name == "iceberg.jpg":
            continue
        base = _read(fixtures_dir, name)
        datas += header_damage(base, rng, 24)
        datas += _damaged(base, 1000 + k, 4, 14)
    c = jpeg.Context()
    _assert_same(jpeg, c, datas)
    c.close()


def test_damaged_baseline_streams_match_oracle(jpeg, ctx, fixtures_dir):
    datas = []
    for name, seed in [("video-001.q50.420.jpeg", 5), ("video-001.restart2.jpeg", 6), ("video-005.gray.jpeg", 7),
                       ("video-001.cmyk.jpeg", 8)]:
        datas += _damaged(_read(fixtures_dir, name), seed, 10, 30)
    _assert_same(jpeg, ctx, datas)


def test_native_variant_matches_reference_planes(jpeg, fixtures_dir):
    """jpeg.load mirror (N2): .YCbCr / .Gray planes with makeImg's exact strides, whole planes including the MCU
    padding (stricter than the reference's own `check`, decoder.zig:1803-1836, which looks at blocks whose origin is
    inside bounds)."""
    from tools.multiscan import recode
    cases = [_read(fixtures_dir, name) for name in
             ["video-001.q50.420.jpeg", "video-001.q50.422.jpeg", "video-001.jpeg", "video-005.gray.jpeg",
              "video-001.221212.jpeg", "video-001.q50.411.jpeg", "video-001.q50.420.progressive.jpeg",
              "video-001.progressive.jpeg", "video-005.gray.q50.progressive.jpeg"]]
    # component scans of their own: blocks outside the image are never coded, the planes keep makeImg's zeros there
    cases.append(recode(S.encode(60010, 97, 75, subsampling="4:2:0"), [[0], [1], [2]], 0))
    for data in cases:
        ref = O.decode(data)
        img = jpeg.loadFromBuffer(data)
        assert img.tag == ref.variant_name
        r = img.bounds()
        assert (r.dX(), r.dY()) == (ref.width, ref.height)
        if img.tag == "Gray":
            got = img.Gray.pixels.reshape(-1, img.Gray.stride)
            assert img.Gray.stride == ref.stride
            assert got.shape == ref.pix.shape and np.array_equal(got, ref.pix)  # MCU padding included
        else:
            m = img.YCbCr
            assert (m.y_stride, m.c_stride) == (ref.y_stride, ref.c_stride)
            assert m.subsample_ratio.name == ref.subsample_ratio
            assert np.array_equal(m.y.reshape(-1, m.y_stride), ref.y)
            assert np.array_equal(m.cb.reshape(-1, m.c_stride), ref.cb)
            assert np.array_equal(m.cr.reshape(-1, m.c_stride), ref.cr)
        assert np.array_equal(img.rgbaPixels().reshape(ref.height, ref.width, 4), ref.rgbaPixels())


def test_native_planes_on_demand_after_a_fused_decode(jpeg, ctx, fixtures_dir):
    """zpx_batch_fetch_native after a plain decode (ZPX_OPT_NATIVE_PLANES = 0): the RGBA came from the fused kernel, the
    planes of those images are reconstructed when asked for, from the resident coefficients: Image.pixels byte for byte
    (whole planes, MCU padding and the reference's never-reconstructed blocks included); a second fetch reuses them."""
    from tools.multiscan import recode
    names = ["video-001.q50.420.jpeg", "video-001.q50.422.jpeg", "video-001.q50.444.jpeg", "video-001.q50.411.jpeg",
             "video-005.gray.jpeg", "video-001.221212.jpeg", "video-001.cmyk.jpeg", "video-001.rgb.jpeg",
             "video-001.q50.420.progressive.jpeg", "video-005.gray.q50.progressive.jpeg", "video-001.restart2.jpeg"]
    datas = [_read(fixtures_dir, n) for n in names]
    datas.append(recode(S.encode(60010, 97, 75, subsampling="4:2:0"), [[0], [1], [2]], 0))
    datas += S.make_batch(2, 2, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    refs = [O.decode(d) for d in datas]
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        b.decode()
        launches = ctx.kernel_launches
        for attempt in range(2):
            nat, st = b.fetch_native()
            assert st == [0] * len(datas)
            for i, (got, ref) in enumerate(zip(nat, refs)):
                assert got.size == ref.pixels.size and np.array_equal(got, ref.pixels), (i, attempt)
            if attempt == 0:
                assert ctx.kernel_launches > launches  # the planes were made now ...
                launches = ctx.kernel_launches
            else:
                assert ctx.kernel_launches == launches + 1  # ... and only the CMYK interleave runs again
        outs, st = b.fetch_rgba()
        for o, ref in zip(outs, refs):
            assert np.array_equal(o.reshape(-1), ref.rgbaPixels().reshape(-1))


def test_native_cmyk_variant(jpeg, fixtures_dir):
    """4-component frames come back as Image{.CMYK} holding applyBlack's interleave (decoder.zig:852-901;
    the YCbCrK branch :811-846 by intent), and rgbaPixels() of it is the GPU's RGBA."""
    cases = [_read(fixtures_dir, "video-001.cmyk.jpeg"), S.encode(51234, 120, 88, "CMYK"),
             S.encode(51235, 100, 72, "CMYK", ycck=True)]
    for data in cases:
        ref = O.decode(data)
        img = jpeg.loadFromBuffer(data)
        assert img.tag == "CMYK" == ref.variant_name
        m = img.CMYK
        assert m.stride == ref.stride == 4 * ref.width
        assert np.array_equal(m.pixels.reshape(ref.height, ref.width, 4), ref.pix.reshape(ref.height, ref.width, 4))
        assert np.array_equal(img.rgbaPixels().reshape(ref.height, ref.width, 4), ref.rgbaPixels())
        c = m.at(3, 5)
        assert c.toRGBA()[0] >> 8 == int(ref.rgbaPixels()[5, 3, 0])


def test_batch_shapes_at_the_edges(jpeg, fixtures_dir):
    """Empty batch, a batch with nothing decodable, thousands of tiny images (planner / tile-table limits), and a
    context reused across batches of very different sizes."""
    c = jpeg.Context([0])
    outs, st = jpeg.decodeBatchOneCall([], c)
    assert outs == [] and st == []
    with jpeg.Batch(c, []) as b:
        b.upload()
        b.decode()
        assert b.fetch_rgba() == ([], [])
    junk = [b"", b"\xff\xd8", b"not a jpeg at all", b"\xff\xd8\xff\xd9"]
    outs, st = jpeg.decodeBatchOneCall(junk, c)
    assert all(o is None for o in outs) and all(s != 0 for s in st)
    for d, s in zip(junk, st):
        assert jpeg.lib.zpx_error_name(s).decode() == _oracle_rgba(d)[1]
    tiny = [S.encode(53000 + i, 1 + i % 9, 1 + (i * 7) % 11, subsampling=["4:2:0", "4:4:4", "4:2:2"][i % 3]) for i in range(12)]
    tiny += [S.encode(53100, 8, 8, mode="L"), S.encode(53101, 3, 17, mode="CMYK")]
    want = [_oracle_rgba(d)[0] for d in tiny]
    batch = [tiny[i % len(tiny)] for i in range(6000)]
    for mode in (0, 1, 2):
        c.set_option(1, mode)
        with jpeg.Batch(c, batch) as b:
            b.upload()
            b.decode()
            outs, st = b.fetch_rgba()
        assert not any(st)
        for i in range(0, 6000, 37):
            assert np.array_equal(outs[i], want[i % len(tiny)]), (mode, i)
    c.set_option(1, 0)
    outs, st = jpeg.decodeBatchOneCall(batch[:3000], c)   # through the chunk pipeline as well
    assert not any(st) and all(np.array_equal(outs[i], want[i % len(tiny)]) for i in range(0, 3000, 41))
    big = _read(fixtures_dir, "iceberg.jpg")
    _assert_same(jpeg, c, [big, tiny[0], big])
    c.close()


def test_gpu_resident_hand_off(jpeg, fixtures_dir):
    """a consumer on the GPU takes the RGBA where the kernels left it (zpx_batch_device_rgba), on the
    caller's own stream, with no device->host copy in between."""
    import torch

    class _Dev:  # __cuda_array_interface__ view of library-owned device memory
        def __init__(self, ptr, h, w):
            self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "|u1", "data": (ptr, False), "version": 3}

    names = ["video-001.q50.420.jpeg", "video-005.gray.jpeg", "video-001.cmyk.jpeg", "video-001.progressive.jpeg"]
    datas = [_read(fixtures_dir, n) for n in names]
    c = jpeg.Context([0])
    stream = torch.cuda.Stream()
    with jpeg.Batch(c, datas) as b:
        b.upload()
        b.decode(stream.cuda_stream)  # asynchronous: the caller's stream orders decode and consumer
        with torch.cuda.stream(stream):
            views = []
            for i in range(len(datas)):
                inf = b.info(i)
                ptr = b.device_rgba_ptr(i)
                assert ptr != 0
                views.append(torch.as_tensor(_Dev(ptr, inf.height, inf.width), device="cuda"))
            sums = [int(v.sum(dtype=torch.int64)) for v in views]   # a reduction on the GPU, then 8 bytes per image
            copies = [v.clone() for v in views]
        stream.synchronize()
        assert all(s == 0 for s in b.status())
    for d, got, total in zip(datas, copies, sums):
        want = O.decode(d).rgbaPixels()
        assert np.array_equal(got.cpu().numpy(), want)
        assert total == int(want.astype(np.int64).sum())
    c.close()


def test_dispatcher_routes_jpegs_to_the_batch_path(jpeg, fixtures_dir):
    """zpix.fromBuffer / fromBuffers (src/root.zig:24-40): probe, JPEG -> GPU path, anything else UnknownImageFormat."""
    import zpix_b200 as zpix
    a = _read(fixtures_dir, "video-001.q50.420.jpeg")
    b = _read(fixtures_dir, "video-005.gray.jpeg")
    png = b"\x89PNG\r\n\x1a\n" + bytes(32)
    res = zpix.fromBuffers([a, png, b, b""])
    assert isinstance(res[1], ValueError) and isinstance(res[3], ValueError)
    assert np.array_equal(res[0].rgbaPixels().reshape(-1), O.decode(a).rgbaPixels().reshape(-1))
    assert np.array_equal(res[2].rgbaPixels().reshape(-1), O.decode(b).rgbaPixels().reshape(-1))
    one = zpix.fromBuffer(a)
    assert one.tag == "YCbCr" and np.array_equal(one.rgbaPixels().reshape(-1), res[0].rgbaPixels().reshape(-1))
    with pytest.raises(ValueError):
        zpix.fromBuffer(png)


def test_all_devices_of_the_box_in_one_context(jpeg, fixtures_dir):
    """the host scheduler splits a batch over the GPUs of one box (contiguous ranges balanced by
    entropy-coded bytes, one host thread per device), no collective.  Runs on however many GPUs are visible."""
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("one GPU visible")
    names = BASELINE_FIXTURES + PROGRESSIVE_FIXTURES
    datas = [_read(fixtures_dir, n) for n in names] * 3
    datas += S.make_batch(2, 6, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    datas += [S.encode(52001 + i, 640, 480, subsampling="4:4:4") for i in range(5)]
    ctx = jpeg.Context(list(range(nd)))
    assert ctx.num_devices == nd
    for mode in (0, 2):
        ctx.set_option(1, mode)
        _assert_same(jpeg, ctx, datas)
    with jpeg.Batch(ctx, datas) as b:
        b.upload()
        b.decode()
        used = [di for di in range(nd) if b.timing(di)["images"] > 0]
    assert len(used) >= 2
    ctx.close()


def test_one_call_chunk_pipeline(jpeg, fixtures_dir):
    """zpx_decode_batch_rgba cuts large batches into chunks that alternate between two sets of device
    buffers; results and per-image status must not depend on the chunking."""
    names = BASELINE_FIXTURES + PROGRESSIVE_FIXTURES
    datas = [_read(fixtures_dir, n) for n in names]
    bad = datas[3][: len(datas[3]) // 2]
    batch = (datas + [bad]) * 12  # 312 inputs -> chunks of 128/64/32
    want = [_oracle_rgba(d) for d in datas + [bad]]
    for chunk in (0, 64, 32, -1):
        c = jpeg.Context()
        c.set_option(4, chunk)
        outs, st = jpeg.decodeBatchOneCall(batch, c)
        for i, (o, s) in enumerate(zip(outs, st)):
            w, err = want[i % len(want)]
            if w is None:
                assert s != 0 and jpeg.lib.zpx_error_name(s).decode() == err
            else:
                assert s == 0 and np.array_equal(o, w), (chunk, i)
        c.close()


def test_one_resident_batch_per_context(jpeg, fixtures_dir):
    """A context has one set of device buffers: a second upload evicts the first batch (BAD_STATE, never
    silently wrong pixels)."""
    c = jpeg.Context()
    a = jpeg.Batch(c, [_read(fixtures_dir, "video-001.jpeg")])
    b = jpeg.Batch(c, [_read(fixtures_dir, "video-005.gray.jpeg")])
    a.upload()
    a.decode()
    b.upload()
    with pytest.raises(jpeg.JpegError) as e:
        a.fetch_rgba()
    assert e.value.name == "BadState"
    b.decode()
    outs, st = b.fetch_rgba()
    assert st == [0] and np.array_equal(outs[0], O.decode(_read(fixtures_dir, "video-005.gray.jpeg")).rgbaPixels())
    a.close()
    b.close()
    c.close()


def test_empty_batch_and_api_contract(jpeg, ctx):
    outs, st, _ = _gpu_batch(jpeg, ctx, [])
    assert outs == [] and st == []
    res = jpeg.decodeBatch([b"not a jpeg", b""], ctx)
    assert all(isinstance(r, jpeg.JpegError) for r in res)
    assert res[0].name == "InvalidSOIMarker" and res[1].name == "UnexpectedEof"


def test_noise_images_at_quality_extremes(jpeg, ctx):
    """K2 through real files (block-level property tests live in test_gpu_blocks.py): noise at the extremes of the 8-bit
    range (quality 100 and quality 1 images of pure noise) still matches the oracle bit for bit."""
    rng = np.random.default_rng(5)
    from PIL import Image
    import io

    datas = []
    for q in (1, 30, 100):
        for sub in ("4:4:4", "4:2:0"):
            px = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
            px[::2] = 255 - px[::2] // 8  # harsh edges -> large coefficients, saturating IDCT outputs
            buf = io.BytesIO()
            Image.fromarray(px, "RGB").save(buf, "JPEG", quality=q, subsampling=sub)
            datas.append(buf.getvalue())
    _assert_same(jpeg, ctx, datas)


def test_cpp_host_layer_end_to_end(fixtures_dir, tmp_path):
    """cpp/zpix.hpp (C++ mirror of jpeg.load / loadBatch / rgbaPixels) over the same C ABI: hashes of the RGBA bytes
    equal the oracle's."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "zpix_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "cpp", "zpix_demo.cpp"), "-o", str(exe),
                           "-L" + os.path.join(root, "zpix_b200"), "-lzpixcuda", "-Wl,-rpath," + os.path.join(root, "zpix_b200")])
    names = ["video-001.jpeg", "video-005.gray.jpeg", "video-001.cmyk.jpeg", "video-001.q50.420.progressive.jpeg"]
    paths = [os.path.join(fixtures_dir, n) for n in names]
    r = subprocess.run([str(exe)] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(names)
    for p, line in zip(paths, lines):
        want = O.decode(open(p, "rb").read())
        h = 1469598103934665603
        for b in want.rgbaPixels().tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert line.split()[1] == f"{want.width}x{want.height}"
        assert line.split()[2] == f"fnv1a={h:016x}", line
    # the batch dispatcher (zpix::fromBuffers, src/root.zig:24-40): native variants, a PNG signature -> UnknownImageFormat
    r = subprocess.run([str(exe), "--native"] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for p, line in zip(paths, r.stdout.strip().splitlines()):
        want = O.decode(open(p, "rb").read())
        h = 1469598103934665603
        for b in want.pixels.tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert line.split()[1] == f"variant={want.variant}" and line.split()[3] == f"fnv1a={h:016x}", line


def test_image_over_the_memory_budget_fails_alone(jpeg, fixtures_dir, monkeypatch):
    """An image whose device buffers do not fit gets error.OutOfMemory (the reference's answer to a failed makeImg
    allocation, decoder.zig:1708-1783) as ITS status; the rest of the batch decodes.  ZPX_IMAGE_BUDGET_MB shrinks the
    budget so that an ordinary file plays the part of the 65535 x 65535 one."""
    small = _read(fixtures_dir, "video-001.q50.420.jpeg")
    big = S.encode(91000, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    monkeypatch.setenv("ZPX_IMAGE_BUDGET_MB", "4")   # 1080p needs ~17 MB, the fixtures ~0.1 MB
    c = jpeg.Context([0])
    outs, st = jpeg.decodeBatchOneCall([small, big, small], c)
    assert st[0] == 0 and st[2] == 0 and jpeg.lib.zpx_error_name(st[1]).decode() == "OutOfMemory"
    want = O.decode(small).rgbaPixels()
    assert np.array_equal(outs[0], want) and np.array_equal(outs[2], want) and outs[1] is None
    monkeypatch.delenv("ZPX_IMAGE_BUDGET_MB")
    outs, st = jpeg.decodeBatchOneCall([small, big, small], c)
    assert st == [0, 0, 0] and np.array_equal(outs[1], O.decode(big).rgbaPixels())
    c.close()


def test_results_read_after_a_decode_on_the_callers_stream(jpeg, fixtures_dir):
    """zpx_batch_decode(b, stream) returns once the work is enqueued; status and pixels fetched right after, without
    the caller synchronising, must still be the decode's (the library orders its own stream after the caller's)."""
    import torch
    datas = S.make_batch(2, 24, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    bad = datas[5][: len(datas[5]) // 2]
    datas[5] = bad
    c = jpeg.Context([0])
    stream = torch.cuda.Stream()
    with jpeg.Batch(c, datas) as b:
        b.upload()
        for _ in range(3):
            b.decode(stream.cuda_stream)
            outs, st = b.fetch_rgba()      # no stream.synchronize() in between
            assert st[5] != 0 and all(s == 0 for i, s in enumerate(st) if i != 5)
            for i in (0, 7, 23):
                assert np.array_equal(outs[i], O.decode(datas[i]).rgbaPixels())
    c.close()


def test_self_synchronising_decoder_on_the_callers_stream(jpeg, fixtures_dir):
    """streams without restart markers decoded on a caller's stream: the synchronisation rounds are enqueued blind (gated
    on device flags, nothing read back).  Default sub-sequences converge within them; 32-byte sub-sequences need more
    rounds than were enqueued would be noticed by the status call and repaired by decoding again with the host loop
    (forced here through the option's test value).  Same
    pixels as the oracle either way, also when the device pointer is taken after zpx_batch_status alone."""
    import torch
    datas = S.make_batch(3, 6, 512, 512, first=5000, mode="YCbCr", subsampling="4:4:4") + S.make_batch(3, 4, 512, 512, mode="L")
    datas += [_read(fixtures_dir, "iceberg.jpg"), S.encode(63000, 1920, 1080, subsampling="4:2:0")]
    want = [O.decode(d).rgbaPixels() for d in datas]
    stream = torch.cuda.Stream()
    for mode, sub, rounds in ((2, 0, 2), (2, 32, 2), (2, 32, 0), (2, 0, -1), (0, 0, 2)):
        c = jpeg.Context([0])
        c.set_option(1, mode)
        c.set_option(3, sub)
        c.set_option(12, rounds)  # ZPX_OPT_GATED_SWEEPS; -1 = test hook: the rounds count as not converged
        with jpeg.Batch(c, datas) as b:
            b.upload()
            for _ in range(2):
                launches = c.kernel_launches
                b.decode(stream.cuda_stream)
                enq = c.kernel_launches - launches
                assert b.status() == [0] * len(datas)
                if rounds == -1:
                    assert c.kernel_launches - launches > enq  # the status call decoded again
                ptr = b.device_rgba_ptr(len(datas) - 2)
                assert ptr
                outs, st = b.fetch_rgba()
                for o, w in zip(outs, want):
                    assert np.array_equal(o, w), (mode, sub)
        c.close()


def test_native_batch_one_call(jpeg, fixtures_dir):
    """zpx_decode_batch_native: the whole batch comes back as the Image variants jpeg.load returns -- planes of the
    fused kernel (MCU padding included) for the ordinary files, unfused kernels for the rest -- through the chunk
    pipeline, with a failing file in the middle."""
    names = ["video-001.q50.420.jpeg", "video-001.q50.422.jpeg", "video-001.jpeg", "video-005.gray.jpeg", "video-001.221212.jpeg",
             "video-001.q50.411.jpeg", "video-001.q50.410.jpeg", "video-001.q50.440.jpeg", "video-001.q50.420.progressive.jpeg",
             "video-001.rgb.jpeg", "video-001.cmyk.jpeg", "video-001.restart2.jpeg"]
    datas = [_read(fixtures_dir, n) for n in names]
    datas += S.make_batch(2, 3, 1920, 1080, subsampling="4:2:0", restart_rows=1)
    datas += [S.encode(54000, 513, 511, mode="L"), S.encode(54001, 641, 479, subsampling="4:4:4")]
    bad = datas[0][: len(datas[0]) // 2]
    batch = (datas + [bad]) * 6
    for chunk in (0, 16, -1):
        c = jpeg.Context([0])
        c.set_option(4, chunk)
        res = jpeg.decodeBatchNative(batch, c)
        for k, (d, img) in enumerate(zip(batch, res)):
            if d is bad:
                assert isinstance(img, jpeg.JpegError) and img.name == "UnexpectedEof"
                continue
            ref = O.decode(d)
            assert img.tag == ref.variant_name, k
            if img.tag == "Gray":
                assert np.array_equal(img.Gray.pixels.reshape(-1, img.Gray.stride), ref.pix)
            elif img.tag == "YCbCr":
                m = img.YCbCr
                assert (m.y_stride, m.c_stride) == (ref.y_stride, ref.c_stride)
                assert np.array_equal(m.y.reshape(-1, m.y_stride), ref.y)
                assert np.array_equal(m.cb.reshape(-1, m.c_stride), ref.cb)
                assert np.array_equal(m.cr.reshape(-1, m.c_stride), ref.cr)
            else:
                assert np.array_equal(img.payload.pixels.reshape(-1), ref.pix.reshape(-1))
        c.close()


def test_full_size_batches_every_image(jpeg):
    """The BASELINE.json batches at FULL size, every output image checked (no sampling): cfg2 = 1024 distinct 1080p
    4:2:0 files with DRI, each compared with the oracle on all host threads; cfg3 = 4096 x 512x512 (gray + 4:4:4, no
    DRI) and cfg4 = 512 x 2160p 4:2:2 with and without DRI, built from 64 / 16 distinct files: the first copy of each
    against the oracle, every other copy against that one on the GPU.  Size-independent check on top: the number of
    pixels and the per-image status of every output."""
    import ctypes as C
    import bench

    c = jpeg.Context([0])
    # cfg2 through the one-call API into one pinned buffer
    datas = S.make_batch(2, 1024, 1920, 1080, cache_dir=bench.CACHE, subsampling="4:2:0", restart_rows=1)
    n, out_bytes = len(datas), 4 * 1920 * 1080
    pinned = jpeg.lib.zpx_host_alloc(out_bytes * n)
    assert pinned
    try:
        outs = (C.c_void_p * n)(*[pinned + i * out_bytes for i in range(n)])
        keep = [np.frombuffer(d, np.uint8) for d in datas]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in keep])
        lens = (C.c_size_t * n)(*[a.size for a in keep])
        st = (C.c_int32 * n)()
        assert jpeg.lib.zpx_decode_batch_rgba(c.handle, ptrs, lens, n, outs, None, st) == 0
        assert not any(st)
        par = bench.parity_host(datas, pinned, out_bytes, os.cpu_count() or 1)
        assert par == {"images": 1024, "mismatch": 0, "against": par["against"]}
    finally:
        jpeg.lib.zpx_host_free(pinned)

    def rep(base, k):
        return [base[i % len(base)] for i in range(k)]

    g = S.make_batch(3, 32, 512, 512, cache_dir=bench.CACHE, mode="L")
    y = S.make_batch(3, 32, 512, 512, cache_dir=bench.CACHE, first=5000, mode="YCbCr", subsampling="4:4:4")
    cfg3 = rep([x for pair in zip(g, y) for x in pair], 4096)
    cfg4a = rep(S.make_batch(4, 16, 3840, 2160, cache_dir=bench.CACHE, mode="YCbCr", subsampling="4:2:2", restart_rows=1), 512)
    cfg4b = rep(S.make_batch(4, 16, 3840, 2160, cache_dir=bench.CACHE, mode="YCbCr", subsampling="4:2:2"), 512)
    for batch_data, distinct in ((cfg3, 64), (cfg4a, 16), (cfg4b, 16)):
        with jpeg.Batch(c, batch_data) as b:
            b.upload()
            b.decode()
            assert not any(b.status())
            par = bench.parity_device(b, batch_data, distinct)
        assert par["images"] == len(batch_data) and par["distinct"] == distinct and par["mismatch"] == 0, par
    c.close()
