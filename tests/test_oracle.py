"""Pins the CPU oracle (oracle/zpix_oracle.c) against every test the reference
holds for the JPEG path (reference src/jpeg/decoder.zig:1843-2279) and against
the independent goldens of SURVEY.md Appendix C.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

# sha256(RGBA)[:16] from an independent restatement.
APPENDIX_C = {
    "iceberg.jpg": "d2943764ea20b49f",
    "video-001.jpeg": "da0a9efc582a7aa8",
    "video-001.progressive.jpeg": "da0a9efc582a7aa8",
    "video-001.q50.410.jpeg": "427ce31bc36775b5",
    "video-001.q50.410.progressive.jpeg": "427ce31bc36775b5",
    "video-001.q50.411.jpeg": "a19a6c550a0f2cd0",
    "video-001.q50.411.progressive.jpeg": "a19a6c550a0f2cd0",
    "video-001.q50.420.jpeg": "57ccaab2cc4d4931",
    "video-001.q50.420.progressive.jpeg": "57ccaab2cc4d4931",
    "video-001.q50.422.jpeg": "1dd3cddf25fbebed",
    "video-001.q50.422.progressive.jpeg": "1dd3cddf25fbebed",
    "video-001.q50.440.jpeg": "095f63dfa59bf1c5",
    "video-001.q50.440.progressive.jpeg": "095f63dfa59bf1c5",
    "video-001.q50.444.jpeg": "ae4bc7213e461165",
    "video-001.q50.444.progressive.jpeg": "ae4bc7213e461165",
    "video-001.separate.dc.progression.jpeg": "09646df57398bbb1",
    "video-001.separate.dc.progression.progressive.jpeg": "09646df57398bbb1",
    "video-005.gray.q50.jpeg": "a7684585c0e7fa0a",
    "video-005.gray.q50.progressive.jpeg": "a7684585c0e7fa0a",
    "video-005.gray.q50.2x2.jpeg": "a7684585c0e7fa0a",
    "video-005.gray.q50.2x2.progressive.jpeg": "a7684585c0e7fa0a",
    "video-005.gray.jpeg": "e1a0d584f58c3fc8",
    "video-001.221212.jpeg": "6973197b4a85ee47",
    "video-001.restart2.jpeg": "4ee401171a2a0a6c",
    "video-001.rgb.jpeg": "75c1b55eed6db5ef",
    "video-001.cmyk.jpeg": "b6fc252fb21fd9fc",
}
APPENDIX_C_ICEBERG_FULL = "d2943764ea20b49f84921e19d95a6438fac21a2b7160c6b6455a02da4cc694eb"
APPENDIX_C_ICEBERG_PLANES = {
    "y": "2a41893538e7d346d18cc64d879a4c28926593744d14dacb49570be3cf9eb736",
    "cb": "106dccf7aab9bc4f0ecb7c78e4ee5cae281491bf39f9fe7e1a1485a9472c1ead",
    "cr": "027a346d588ddc8215cea3a5712ab684af1dda6965b06d1556b1197bff44fc3b",
}
EXPECTED_VARIANT = {
    "video-001.cmyk.jpeg": "CMYK",
    "video-001.rgb.jpeg": "RGBA",
    "video-005.gray.jpeg": "Gray",
    "video-005.gray.q50.jpeg": "Gray",
    "video-005.gray.q50.2x2.jpeg": "Gray",
}


def _read(d, name):
    with open(os.path.join(d, name), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", sorted(APPENDIX_C))
def test_appendix_c_hashes(fixtures_dir, golden_dir, name):
    img = O.decode(_read(fixtures_dir, name))
    h = hashlib.sha256(img.rgbaPixels().tobytes()).hexdigest()
    assert h[:16] == APPENDIX_C[name]
    full = json.load(open(os.path.join(golden_dir, "rgba_sha256.json")))
    assert full[name] == h
    want_variant = "Gray" if "gray" in name else EXPECTED_VARIANT.get(name, "YCbCr")
    assert img.variant_name == want_variant
    if name == "iceberg.jpg":
        assert h == APPENDIX_C_ICEBERG_FULL
        assert (img.width, img.height) == (2048, 2048)
        for k, want in APPENDIX_C_ICEBERG_PLANES.items():
            plane = getattr(img, k)[:2048, :2048]
            assert hashlib.sha256(np.ascontiguousarray(plane).tobytes()).hexdigest() == want
    else:
        assert (img.width, img.height) == (150, 103)


def _check_planes(bounds_w, bounds_h, p0, p1):
    """decoder.zig:1803-1836: planes must agree on every 8x8 block whose origin is inside bounds."""
    s0, s1 = p0.shape[1], p1.shape[1]
    assert s0 > 0 and s0 % 8 == 0 and s1 > 0 and s1 % 8 == 0
    for y in range(0, min(p0.shape[0], p1.shape[0]), 8):
        for x in range(0, min(s0, s1), 8):
            if x >= bounds_w or y >= bounds_h:
                continue
            assert np.array_equal(p0[y:y + 8, x:x + 8], p1[y:y + 8, x:x + 8]), (x, y)


@pytest.mark.parametrize(
    "base",
    [
        "video-001", "video-001.q50.410", "video-001.q50.411", "video-001.q50.420", "video-001.q50.422",
        "video-001.q50.440", "video-001.q50.444", "video-005.gray.q50", "video-005.gray.q50.2x2",
        "video-001.separate.dc.progression",
    ],
)
def test_decode_plus_progressive(fixtures_dir, base):
    """decoder.zig:1843-1920"""
    m0 = O.decode(_read(fixtures_dir, base + ".jpeg"))
    m1 = O.decode(_read(fixtures_dir, base + ".progressive.jpeg"))
    assert m0.bounds() == m1.bounds() == (0, 0, 150, 103)
    assert m0.variant == m1.variant
    if m0.variant == O.GRAY:
        _check_planes(150, 103, m0.pix, m1.pix)
    else:
        assert m0.variant == O.YCBCR
        _check_planes(150, 103, m0.y, m1.y)
        _check_planes(150, 103, m0.cb, m1.cb)
        _check_planes(150, 103, m0.cr, m1.cr)


@pytest.mark.parametrize(
    "name",
    ["video-001.cmyk", "video-001.221212", "video-005.gray", "video-001.rgb", "video-001.separate.dc.progression"],
)
def test_decode_assorted(fixtures_dir, name):
    """decoder.zig:1922-1940"""
    O.decode(_read(fixtures_dir, name + ".jpeg"))


def test_truncated_sos(fixtures_dir):
    """decoder.zig:1942-1963"""
    b = _read(fixtures_dir, "video-005.gray.q50.jpeg")
    i = b.index(b"\xff\xda") + 2
    for k in range(i, min(i + 10, len(b))):
        with pytest.raises(O.OracleError) as e:
            O.decode(b[:k])
        assert e.value.name == "UnexpectedEof"


def test_large_image_with_short_data(golden_dir):
    """decoder.zig:1965-2027"""
    with pytest.raises(O.OracleError) as e:
        O.decode(_read(golden_dir, "fuzz_issue10413.bin"))
    assert e.value.name == "UnexpectedEof"


def test_padded_rst_marker(golden_dir):
    """decoder.zig:2029-2205"""
    O.decode(_read(golden_dir, "padded_rst_issue28717.jpg"))


def test_issue56724(fixtures_dir):
    """decoder.zig:2207-2226"""
    with pytest.raises(O.OracleError) as e:
        O.decode(_read(fixtures_dir, "video-001.jpeg")[:24])
    assert e.value.name == "UnexpectedEof"


BAD_RST_CASES = [
    (True, b""), (True, b"\x00"), (True, b"\x61"), (True, b"\x61\x62\x63\xff\x00\x64"), (True, b"\xff"),
    (True, b"\xff\x00"), (True, b"\xff\xff\xff\x00\xff\x00\x00\xff\xff\xff"),
    (False, b"\xff\x03"), (False, b"\xff\xd5"), (False, b"\xff\xff\xd5"),
]


@pytest.mark.parametrize("want_pass,infix", BAD_RST_CASES)
def test_bad_restart_marker(fixtures_dir, want_pass, infix):
    """decoder.zig:2228-2279"""
    data = _read(fixtures_dir, "video-001.restart2.jpeg")
    assert len(data) == 4855 and data[2816:2818] == b"\xff\xd1"
    spliced = data[:2816] + infix + data[2816:]
    if want_pass:
        img = O.decode(spliced)
        # junk before an RST is skipped without touching the pixels
        assert np.array_equal(img.rgbaPixels(), O.decode(data).rgbaPixels())
    else:
        with pytest.raises(O.OracleError) as e:
            O.decode(spliced)
        assert e.value.name == "BadRSTMarker"


def test_decode_config(fixtures_dir):
    """decoder.zig:178-218"""
    assert O.decode_config(_read(fixtures_dir, "video-001.jpeg")) == (150, 103, "YCbCr")
    assert O.decode_config(_read(fixtures_dir, "video-005.gray.jpeg")) == (150, 103, "Gray")
    assert O.decode_config(_read(fixtures_dir, "video-001.cmyk.jpeg")) == (150, 103, "YCbCr")
    assert O.decode_config(_read(fixtures_dir, "iceberg.jpg")) == (2048, 2048, "YCbCr")


def test_not_a_jpeg():
    with pytest.raises(O.OracleError) as e:
        O.decode(b"\x89PNG\r\n\x1a\n")
    assert e.value.name == "InvalidSOIMarker"
    with pytest.raises(O.OracleError) as e:
        O.decode(b"")
    assert e.value.name == "UnexpectedEof"


def test_sanity_vs_libjpeg(fixtures_dir):
    """Loose bound only (libjpeg-turbo is NOT an oracle: different IDCT, fancy upsampling)."""
    import io

    from PIL import Image as PILImage

    for name, bound in [("video-005.gray.jpeg", 1), ("video-001.q50.444.jpeg", 3), ("video-001.jpeg", 3)]:
        data = _read(fixtures_dir, name)
        ours = O.decode(data).rgbaPixels()[:, :, :3].astype(int)
        ref = np.asarray(PILImage.open(io.BytesIO(data)).convert("RGB")).astype(int)
        assert np.abs(ours - ref).max() <= bound, name


def test_idct_known_answers():
    """idct.zig:77-201: zero block -> zero; pure DC -> constant; row shortcut is value-neutral."""
    z = np.zeros(64, np.int32)
    assert not O.idct(z).any()
    b = z.copy()
    b[0] = 8 * 16
    # row pass: dc<<3 ; column pass: ((v<<8)+8192 ...)>>14 of a constant column
    out = O.idct(b)
    assert (out == out[0]).all() and out[0] == 16
    rng = np.random.default_rng(7)
    for _ in range(200):
        b = rng.integers(-1024, 1024, 64).astype(np.int32)
        b[rng.integers(0, 8) * 8 + 1: (rng.integers(0, 8) + 1) * 8] = 0
        ref = _idct_general(b)
        assert np.array_equal(O.idct(b), ref)


def _idct_general(src):
    """Same arithmetic WITHOUT the all-AC-zero row shortcut (SURVEY B3 says it is value-neutral)."""
    w1, w2, w3, w5, w6, w7, r2 = 2841, 2676, 2408, 1609, 1108, 565, 181
    s = [int(v) for v in src]

    def w(v):  # wrap to i32
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v & 0x80000000 else v

    for y in range(8):
        r = s[y * 8:y * 8 + 8]
        x0 = w((r[0] << 11) + 128); x1 = w(r[4] << 11); x2 = r[6]; x3 = r[2]; x4 = r[1]; x5 = r[7]; x6 = r[5]; x7 = r[3]
        x8 = w(w7 * (x4 + x5)); x4 = w(x8 + (w1 - w7) * x4); x5 = w(x8 - (w1 + w7) * x5)
        x8 = w(w3 * (x6 + x7)); x6 = w(x8 - (w3 - w5) * x6); x7 = w(x8 - (w3 + w5) * x7)
        x8 = w(x0 + x1); x0 = w(x0 - x1); x1 = w(w6 * (x3 + x2)); x2 = w(x1 - (w2 + w6) * x2); x3 = w(x1 + (w2 - w6) * x3)
        x1 = w(x4 + x6); x4 = w(x4 - x6); x6 = w(x5 + x7); x5 = w(x5 - x7)
        x7 = w(x8 + x3); x8 = w(x8 - x3); x3 = w(x0 + x2); x0 = w(x0 - x2)
        x2 = w(r2 * (x4 + x5) + 128) >> 8; x4 = w(r2 * (x4 - x5) + 128) >> 8
        s[y * 8:y * 8 + 8] = [w(x7 + x1) >> 8, w(x3 + x2) >> 8, w(x0 + x4) >> 8, w(x8 + x6) >> 8,
                              w(x8 - x6) >> 8, w(x0 - x4) >> 8, w(x3 - x2) >> 8, w(x7 - x1) >> 8]
    for x in range(8):
        c = [s[x + 8 * k] for k in range(8)]
        y0 = w((c[0] << 8) + 8192); y1 = w(c[4] << 8); y2 = c[6]; y3 = c[2]; y4 = c[1]; y5 = c[7]; y6 = c[5]; y7 = c[3]
        y8 = w(w7 * (y4 + y5) + 4); y4 = w(y8 + (w1 - w7) * y4) >> 3; y5 = w(y8 - (w1 + w7) * y5) >> 3
        y8 = w(w3 * (y6 + y7) + 4); y6 = w(y8 - (w3 - w5) * y6) >> 3; y7 = w(y8 - (w3 + w5) * y7) >> 3
        y8 = w(y0 + y1); y0 = w(y0 - y1); y1 = w(w6 * (y3 + y2) + 4); y2 = w(y1 - (w2 + w6) * y2) >> 3; y3 = w(y1 + (w2 - w6) * y3) >> 3
        y1 = w(y4 + y6); y4 = w(y4 - y6); y6 = w(y5 + y7); y5 = w(y5 - y7)
        y7 = w(y8 + y3); y8 = w(y8 - y3); y3 = w(y0 + y2); y0 = w(y0 - y2)
        y2 = w(r2 * (y4 + y5) + 128) >> 8; y4 = w(r2 * (y4 - y5) + 128) >> 8
        outs = [w(y7 + y1) >> 14, w(y3 + y2) >> 14, w(y0 + y4) >> 14, w(y8 + y6) >> 14,
                w(y8 - y6) >> 14, w(y0 - y4) >> 14, w(y3 - y2) >> 14, w(y7 - y1) >> 14]
        for k in range(8):
            s[x + 8 * k] = outs[k]
    return np.array(s, dtype=np.int32)


def test_colour_formulas():
    """color.zig:90-126 + image.zig:122-125 against the closed forms of SURVEY A.6."""
    rng = np.random.default_rng(3)
    for y, cb, cr in rng.integers(0, 256, (2000, 3)):
        yy = int(y) * 0x10101
        c1, c2 = int(cb) - 128, int(cr) - 128

        def ch(v):
            return 0 if v < 0 else (255 if v >= (1 << 24) else v >> 16)

        want = (ch(yy + 91881 * c2), ch(yy - 22554 * c1 - 46802 * c2), ch(yy + 116130 * c1), 255)
        assert O.ycbcr_to_rgba8(int(y), int(cb), int(cr)) == want
    for c, m, y, k in rng.integers(0, 256, (2000, 4)):
        wk = 0xFFFF - int(k) * 0x101
        want = tuple(((0xFFFF - int(v) * 0x101) * wk // 0xFFFF) >> 8 for v in (c, m, y)) + (255,)
        assert O.cmyk_to_rgba8(int(c), int(m), int(y), int(k)) == want


def test_tap_matches_image(fixtures_dir):
    """The coefficient tap (used by the entropy-kernel tests) reproduces the planes through dequant+IDCT."""
    img, recs = O.decode(_read(fixtures_dir, "video-001.q50.420.jpeg"), tap=True)
    assert recs.shape == (420, 67)
    assert set(recs[:, 0]) == {0, 1, 2}
