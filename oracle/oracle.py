"""ctypes binding of the CPU oracle (oracle/zpix_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(zpix_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzpix_oracle.so")

GRAY, YCBCR, RGBA, CMYK = 0, 1, 2, 3
VARIANT_NAMES = {GRAY: "Gray", YCBCR: "YCbCr", RGBA: "RGBA", CMYK: "CMYK"}
RATIO_NAMES = {0: "Ratio444", 1: "Ratio422", 2: "Ratio420", 3: "Ratio440", 4: "Ratio411", 5: "Ratio410"}


class _ZoImage(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("pixels", C.POINTER(C.c_uint8)),
        ("pixels_len", C.c_size_t),
        ("pix", C.POINTER(C.c_uint8)),
        ("stride", C.c_size_t),
        ("y", C.POINTER(C.c_uint8)),
        ("cb", C.POINTER(C.c_uint8)),
        ("cr", C.POINTER(C.c_uint8)),
        ("y_stride", C.c_size_t),
        ("c_stride", C.c_size_t),
        ("subsample_ratio", C.c_int32),
        ("ycck_intent", C.c_int32),
        ("eob_carry", C.c_int32),
        ("pad0", C.c_int32),
    ]


class _ZoBlockRec(C.Structure):
    _fields_ = [("comp", C.c_int32), ("bx", C.c_int32), ("by", C.c_int32), ("coef", C.c_int32 * 64)]


class _ZoTap(C.Structure):
    _fields_ = [("recs", C.POINTER(_ZoBlockRec)), ("cap", C.c_size_t), ("count", C.c_size_t)]


class _ZoConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("color_model", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc if the .so is missing or stale."""
    src = os.path.join(_HERE, "zpix_oracle.c")
    hdr = os.path.join(_HERE, "zpix_oracle.h")
    stale = (
        force
        or not os.path.exists(_SO)
        or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libzpix_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.zo_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_ZoImage)]
        L.zo_decode.restype = C.c_int
        L.zo_decode_tap.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_ZoImage), C.POINTER(_ZoTap)]
        L.zo_decode_tap.restype = C.c_int
        L.zo_decode_config.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_ZoConfig)]
        L.zo_decode_config.restype = C.c_int
        L.zo_rgba_pixels.argtypes = [C.POINTER(_ZoImage), C.c_void_p]
        L.zo_rgba_pixels.restype = None
        L.zo_last_eob_carry.argtypes = []
        L.zo_last_eob_carry.restype = C.c_int
        L.zo_last_coef_overflow.argtypes = []
        L.zo_last_coef_overflow.restype = C.c_int
        L.zo_free.argtypes = [C.POINTER(_ZoImage)]
        L.zo_free.restype = None
        L.zo_error_name.argtypes = [C.c_int]
        L.zo_error_name.restype = C.c_char_p
        L.zo_idct.argtypes = [C.c_void_p]
        L.zo_idct.restype = None
        L.zo_ycbcr_to_rgba16.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
        L.zo_cmyk_to_rgba16.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
        L.zo_load_rgba.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.zo_load_rgba.restype = C.c_int
        L.zo_reconstruct_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.zo_reconstruct_blocks.restype = None
        L.zo_ycbcr_to_rgba8_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.zo_ycbcr_to_rgba8_batch.restype = None
        L.zo_cmyk_to_rgba8_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.zo_cmyk_to_rgba8_batch.restype = None
        L.zo_compare_rgba.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.zo_compare_rgba.restype = C.c_int
        _lib = L
    return _lib


class OracleError(Exception):
    """Carries the Zig error name the reference decoder would return."""

    def __init__(self, code: int):
        self.code = code
        self.name = lib().zo_error_name(code).decode()
        super().__init__(self.name)


class Image:
    """Result of jpeg.load: one of the four variants the JPEG decoder returns."""

    def __init__(self, raw: _ZoImage):
        self.variant = raw.variant
        self.variant_name = VARIANT_NAMES[raw.variant]
        self.width, self.height = raw.width, raw.height
        self.ycck_intent = bool(raw.ycck_intent)
        self.eob_carry = bool(raw.eob_carry)
        self.subsample_ratio = RATIO_NAMES.get(raw.subsample_ratio) if raw.variant == YCBCR else None
        L = lib()
        rgba = np.empty((raw.height, raw.width, 4), dtype=np.uint8)
        L.zo_rgba_pixels(C.byref(raw), rgba.ctypes.data)
        self._rgba = rgba
        buf = np.ctypeslib.as_array(raw.pixels, shape=(raw.pixels_len,)).copy() if raw.pixels_len else np.zeros(0, np.uint8)
        self.pixels = buf
        base = C.addressof(raw.pixels.contents) if raw.pixels_len else 0

        def off(p):
            return C.addressof(p.contents) - base

        if raw.variant == YCBCR:
            self.y_stride, self.c_stride = raw.y_stride, raw.c_stride
            yo, cbo, cro = off(raw.y), off(raw.cb), off(raw.cr)
            ylen = cbo - yo
            clen = cro - cbo
            self.y = buf[yo:yo + ylen].reshape(-1, raw.y_stride)
            self.cb = buf[cbo:cbo + clen].reshape(-1, raw.c_stride)
            self.cr = buf[cro:cro + clen].reshape(-1, raw.c_stride)
        else:
            self.stride = raw.stride
            self.pix = buf.reshape(-1, raw.stride) if raw.stride else buf

    def bounds(self):
        return (0, 0, self.width, self.height)

    def rgbaPixels(self) -> np.ndarray:
        """Image.rgbaPixels (image.zig:103): tight H x W x 4 uint8."""
        return self._rgba


def last_eob_carry() -> bool:
    """True if the latest decode() met a scan that started inside an End-Of-Band run (also when it raised)."""
    return bool(lib().zo_last_eob_carry())


def last_coef_overflow() -> bool:
    """True if a coefficient left the int16 range during the latest decode() (also when it raised)."""
    return bool(lib().zo_last_coef_overflow())


def decode(data: bytes, tap: bool = False):
    """jpeg.loadFromBuffer.  Returns Image (and the block records when tap=True)."""
    L = lib()
    raw = _ZoImage()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data) if len(data) else (C.c_uint8 * 1)()
    if not tap:
        e = L.zo_decode(buf, len(data), C.byref(raw))
        if e != 0:
            raise OracleError(e)
        try:
            return Image(raw)
        finally:
            L.zo_free(C.byref(raw))
    t = _ZoTap(None, 0, 0)
    e = L.zo_decode_tap(buf, len(data), C.byref(raw), C.byref(t))
    if e != 0:
        raise OracleError(e)
    L.zo_free(C.byref(raw))
    n = t.count
    recs = (_ZoBlockRec * max(n, 1))()
    t = _ZoTap(recs, n, 0)
    e = L.zo_decode_tap(buf, len(data), C.byref(raw), C.byref(t))
    if e != 0:
        raise OracleError(e)
    try:
        img = Image(raw)
    finally:
        L.zo_free(C.byref(raw))
    arr = np.frombuffer(recs, dtype=np.int32).reshape(max(n, 1), 67)[:n]
    return img, arr  # columns: comp, bx, by, coef[64]


def load(path: str) -> Image:
    """jpeg.load"""
    with open(path, "rb") as f:
        return decode(f.read())


def decode_config(data: bytes):
    L = lib()
    cfg = _ZoConfig()
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data.ljust(1, b"\0")) if len(data) == 0 else (C.c_uint8 * len(data)).from_buffer_copy(data)
    e = L.zo_decode_config(buf, len(data), C.byref(cfg))
    if e != 0:
        raise OracleError(e)
    return cfg.width, cfg.height, VARIANT_NAMES[cfg.color_model]


def load_rgba_timed(data: bytes) -> int:
    """decode + rgbaPixels without copying the result to Python (CPU baseline). Returns error code."""
    L = lib()
    return L.zo_load_rgba(data, len(data), None, 0, None, None)


def idct(block: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(block, dtype=np.int32).copy()
    lib().zo_idct(b.ctypes.data)
    return b


def ycbcr_to_rgba8(y: int, cb: int, cr: int):
    out = (C.c_uint32 * 4)()
    lib().zo_ycbcr_to_rgba16(y, cb, cr, out)
    return tuple(v >> 8 for v in out)


def cmyk_to_rgba8(c: int, m: int, y: int, k: int):
    out = (C.c_uint32 * 4)()
    lib().zo_cmyk_to_rgba16(c, m, y, k, out)
    return tuple(v >> 8 for v in out)


def reconstruct_blocks(coef: np.ndarray, quant_zz: np.ndarray) -> np.ndarray:
    """reconstructBlock (decoder.zig:1553-1634) on free-standing blocks: coef (n, 64) natural order, quant_zz (64,)
    zig-zag order -> (n, 8, 8) uint8."""
    c = np.ascontiguousarray(coef, dtype=np.int32).reshape(-1, 64)
    q = np.ascontiguousarray(quant_zz, dtype=np.int32).reshape(64)
    out = np.empty((c.shape[0], 8, 8), np.uint8)
    lib().zo_reconstruct_blocks(c.ctypes.data, q.ctypes.data, c.shape[0], out.ctypes.data)
    return out


def ycbcr_to_rgba8_batch(ycc: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(ycc, dtype=np.uint8).reshape(-1, 3)
    out = np.empty((a.shape[0], 4), np.uint8)
    lib().zo_ycbcr_to_rgba8_batch(a.ctypes.data, a.shape[0], out.ctypes.data)
    return out


def cmyk_to_rgba8_batch(cmyk: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(cmyk, dtype=np.uint8).reshape(-1, 4)
    out = np.empty((a.shape[0], 4), np.uint8)
    lib().zo_cmyk_to_rgba8_batch(a.ctypes.data, a.shape[0], out.ctypes.data)
    return out


def compare_rgba(data: bytes, got_ptr: int, got_len: int) -> int:
    """jpeg.load + rgbaPixels of `data` against the bytes at got_ptr: 0 identical, 1 different, < 0 -(error code).
    Releases the GIL: callable from a thread pool."""
    return lib().zo_compare_rgba(data, len(data), got_ptr, got_len)
