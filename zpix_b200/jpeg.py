"""Host-side mirror of zpix's `jpeg` module (reference src/jpeg/root.zig) over the CUDA library.

    load(path)              src/jpeg/root.zig:36   -> image.Image   (native variant, as the reference)
    loadFromBuffer(buf)     src/jpeg/root.zig:10
    decode(reader)          src/jpeg/decoder.zig:155 (reader = anything with .read())
    decodeConfig(reader)    src/jpeg/decoder.zig:178
    probeBuffer / probePath src/jpeg/root.zig:17-34
    decodeBatch(buffers)    NEW, beside loadFromBuffer: list[Image{.RGBA}] (rgbaPixels semantics)
    loadBatch(paths)        NEW, beside load

Errors are raised as JpegError carrying the Zig error name (`error.UnexpectedEof` -> "UnexpectedEof").
Everything is decoded by libzpixcuda.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import image as _image
from ._lib import ZpxImageInfo, ZpxTiming, lib

VARIANTS = {0: "Gray", 1: "YCbCr", 2: "RGBA", 3: "CMYK"}


class JpegError(Exception):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        self.name = lib.zpx_error_name(code).decode()
        super().__init__(f"error.{self.name}" + (f" ({detail})" if detail else ""))


def _check(ctx, code):
    if code != 0:
        detail = ""
        if code == 100 and ctx is not None:
            detail = lib.zpx_last_cuda_error_string(ctx).decode()
        raise JpegError(code, detail)


class Context:
    """Owns the zpx_ctx (device memory, streams).  Not thread-safe (one per thread)."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        if devices:
            arr = (C.c_int32 * len(devices))(*devices)
            code = lib.zpx_ctx_create(arr, len(devices), C.byref(self._h))
        else:
            code = lib.zpx_ctx_create(None, 0, C.byref(self._h))
        _check(None, code)

    def set_option(self, opt: int, value: int):
        _check(self._h, lib.zpx_ctx_set_option(self._h, opt, value))

    @property
    def handle(self):
        return self._h

    @property
    def num_devices(self):
        return lib.zpx_ctx_num_devices(self._h)

    @property
    def kernel_launches(self):
        return lib.zpx_ctx_kernel_launches(self._h)

    def close(self):
        if self._h:
            lib.zpx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class Batch:
    """open -> upload -> decode -> fetch, the staged form of decodeBatch (used by the benchmark)."""

    def __init__(self, ctx: Context, buffers: Sequence[bytes]):
        self.ctx = ctx
        self.n = len(buffers)
        self._keep = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in buffers]
        ptrs = (C.c_void_p * max(self.n, 1))(*[a.ctypes.data if a.size else None for a in self._keep])
        lens = (C.c_size_t * max(self.n, 1))(*[a.size for a in self._keep])
        self._h = C.c_void_p()
        _check(ctx.handle, lib.zpx_batch_open(ctx.handle, ptrs, lens, self.n, C.byref(self._h)))

    def info(self, i) -> ZpxImageInfo:
        inf = ZpxImageInfo()
        _check(self.ctx.handle, lib.zpx_batch_info(self._h, i, C.byref(inf)))
        return inf

    def infos(self):
        return [self.info(i) for i in range(self.n)]

    def upload(self):
        _check(self.ctx.handle, lib.zpx_batch_upload(self._h))

    def decode(self, stream: int = 0):
        _check(self.ctx.handle, lib.zpx_batch_decode(self._h, C.c_void_p(stream) if stream else None))

    def status(self) -> List[int]:
        st = (C.c_int32 * max(self.n, 1))()
        _check(self.ctx.handle, lib.zpx_batch_status(self._h, st))
        return list(st)[: self.n]

    def timing(self, device_index=0) -> dict:
        t = ZpxTiming()
        _check(self.ctx.handle, lib.zpx_batch_timing(self._h, device_index, C.byref(t)))
        return t.as_dict()

    def fetch_rgba(self, outs: Optional[List[Optional[np.ndarray]]] = None):
        """Returns (list of HxWx4 uint8 arrays or None, list of status)."""
        infos = self.infos()
        if outs is None:
            outs = [np.empty((inf.height, inf.width, 4), np.uint8) if inf.status == 0 else None for inf in infos]
        ptrs = (C.c_void_p * max(self.n, 1))(*[o.ctypes.data if o is not None else None for o in outs])
        st = (C.c_int32 * max(self.n, 1))()
        _check(self.ctx.handle, lib.zpx_batch_fetch_rgba(self._h, ptrs, None, st))
        st = list(st)[: self.n]
        return [o if s == 0 else None for o, s in zip(outs, st)], st

    def fetch_native(self):
        infos = self.infos()
        outs = [np.empty(inf.native_len, np.uint8) if inf.status == 0 else None for inf in infos]
        ptrs = (C.c_void_p * max(self.n, 1))(*[o.ctypes.data if o is not None else None for o in outs])
        st = (C.c_int32 * max(self.n, 1))()
        _check(self.ctx.handle, lib.zpx_batch_fetch_native(self._h, ptrs, st))
        return outs, list(st)[: self.n]

    def coefficients(self, i) -> np.ndarray:
        nb = C.c_size_t()
        _check(self.ctx.handle, lib.zpx_batch_fetch_coefficients(self._h, i, None, 0, C.byref(nb)))
        out = np.empty((nb.value, 64), np.int16)
        _check(self.ctx.handle, lib.zpx_batch_fetch_coefficients(self._h, i, out.ctypes.data, nb.value, C.byref(nb)))
        return out

    def device_rgba_ptr(self, i) -> int:
        return lib.zpx_batch_device_rgba(self._h, i) or 0

    def close(self):
        if self._h:
            lib.zpx_batch_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class SyntheticBatch(Batch):
    """Test hook (zpx_batch_open_synthetic): ONE sequential frame given by its geometry and quantisers; its int16
    coefficient blocks are injected with set_coefficients(), so decode() runs the reconstruction kernels only.
    comp_hv: per component (h, v); quant_zz: (ncomp, 64) zig-zag order; mode 0 gray, 1 YCbCr, 2 RGB, 3 CMYK, 4 YCbCrK."""

    def __init__(self, ctx: Context, width: int, height: int, comp_hv, quant_zz, mode: int):
        self.ctx = ctx
        self.n = 1
        self._keep = []
        hv = (C.c_uint8 * len(comp_hv))(*[(h << 4) | v for h, v in comp_hv])
        q = np.ascontiguousarray(quant_zz, dtype=np.uint16).reshape(len(comp_hv), 64)
        self._h = C.c_void_p()
        _check(ctx.handle, lib.zpx_batch_open_synthetic(ctx.handle, width, height, len(comp_hv), hv, q.ctypes.data, mode, C.byref(self._h)))

    def set_coefficients(self, blocks: np.ndarray):
        b = np.ascontiguousarray(blocks, dtype=np.int16).reshape(-1, 64)
        _check(self.ctx.handle, lib.zpx_batch_set_coefficients(self._h, 0, b.ctypes.data, b.shape[0]))


def test_colour(ctx: Context, mode: int, samples: np.ndarray) -> np.ndarray:
    """Test hook (zpx_test_colour): the kernels' colour functions on free-standing samples -> (n, 4) uint8."""
    a = np.ascontiguousarray(samples, dtype=np.uint8)
    n = a.shape[0]
    out = np.empty((n, 4), np.uint8)
    _check(ctx.handle, lib.zpx_test_colour(ctx.handle, mode, a.ctypes.data, n, out.ctypes.data))
    return out


def _rgba_image(inf: ZpxImageInfo, rgba: np.ndarray) -> _image.Image:
    rect = _image.Rectangle.init(0, 0, inf.width, inf.height)
    flat = rgba.reshape(-1)
    return _image.Image("RGBA", _image.RGBAImage(flat, 4 * inf.width, rect), device_rgba=flat)


def decodeBatch(buffers: Sequence[bytes], ctx: Optional[Context] = None, raise_on_error: bool = False):
    """NEW entry beside loadFromBuffer: decode a batch to `Image{.RGBA}` (rgbaPixels bytes).
    Returns a list with an Image per input, or a JpegError instance for inputs that failed."""
    ctx = ctx or default_context()
    with Batch(ctx, buffers) as b:
        b.upload()
        b.decode()
        outs, st = b.fetch_rgba()
        infos = b.infos()
    res = []
    for inf, o, s in zip(infos, outs, st):
        if s != 0:
            if raise_on_error:
                raise JpegError(s)
            res.append(JpegError(s))
        else:
            res.append(_rgba_image(inf, o))
    return res


def decodeBatchOneCall(buffers: Sequence[bytes], ctx: Optional[Context] = None):
    """The single C-ABI call `zpx_decode_batch_rgba` (what the Zig `jpeg.decodeBatch` binds): header
    parse, upload, kernels and download, pipelined in chunks for large batches.
    Returns (list of HxWx4 uint8 arrays or None, list of status codes)."""
    ctx = ctx or default_context()
    n = len(buffers)
    keep = [np.frombuffer(b, dtype=np.uint8) for b in buffers]
    infos = []
    for a in keep:
        inf = ZpxImageInfo()
        lib.zpx_probe(a.ctypes.data if a.size else None, a.size, C.byref(inf))
        infos.append(inf)
    outs = [np.empty((i.height, i.width, 4), np.uint8) if (i.status == 0 and i.width > 0 and i.height > 0) else None for i in infos]
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in keep])
    lens = (C.c_size_t * max(n, 1))(*[a.size for a in keep])
    optr = (C.c_void_p * max(n, 1))(*[o.ctypes.data if o is not None else None for o in outs])
    st = (C.c_int32 * max(n, 1))()
    _check(ctx.handle, lib.zpx_decode_batch_rgba(ctx.handle, ptrs, lens, n, optr, None, st))
    st = list(st)[:n]
    return [o if s == 0 else None for o, s in zip(outs, st)], st


def decodeBatchNative(buffers: Sequence[bytes], ctx: Optional[Context] = None):
    """The single C-ABI call `zpx_decode_batch_native`: like decodeBatchOneCall, but every image comes back as the
    variant jpeg.load returns (Image{.YCbCr} / {.Gray} planes with makeImg's strides, {.RGBA}, {.CMYK}).
    Returns a list with an Image per input, or a JpegError instance for inputs that failed."""
    ctx = ctx or default_context()
    n = len(buffers)
    keep = [np.frombuffer(b, dtype=np.uint8) for b in buffers]
    infos = []
    for a in keep:
        inf = ZpxImageInfo()
        lib.zpx_probe(a.ctypes.data if a.size else None, a.size, C.byref(inf))
        infos.append(inf)
    outs = [np.empty(i.native_len, np.uint8) if (i.status == 0 and i.width > 0 and i.height > 0) else None for i in infos]
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in keep])
    lens = (C.c_size_t * max(n, 1))(*[a.size for a in keep])
    optr = (C.c_void_p * max(n, 1))(*[o.ctypes.data if o is not None else None for o in outs])
    st = (C.c_int32 * max(n, 1))()
    _check(ctx.handle, lib.zpx_decode_batch_native(ctx.handle, ptrs, lens, n, optr, st))
    res = []
    for inf, o, s in zip(infos, outs, list(st)[:n]):
        if s != 0 or o is None:
            res.append(JpegError(s if s != 0 else inf.status))
            continue
        img = _native_image(inf, o, np.zeros(0, np.uint8))
        img._device_rgba = None  # rgbaPixels() of these images runs the reference's per-pixel formulas on the host
        res.append(img)
    return res


def loadBatch(paths: Sequence[str], ctx: Optional[Context] = None, raise_on_error: bool = False):
    bufs = []
    for p in paths:
        with open(p, "rb") as f:
            bufs.append(f.read())
    return decodeBatch(bufs, ctx, raise_on_error)


def _native_image(inf: ZpxImageInfo, native: np.ndarray, rgba: np.ndarray) -> _image.Image:
    rect = _image.Rectangle.init(0, 0, inf.width, inf.height)
    v = VARIANTS[inf.variant]
    flat = rgba.reshape(-1)
    if v == "Gray":
        return _image.Image("Gray", _image.GrayImage(native, inf.y_stride, rect), device_rgba=flat)
    if v == "YCbCr":
        ratio = _image.YCbCrSubsample(inf.subsample_ratio)
        y = native[: inf.native_cb_off]
        cb = native[inf.native_cb_off: inf.native_cr_off]
        cr = native[inf.native_cr_off:]
        return _image.Image("YCbCr", _image.YCbCrImage(y, cb, cr, inf.y_stride, inf.c_stride, ratio, rect, native), device_rgba=flat)
    if v == "RGBA":
        return _image.Image("RGBA", _image.RGBAImage(native, 4 * inf.width, rect), device_rgba=flat)
    return _image.Image("CMYK", _image.CMYKImage(native, 4 * inf.width, rect), device_rgba=flat)


def loadFromBuffer(buffer: bytes, ctx: Optional[Context] = None) -> _image.Image:
    """src/jpeg/root.zig:10.  Returns the same Image variant the reference returns (.Gray/.YCbCr/.RGBA),
    .CMYK for 4-component frames), planes / interleave computed on the GPU."""
    ctx = ctx or default_context()
    ctx.set_option(9, 1)  # ZPX_OPT_NATIVE_PLANES: the fused kernel writes the planes beside the RGBA
    try:
        with Batch(ctx, [buffer]) as b:
            inf = b.info(0)
            if inf.status != 0:
                raise JpegError(inf.status)
            b.upload()
            b.decode()
            outs, st = b.fetch_rgba()
            if st[0] != 0:
                raise JpegError(st[0])
            nat, st2 = b.fetch_native()
            if st2[0] != 0:
                raise JpegError(st2[0])
            return _native_image(inf, nat[0], outs[0])
    finally:
        ctx.set_option(9, 0)


def load(path: str, ctx: Optional[Context] = None) -> _image.Image:
    """src/jpeg/root.zig:36"""
    with open(path, "rb") as f:
        return loadFromBuffer(f.read(), ctx)


def decode(reader, ctx: Optional[Context] = None) -> _image.Image:
    """src/jpeg/decoder.zig:155 (reader: object with .read())"""
    return loadFromBuffer(reader.read(), ctx)


def decodeConfig(reader) -> _image.Config:
    """src/jpeg/decoder.zig:178-218; host only."""
    data = reader.read() if hasattr(reader, "read") else bytes(reader)
    inf = ZpxImageInfo()
    arr = np.frombuffer(data, np.uint8)
    code = lib.zpx_probe(arr.ctypes.data if arr.size else None, arr.size, C.byref(inf))
    if code != 0:
        raise JpegError(code)
    return _image.Config(inf.width, inf.height, "Gray" if inf.num_components == 1 else "YCbCr")


def probeBuffer(buffer: bytes) -> bool:
    """src/jpeg/root.zig:17-20"""
    return len(buffer) >= 2 and buffer[0] == 0xFF and buffer[1] == 0xD8


def probePath(path: str) -> bool:
    """src/jpeg/root.zig:23-34"""
    with open(path, "rb") as f:
        return probeBuffer(f.read(2))
