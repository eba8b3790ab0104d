"""zpix_b200 -- B200-native batched JPEG decode behind zpix's `jpeg.load` / `Image` API.

The product is libzpixcuda.so (CUDA kernels + C ABI, include/zpix_cuda.h).  This package is the
host-side mirror of the reference's module API for that path (`jpeg`, `image`, `color`), used by
the tests and the benchmark.  Importing it loads the CUDA library and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401  (raises ImportError with the build command if the .so is missing)
from . import color, image, jpeg  # noqa: F401

__all__ = ["jpeg", "image", "color"]


def fromBuffer(buffer: bytes):
    """zpix.fromBuffer (reference src/root.zig:35-40), JPEG only in this build."""
    if jpeg.probeBuffer(buffer):
        return jpeg.loadFromBuffer(buffer)
    raise ValueError("error.UnknownImageFormat")


def fromFilePath(path: str):
    """zpix.fromFilePath (reference src/root.zig:24-32), JPEG only in this build."""
    if jpeg.probePath(path):
        return jpeg.load(path)
    raise ValueError("error.UnknownImageFormat")


def fromBuffers(buffers, ctx=None):
    """Batch form of fromBuffer: probe every buffer, send the JPEGs to the GPU batch path in one
    call and answer `error.UnknownImageFormat` for the rest (PNG / QOI / BMP stay with the reference's CPU
    decoders, which are out of scope here).  Returns one Image or one exception instance per input."""
    is_jpeg = [jpeg.probeBuffer(b) for b in buffers]
    decoded = iter(jpeg.decodeBatch([b for b, j in zip(buffers, is_jpeg) if j], ctx))
    return [next(decoded) if j else ValueError("error.UnknownImageFormat") for j in is_jpeg]
