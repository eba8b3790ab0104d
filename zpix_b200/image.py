"""Host-side mirror of zpix's `image` module (reference src/image/image.zig, geometry.zig).

Same names, fields and semantics as the Zig types the JPEG path returns:
`Image` is the tagged union with `bounds()`, `at(x, y)`, `rgbaPixels()`, `free()`;
payloads are `GrayImage`, `YCbCrImage`, `RGBAImage`, `CMYKImage`.  Pixel storage is a numpy
uint8 array (the Zig `pixels: []u8` slice).  `rgbaPixels()` of an image produced by the GPU batch
path returns the bytes the device computed (image.zig:103-130 runs on the GPU, not here).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional

import numpy as np

from .color import Color


@dataclass(frozen=True)
class Point:
    x: int
    y: int

    def In(self, r: "Rectangle") -> bool:  # geometry.zig:6-8
        return r.min.x <= self.x < r.max.x and r.min.y <= self.y < r.max.y


@dataclass(frozen=True)
class Rectangle:
    min: Point
    max: Point

    @staticmethod
    def init(x0, y0, x1, y1):  # geometry.zig:19-30
        return Rectangle(Point(min(x0, x1), min(y0, y1)), Point(max(x0, x1), max(y0, y1)))

    def dX(self):
        return self.max.x - self.min.x

    def dY(self):
        return self.max.y - self.min.y

    def size(self):
        return Point(self.dX(), self.dY())

    def Intersect(self, o: "Rectangle") -> Optional["Rectangle"]:  # geometry.zig:45-54
        x0, y0 = max(self.min.x, o.min.x), max(self.min.y, o.min.y)
        x1, y1 = min(self.max.x, o.max.x), min(self.max.y, o.max.y)
        if x0 >= x1 or y0 >= y1:
            return None
        return Rectangle.init(x0, y0, x1, y1)


class YCbCrSubsample(Enum):  # image.zig:465-472
    Ratio444 = 0
    Ratio422 = 1
    Ratio420 = 2
    Ratio440 = 3
    Ratio411 = 4
    Ratio410 = 5


@dataclass
class Config:  # image.zig:16-20
    width: int
    height: int
    color_model: str


@dataclass
class GrayImage:  # image.zig:633-695
    pixels: np.ndarray
    stride: int
    rect: Rectangle

    def bounds(self):
        return self.rect

    def pixOffset(self, x, y):
        return (y - self.rect.min.y) * self.stride + (x - self.rect.min.x)

    def at(self, x, y):
        if not Point(x, y).In(self.rect):
            return Color.fromGray(0)
        return Color.fromGray(int(self.pixels[self.pixOffset(x, y)]))


@dataclass
class RGBAImage:  # image.zig:133-227
    pixels: np.ndarray
    stride: int
    rect: Rectangle

    def bounds(self):
        return self.rect

    def pixOffset(self, x, y):
        return (y - self.rect.min.y) * self.stride + (x - self.rect.min.x) * 4

    def at(self, x, y):
        if not Point(x, y).In(self.rect):
            return Color.fromRGBA(0, 0, 0, 0)
        i = self.pixOffset(x, y)
        return Color.fromRGBA(*(int(v) for v in self.pixels[i:i + 4]))


@dataclass
class CMYKImage:  # image.zig:762-823
    pixels: np.ndarray
    stride: int
    rect: Rectangle

    def bounds(self):
        return self.rect

    def pixOffset(self, x, y):
        return (y - self.rect.min.y) * self.stride + (x - self.rect.min.x) * 4

    def at(self, x, y):
        if not Point(x, y).In(self.rect):
            return Color.fromCMYK(0, 0, 0, 0)
        i = self.pixOffset(x, y)
        return Color.fromCMYK(*(int(v) for v in self.pixels[i:i + 4]))


@dataclass
class YCbCrImage:  # image.zig:474-631
    y: np.ndarray
    cb: np.ndarray
    cr: np.ndarray
    y_stride: int
    c_stride: int
    subsample_ratio: YCbCrSubsample
    rect: Rectangle
    pixels: np.ndarray

    def bounds(self):
        return self.rect

    def yOffset(self, x, y):
        return (y - self.rect.min.y) * self.y_stride + (x - self.rect.min.x)

    def cOffset(self, x, y):  # image.zig:594-605
        s, r, m = self.c_stride, self.subsample_ratio, self.rect.min
        if r == YCbCrSubsample.Ratio422:
            return (y - m.y) * s + (x // 2 - m.x // 2)
        if r == YCbCrSubsample.Ratio420:
            return (y // 2 - m.y // 2) * s + (x // 2 - m.x // 2)
        if r == YCbCrSubsample.Ratio440:
            return (y // 2 - m.y // 2) * s + (x - m.x)
        if r == YCbCrSubsample.Ratio411:
            return (y - m.y) * s + (x // 4 - m.x // 4)
        if r == YCbCrSubsample.Ratio410:
            return (y // 2 - m.y // 2) * s + (x // 4 - m.x // 4)
        return (y - m.y) * s + (x - m.x)

    def YCbCrAt(self, x, y):
        if not Point(x, y).In(self.rect):
            return Color.fromYCbCr(0, 0, 0)
        yi, ci = self.yOffset(x, y), self.cOffset(x, y)
        return Color.fromYCbCr(int(self.y[yi]), int(self.cb[ci]), int(self.cr[ci]))

    at = YCbCrAt


class Image:
    """`image.Image` tagged union (image.zig:24-131). `tag` in {'Gray','YCbCr','RGBA','CMYK'}."""

    def __init__(self, tag: str, payload, device_rgba: Optional[np.ndarray] = None):
        self.tag = tag
        self.payload = payload
        self._device_rgba = device_rgba

    # union accessors, e.g. img.RGBA like `switch (img) { .RGBA => |m| ... }`
    def __getattr__(self, name):
        if name in ("Gray", "YCbCr", "RGBA", "CMYK"):
            return self.payload if self.tag == name else None
        raise AttributeError(name)

    def bounds(self) -> Rectangle:
        return self.payload.bounds()

    def at(self, x, y) -> Color:
        return self.payload.at(x, y)

    def free(self, allocator=None):  # image.zig:68-99; numpy owns the memory
        self.payload.pixels = None
        self._device_rgba = None

    def rgbaPixels(self, allocator=None) -> np.ndarray:
        """image.zig:103-130: tight W*H*4 RGBA bytes.  For images from the batch path these are the
        bytes the fused GPU kernel produced; otherwise at()/toRGBA() per pixel as the reference does."""
        if self._device_rgba is not None:
            return self._device_rgba
        r = self.bounds()
        out = np.empty((r.dY(), r.dX(), 4), np.uint8)
        for y in range(r.min.y, r.max.y):
            for x in range(r.min.x, r.max.x):
                c = self.at(x, y).toRGBA()
                out[y - r.min.y, x - r.min.x] = [c[0] >> 8, c[1] >> 8, c[2] >> 8, c[3] >> 8]
        return out.reshape(-1)
