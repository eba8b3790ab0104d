"""Host-side mirror of zpix's `color` module (reference src/color/color.zig).

Only the interface the JPEG path touches: the `Color` variants the decoder's images
yield and `toRGBA()` (color.zig:31-132).  Plain Python integers, same arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass


def _clamp16(v: int) -> int:
    # color.zig:100: (v & 0xff000000) == 0 ? v >> 8 : ~(v >> 31) & 0xffff   (v is i32)
    if (v & 0xFFFFFFFF) & 0xFF000000 == 0:
        return v >> 8
    return 0 if v < 0 else 0xFFFF


@dataclass(frozen=True)
class Color:
    """Tagged union `color.Color` (color.zig:13-23); kind in {'gray','ycbcr','cmyk','rgba'}."""

    kind: str
    v: tuple

    @staticmethod
    def fromGray(y):
        return Color("gray", (y,))

    @staticmethod
    def fromYCbCr(y, cb, cr):
        return Color("ycbcr", (y, cb, cr))

    @staticmethod
    def fromCMYK(c, m, y, k):
        return Color("cmyk", (c, m, y, k))

    @staticmethod
    def fromRGBA(r, g, b, a):
        return Color("rgba", (r, g, b, a))

    def toRGBA(self):
        """Alpha-premultiplied 16-bit (r, g, b, a), color.zig:31-132."""
        if self.kind == "gray":
            y = self.v[0] | (self.v[0] << 8)
            return (y, y, y, 0xFFFF)
        if self.kind == "rgba":
            return tuple(c | (c << 8) for c in self.v)
        if self.kind == "ycbcr":
            y, cb, cr = self.v
            yy1, cb1, cr1 = y * 0x10101, cb - 128, cr - 128
            return (
                _clamp16(yy1 + 91881 * cr1),
                _clamp16(yy1 - 22554 * cb1 - 46802 * cr1),
                _clamp16(yy1 + 116130 * cb1),
                0xFFFF,
            )
        if self.kind == "cmyk":
            c, m, y, k = self.v
            w = 0xFFFF - k * 0x101
            return tuple((0xFFFF - ch * 0x101) * w // 0xFFFF for ch in (c, m, y)) + (0xFFFF,)
        raise ValueError(self.kind)
