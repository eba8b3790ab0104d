// zpx_idct.cuh -- exact fixed-point 8x8 IDCT of the reference (src/jpeg/idct.zig:77-201),
// one thread per block, everything in registers.  Every shift of the reference is a rounding point
// and stays where it is; only value-neutral rewrites are used:
//   * the row pass' "all AC zero" shortcut (idct.zig:84-97: every output = s0 << 3) is not taken: the
//     general formulas give the same values, ((s0<<11)+128)>>8 == s0<<3, as long as s0 << 11 does not
//     wrap, i.e. |s0| < 2^20 -- true for every conforming stream (|DC-column value| <= 2047 *
//     255).  Images that may break that bound (a coefficient outside [-4096, 4095] with 8-bit quantisers:
//     only garbage streams have them; the entropy kernels flag them, one max per symbol) take
//     idct_block_q8_exact, which does such rows the reference's way; the unfused path always does;
//   * dequantisation (decoder.zig:1564-1567) is fused into the row-pass loads; the <<11 prescale of
//     columns 0 and 4 is folded into the quantiser ((c*q)<<11 == c*(q<<11) in wrapping 32-bit);
//   * level shift + clamp (decoder.zig:1622-1628: v<-128 -> 0, v>127 -> 255, else v+128) is a
//     saturating pack to s8 followed by ^0x80 on the packed bytes.
// All arithmetic is wrapping 32-bit int, like the oracle.
#pragma once
#include <stdint.h>

namespace zpx {

constexpr int W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565;
constexpr int W1PW7 = W1 + W7, W1MW7 = W1 - W7, W2PW6 = W2 + W6, W2MW6 = W2 - W6, W3PW5 = W3 + W5, W3MW5 = W3 - W5;
constexpr int R2 = 181;

// Arithmetic right shift by K as a multiply-high: floor(x / 2^K) == (x * 2^(32-K)) >> 32 for every int32 x, so the
// value is the reference's `>> K` exactly; the instruction (IMAD.HI) goes to the FMA pipe instead of the ALU pipe,
// which the shifts, adds and packs of the IDCT keep busier (ZPX_SHIFT_VIA_MULHI selects, measured both ways).
#ifndef ZPX_SHIFT_VIA_MULHI
#define ZPX_SHIFT_VIA_MULHI 0
#endif
template <int K>
__device__ __forceinline__ int asr(int x) {
#if ZPX_SHIFT_VIA_MULHI
    return __mulhi(x, 1 << (32 - K));
#else
    return x >> K;
#endif
}

// signed halves of a packed pair of int16
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }

// Row pass on one row.  In: 8 dequantised values, x0/x1 already prescaled (x0 without the +128).
// Out: o[0..7].
__device__ __forceinline__ void idct_row(int x0, int x4, int x3, int x7, int x1, int x6, int x2, int x5, int* o) {
    // names follow idct.zig:100-107: x0=s0<<11+128, x1=s4<<11, x2=s6, x3=s2, x4=s1, x5=s7, x6=s5, x7=s3
    x0 += 128;
    int x8 = W7 * (x4 + x5);
    x4 = x8 + W1MW7 * x4;
    x5 = x8 - W1PW7 * x5;
    x8 = W3 * (x6 + x7);
    x6 = x8 - W3MW5 * x6;
    x7 = x8 - W3PW5 * x7;

    x8 = x0 + x1;
    x0 -= x1;
    x1 = W6 * (x3 + x2);
    x2 = x1 - W2PW6 * x2;
    x3 = x1 + W2MW6 * x3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;

    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (R2 * (x4 + x5) + 128) >> 8;
    x4 = (R2 * (x4 - x5) + 128) >> 8;

    o[0] = asr<8>(x7 + x1);
    o[1] = asr<8>(x3 + x2);
    o[2] = asr<8>(x0 + x4);
    o[3] = asr<8>(x8 + x6);
    o[4] = asr<8>(x8 - x6);
    o[5] = asr<8>(x0 - x4);
    o[6] = asr<8>(x3 - x2);
    o[7] = asr<8>(x7 - x1);
}

// Column pass on one column (stride 8 inside b), in place.  idct.zig:149-199.
__device__ __forceinline__ void idct_col(int* b) {
    int y0 = (b[8 * 0] << 8) + 8192;
    int y1 = b[8 * 4] << 8;
    int y2 = b[8 * 6], y3 = b[8 * 2], y4 = b[8 * 1], y5 = b[8 * 7], y6 = b[8 * 5], y7 = b[8 * 3];

    int y8 = W7 * (y4 + y5) + 4;
    y4 = (y8 + W1MW7 * y4) >> 3;
    y5 = (y8 - W1PW7 * y5) >> 3;
    y8 = W3 * (y6 + y7) + 4;
    y6 = (y8 - W3MW5 * y6) >> 3;
    y7 = (y8 - W3PW5 * y7) >> 3;

    y8 = y0 + y1;
    y0 -= y1;
    y1 = W6 * (y3 + y2) + 4;
    y2 = (y1 - W2PW6 * y2) >> 3;
    y3 = (y1 + W2MW6 * y3) >> 3;
    y1 = y4 + y6;
    y4 -= y6;
    y6 = y5 + y7;
    y5 -= y7;

    y7 = y8 + y3;
    y8 -= y3;
    y3 = y0 + y2;
    y0 -= y2;
    y2 = (R2 * (y4 + y5) + 128) >> 8;
    y4 = (R2 * (y4 - y5) + 128) >> 8;

    b[8 * 0] = asr<14>(y7 + y1);
    b[8 * 1] = asr<14>(y3 + y2);
    b[8 * 2] = asr<14>(y0 + y4);
    b[8 * 3] = asr<14>(y8 + y6);
    b[8 * 4] = asr<14>(y8 - y6);
    b[8 * 5] = asr<14>(y0 - y4);
    b[8 * 6] = asr<14>(y3 - y2);
    b[8 * 7] = asr<14>(y7 - y1);
}

// pack four int32 to bytes with signed saturation to [-128,127], then +128 (== ^0x80 on each byte)
__device__ __forceinline__ uint32_t pack4_level_shift(int v0, int v1, int v2, int v3) {
    uint32_t t, d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
    return d ^ 0x80808080u;
}

// pack four int32 to bytes with unsigned saturation to [0,255]
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3) {
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
    return d;
}

// signed 16-bit x unsigned 8-bit two-way dot product (SASS IDP.2A.LO/HI.S16.U8): with b = q_even |
// q_odd << 24, lo gives a.lo16 * q_even and hi gives a.hi16 * q_odd -- unpack + dequantise of a
// coefficient in one instruction (8-bit quantiser tables, decoder.zig:645-647).
__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0));
    return d;
}

// Same as dequant_idct_block below for quantisers that fit 8 bits: `qp` holds, per row, four words
// q[2j] | q[2j+1] << 24.  The <<11 prescale of columns 0 and 4 (idct.zig:100-101) is applied after
// the product ((c*q)<<11, wrapping).
// Exact for blocks whose first-column coefficients lie in [-4096, 4095] (see the header comment); images that
// have wider ones are flagged by the entropy kernels and take idct_block_q8_exact instead.
template <typename LoadRow>
__device__ __forceinline__ void dequant_idct_block_q8(LoadRow ld, const uint32_t* __restrict__ qp, uint32_t (&px)[16]) {
    int b[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 c = ld(r);
        const uint4 q = *reinterpret_cast<const uint4*>(qp + r * 4);
        const int s0 = dp2a_lo_su(c.x, q.x) << 11, s1 = dp2a_hi_su(c.x, q.x);
        const int s2 = dp2a_lo_su(c.y, q.y), s3 = dp2a_hi_su(c.y, q.y);
        const int s4 = dp2a_lo_su(c.z, q.z) << 11, s5 = dp2a_hi_su(c.z, q.z);
        const int s6 = dp2a_lo_su(c.w, q.w), s7 = dp2a_hi_su(c.w, q.w);
        idct_row(s0, s1, s2, s3, s4, s5, s6, s7, &b[r * 8]);
    }
#pragma unroll
    for (int x = 0; x < 8; x++) idct_col(&b[x]);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        px[2 * r + 0] = pack4_level_shift(b[r * 8 + 0], b[r * 8 + 1], b[r * 8 + 2], b[r * 8 + 3]);
        px[2 * r + 1] = pack4_level_shift(b[r * 8 + 4], b[r * 8 + 5], b[r * 8 + 6], b[r * 8 + 7]);
    }
}

// Same values as dequant_idct_block_q8 for a block whose coefficient rows NR..7 are all zero (and, with LOCOLS, whose
// columns 4..7 are zero too: nothing outside the top-left 4x4 corner when NR == 4): the zero inputs are constants, so
// the multiplications by them, the row pass of the zero rows (every output of an all-zero row is (0 + 128) >> 8 = 0)
// and the column terms they feed fold away at compile time; every surviving operation is the one the general code
// performs, on the same operands (wrapping 32-bit), hence bit-identical.  The caller takes a variant only when EVERY
// block of the warp qualifies (a warp vote): <4, true> for chroma blocks of photographs, <6, false> for luma blocks
// whose two highest-frequency rows quantise to zero.  (The reference has a data-dependent shortcut of its own at the
// same place: rows whose AC are all zero, idct.zig:84-97.)
template <int NR, bool LOCOLS, typename LoadRow>
__device__ __forceinline__ void dequant_idct_block_q8_sparse(LoadRow ld, const uint32_t* __restrict__ qp, uint32_t (&px)[16]) {
    int b[64];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const uint4 c = ld(r);
        if (LOCOLS) {
            const uint2 q = *reinterpret_cast<const uint2*>(qp + r * 4);
            const int s0 = dp2a_lo_su(c.x, q.x) << 11, s1 = dp2a_hi_su(c.x, q.x);
            const int s2 = dp2a_lo_su(c.y, q.y), s3 = dp2a_hi_su(c.y, q.y);
            idct_row(s0, s1, s2, s3, 0, 0, 0, 0, &b[r * 8]);
        } else {
            const uint4 q = *reinterpret_cast<const uint4*>(qp + r * 4);
            const int s0 = dp2a_lo_su(c.x, q.x) << 11, s1 = dp2a_hi_su(c.x, q.x);
            const int s2 = dp2a_lo_su(c.y, q.y), s3 = dp2a_hi_su(c.y, q.y);
            const int s4 = dp2a_lo_su(c.z, q.z) << 11, s5 = dp2a_hi_su(c.z, q.z);
            const int s6 = dp2a_lo_su(c.w, q.w), s7 = dp2a_hi_su(c.w, q.w);
            idct_row(s0, s1, s2, s3, s4, s5, s6, s7, &b[r * 8]);
        }
    }
#pragma unroll
    for (int k = NR * 8; k < 64; k++) b[k] = 0;
#pragma unroll
    for (int x = 0; x < 8; x++) idct_col(&b[x]);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        px[2 * r + 0] = pack4_level_shift(b[r * 8 + 0], b[r * 8 + 1], b[r * 8 + 2], b[r * 8 + 3]);
        px[2 * r + 1] = pack4_level_shift(b[r * 8 + 4], b[r * 8 + 5], b[r * 8 + 6], b[r * 8 + 7]);
    }
}

// The same block the reference's way where it matters: a row whose dequantised AC are all zero yields s0 << 3
// in every column (idct.zig:84-97), which differs from the general row once s0 << 11 wraps (|s0| >= 2^20; only
// garbage streams).  Out of line and self-contained (reads the block from shared memory again, stores the 8x8
// pixels itself), so that it costs the hot path nothing but the test that calls it.
__device__ __noinline__ void idct_block_q8_exact(const uint4* blk, int key, const uint32_t* __restrict__ qp, uint8_t* dst, int pitch) {
    int b[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 c = blk[r ^ key];
        const uint4 q = *reinterpret_cast<const uint4*>(qp + r * 4);
        const int u0 = dp2a_lo_su(c.x, q.x), s1 = dp2a_hi_su(c.x, q.x);
        const int s2 = dp2a_lo_su(c.y, q.y), s3 = dp2a_hi_su(c.y, q.y);
        const int u4 = dp2a_lo_su(c.z, q.z), s5 = dp2a_hi_su(c.z, q.z);
        const int s6 = dp2a_lo_su(c.w, q.w), s7 = dp2a_hi_su(c.w, q.w);
        if ((s1 | s2 | s3 | u4 | s5 | s6 | s7) == 0) {
            const int dc = (int)((uint32_t)u0 << 3);
#pragma unroll
            for (int k = 0; k < 8; k++) b[r * 8 + k] = dc;
        } else {
            idct_row((int)((uint32_t)u0 << 11), s1, s2, s3, (int)((uint32_t)u4 << 11), s5, s6, s7, &b[r * 8]);
        }
    }
#pragma unroll
    for (int x = 0; x < 8; x++) idct_col(&b[x]);
#pragma unroll
    for (int r = 0; r < 8; r++)
        *reinterpret_cast<uint2*>(dst + r * pitch) =
            make_uint2(pack4_level_shift(b[r * 8 + 0], b[r * 8 + 1], b[r * 8 + 2], b[r * 8 + 3]),
                       pack4_level_shift(b[r * 8 + 4], b[r * 8 + 5], b[r * 8 + 6], b[r * 8 + 7]));
}

// Dequantise + IDCT one block.  `ld(r)` returns row r of the block as a uint4 of 8 int16 (natural
// order); `q` points at the block's quantiser in natural order with columns 0 and 4 pre-multiplied
// by 2048 (int32[64], warp-uniform address).  Result: 8 rows x 2 words of level-shifted pixels.
template <typename LoadRow>
__device__ __forceinline__ void dequant_idct_block(LoadRow ld, const int* __restrict__ q, uint32_t (&px)[16]) {
    int b[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 c = ld(r);
        const int4 qa = *reinterpret_cast<const int4*>(q + r * 8);
        const int4 qb = *reinterpret_cast<const int4*>(q + r * 8 + 4);
        const int s0 = lo16(c.x) * qa.x, s1 = hi16(c.x) * qa.y, s2 = lo16(c.y) * qa.z, s3 = hi16(c.y) * qa.w;
        const int s4 = lo16(c.z) * qb.x, s5 = hi16(c.z) * qb.y, s6 = lo16(c.w) * qb.z, s7 = hi16(c.w) * qb.w;
        // idct_row(x0=s0', x4=s1, x3=s2, x7=s3, x1=s4', x6=s5, x2=s6, x5=s7)
        idct_row(s0, s1, s2, s3, s4, s5, s6, s7, &b[r * 8]);
        // (this unfused path is not the hot one: the reference's all-AC-zero row, idct.zig:84-97, is taken literally;
        // qa.x carries the << 11 prescale of column 0, so the plain product is s0 >> 11 exactly when it did not wrap,
        // and is recomputed from the coefficient otherwise)
        if ((s1 | s2 | s3 | (lo16(c.z) * (qb.x >> 11)) | s5 | s6 | s7) == 0) {
            const int dc = (int)((uint32_t)(lo16(c.x) * (qa.x >> 11)) << 3);
#pragma unroll
            for (int k = 0; k < 8; k++) b[r * 8 + k] = dc;
        }
    }
#pragma unroll
    for (int x = 0; x < 8; x++) idct_col(&b[x]);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        px[2 * r + 0] = pack4_level_shift(b[r * 8 + 0], b[r * 8 + 1], b[r * 8 + 2], b[r * 8 + 3]);
        px[2 * r + 1] = pack4_level_shift(b[r * 8 + 4], b[r * 8 + 5], b[r * 8 + 6], b[r * 8 + 7]);
    }
}

// ---- colour (src/color/color.zig:90-126 followed by the >>8 of image.zig:122-125) ----
// YCbCr -> RGBA8: channel = sat_u8(v >> 16) reproduces the reference's branch exactly for every
// int32 v: (v & 0xff000000)==0 -> v>>16 ; v<0 -> 0 ; else 255.
__device__ __forceinline__ uint32_t ycc_pixel(int y, int rr, int gg, int bb) {
    const int r = y * 0x10101 + rr, g = y * 0x10101 + gg, b = y * 0x10101 + bb;
    return pack4_sat_u8(r >> 16, g >> 16, b >> 16, 255);
}
// The same pixel with the ">> 16" of the first NW channels done by the multiplier: (y << 8) * (0x10101 << 8) + (t << 16)
// as a 64-bit multiply-add is (y * 0x10101 + t) * 2^16 exactly (no 32-bit intermediate, |y * 0x10101 + t| < 2^31), whose
// high word is floor((y * 0x10101 + t) / 2^16) == the reference's arithmetic shift.  One IMAD.WIDE (FMA pipe) instead
// of IMAD + SHF (FMA + ALU pipe).  ZPX_YCC_WIDE selects how many channels take it (measured).
#ifndef ZPX_YCC_WIDE
#define ZPX_YCC_WIDE 0
#endif
__device__ __forceinline__ int madw_hi(int a, int b, long long c) {
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return (int)(d >> 32);
}
struct ChromaTerms {
    int rr, gg, bb;
    long long rr64, gg64, bb64;  // the terms << 16 (only the ones ZPX_YCC_WIDE uses are ever computed)
};
template <int NW>
__device__ __forceinline__ uint32_t ycc_pixel_w(int y, const ChromaTerms& t) {
    const int y8 = y << 8;
    const int r = NW >= 1 ? madw_hi(y8, 0x10101 << 8, t.rr64) : (y * 0x10101 + t.rr) >> 16;
    const int b = NW >= 2 ? madw_hi(y8, 0x10101 << 8, t.bb64) : (y * 0x10101 + t.bb) >> 16;
    const int g = NW >= 3 ? madw_hi(y8, 0x10101 << 8, t.gg64) : (y * 0x10101 + t.gg) >> 16;
    return pack4_sat_u8(r, g, b, 255);
}
// per-chroma-sample terms, shared by every luma pixel that replicates the sample
// (cb - 128, cr - 128 of color.zig:92-93 multiplied out: the same integers, one multiply-add per term)
__device__ __forceinline__ void chroma_terms(int cb, int cr, int& rr, int& gg, int& bb) {
    rr = 91881 * cr - 91881 * 128;
    gg = -22554 * cb + (-46802 * cr + (22554 + 46802) * 128);
    bb = 116130 * cb - 116130 * 128;
}
__device__ __forceinline__ void chroma_terms_w(int cb, int cr, ChromaTerms& t) {
    chroma_terms(cb, cr, t.rr, t.gg, t.bb);
    t.rr64 = (long long)t.rr << 16;
    t.gg64 = (long long)t.gg << 16;
    t.bb64 = (long long)t.bb << 16;
}
// CMYK -> RGBA8 (color.zig:115-121, then >>8).  c,m,y,k are the stored (already inverted) bytes.
__device__ __forceinline__ uint32_t cmyk_pixel(uint32_t c, uint32_t m, uint32_t y, uint32_t k) {
    const uint32_t w = 0xffffu - k * 0x101u;
    const uint32_t r = ((0xffffu - c * 0x101u) * w / 0xffffu) >> 8;
    const uint32_t g = ((0xffffu - m * 0x101u) * w / 0xffffu) >> 8;
    const uint32_t b = ((0xffffu - y * 0x101u) * w / 0xffffu) >> 8;
    return r | (g << 8) | (b << 16) | 0xff000000u;
}

}  // namespace zpx
