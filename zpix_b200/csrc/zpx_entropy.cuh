// zpx_entropy.cuh -- device code shared by the entropy kernels (zpx_k1.cu, zpx_k1s.cu, zpx_k3.cu):
// stuffed-stream bit reader with position tracking, Huffman symbol decode, RECEIVE/EXTEND, error report.
// Reference semantics: src/jpeg/decoder.zig:712-749 (readByteStuffedByte), :909-1022
// (decodeHuffman, ensureNBits, decodeBit(s)), :1115-1134 (receiveExtend).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_internal.h"

namespace zpx {

__constant__ uint8_t c_unzig[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
};

// ---------------------------------------------------------------------------
// bit reader over [start, start+len) of the stuffed stream
// ---------------------------------------------------------------------------
struct BitReader {
    const uint32_t* words;  // 4-byte aligned base covering the range
    uint32_t widx;          // next word to feed
    uint32_t nextw;         // words[widx], loaded one refill ahead so its latency overlaps decoding
    uint32_t first;         // byte offset (relative to words) of the first valid byte
    uint32_t end;           // byte offset (relative to words) one past the last valid byte
    uint64_t buf;           // unread bits, left aligned
    int cnt;                // number of bits in buf (real + zero padding)
    uint32_t fed;           // real data bits fed into buf so far
    uint32_t pad;           // zero padding bits fed after the data ran out
    uint32_t skip;          // the next byte is the 0x00 of an FF 00 pair
    // sub-sequence boundary tracking (self-synchronising decoder): B = data bits fed before the
    // feed position reached byte offset bnd (a multiple of 4)
    uint32_t bnd;
    uint32_t B;
    uint32_t bpassed;

    __device__ __forceinline__ void init(const uint8_t* blob, uint64_t start, uint32_t len) {
        const uint64_t a = start & ~(uint64_t)3;
        words = reinterpret_cast<const uint32_t*>(blob + a);
        first = (uint32_t)(start - a);
        end = first + len;
        widx = 0;
        buf = 0;
        cnt = 0;
        fed = 0;
        pad = 0;
        skip = 0;
        bnd = 0xffffffffu;
        B = 0;
        bpassed = 0;
        nextw = __ldg(words);
        fill();
    }

    // start at raw bit position `bitpos` (relative to `w`), which must lie in a data byte;
    // bytes before it are ignored.  end_off: byte offset of the limit.  boundary: see bnd.
    __device__ __forceinline__ void init_at(const uint32_t* w, uint32_t bitpos, uint32_t end_off, uint32_t boundary) {
        words = w;
        first = bitpos >> 3;
        end = end_off;
        widx = first >> 2;
        buf = 0;
        cnt = 0;
        fed = 0;
        pad = 0;
        skip = 0;
        bnd = boundary;
        B = 0;
        bpassed = 0;
        nextw = __ldg(words + widx);
        fill();
        consume((int)(bitpos & 7));
    }

    // raw bit position (relative to words) of the next unread bit.  Walks back over the buffered
    // data bytes; an FF 00 pair counts as one data byte.  Deterministic for any input; exact
    // whenever the reader started on a data byte of a well-formed stream.
    __device__ __forceinline__ uint32_t rawpos() const {
        const uint8_t* raw = reinterpret_cast<const uint8_t*>(words);
        uint32_t e = widx * 4;
        if (e > end) e = end;
        const uint32_t u = used();
        if (u > fed) return end * 8 + (u - fed);          // overran the data: past-the-end marker
        const uint32_t real = fed - u;                    // data bits fed but not yet consumed
        if (real == 0) return (e + skip) * 8;
        const uint32_t nbytes = (real + 7) >> 3;
        const uint32_t headbits = real - 8 * (nbytes - 1);  // unread bits of the head byte, 1..8
        uint32_t r = e;
        for (uint32_t k = 0; k < nbytes; k++) {
            r -= 1;
            if (r > first && raw[r] == 0x00 && raw[r - 1] == 0xff) r -= 1;
        }
        return r * 8 + (8 - headbits);
    }

    // bits consumed so far
    __device__ __forceinline__ uint32_t used() const { return fed + pad - (uint32_t)cnt; }
    // true if a symbol needed bits the stream does not have
    __device__ __forceinline__ bool overrun() const { return used() > fed; }

    // append up to one word, byte by byte (range edges, FF 00 pairs, zeros past the limit)
    __device__ __forceinline__ void fill_once() {
        const uint32_t off = widx * 4;
        if (off >= bnd && !bpassed) {
            bpassed = 1;
            B = fed;
        }
        if (off >= end) {  // past the limit: zeros
            cnt += 32;
            pad += 32;
            return;
        }
        const uint32_t raw = nextw;
        widx++;
        nextw = __ldg(words + widx);  // at most one word past the limit: the blob is padded
        const uint32_t be = __byte_perm(raw, 0, 0x0123);
        uint32_t acc = 0;
        int nb = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t o = off + j;
            const uint32_t b = (be >> (24 - 8 * j)) & 0xffu;
            if (o < first || o >= end) continue;
            if (skip) {  // the stuffed 0x00
                skip = 0;
                continue;
            }
            acc = (acc << 8) | b;
            nb++;
            if (b == 0xffu) skip = 1;
        }
        if (nb) {
            const uint32_t w = acc << (32 - 8 * nb);
            buf |= ((uint64_t)w << 32) >> cnt;
            cnt += 8 * nb;
            fed += 8 * nb;
        }
    }
    // make sure more than 32 bits are buffered.  Common case (a whole word of data without 0xFF) is
    // straight-line code; the next word is loaded by a predicated load straight into nextw's register,
    // so nothing waits for it until the next refill.
    __device__ __forceinline__ void fill() {
        const bool need = cnt <= 32;
        const uint32_t off = widx * 4;
        const uint32_t raw = nextw;
        const uint32_t nff = ~raw;
        const uint32_t hasff = (nff - 0x01010101u) & raw & 0x80808080u;
        const bool fast = need && (hasff | skip) == 0 && off >= first && off + 4 <= end && (bpassed || off < bnd);
        if (fast) {
            buf |= ((uint64_t)__byte_perm(raw, 0, 0x0123) << 32) >> cnt;
            cnt += 32;
            fed += 32;
            widx++;
        }
        {
            const uint32_t* p = words + widx;
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.global.nc.u32 %0, [%1];\n}\n"
                : "+r"(nextw)
                : "l"(p), "r"((uint32_t)fast));
        }
        if (need && !fast)
            while (cnt <= 32) fill_once();
    }
    __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void consume(int n) {
        buf <<= n;
        cnt -= n;
    }
};

struct HuffSym {
    uint32_t sym;
    int len;  // code length; 0 = no code matches (BadHuffmanCode after 16 bits)
};

// decodeHuffman (decoder.zig:909-970) with a ZPX_LUT_BITS-bit first level
__device__ __forceinline__ HuffSym huff_decode(const ZpxHuffDev* __restrict__ t, uint32_t hi) {
    HuffSym r;
    const uint32_t e = __ldg(&t->lut[hi >> (32 - ZPX_LUT_BITS)]);
    r.len = (int)(e & 0xffu);
    r.sym = e >> 8;
    if (r.len == 0) {
        const uint32_t v16 = hi >> 16;
#pragma unroll 1
        for (int l = ZPX_LUT_BITS + 1; l <= 16; l++) {
            if (v16 < __ldg(&t->limit[l])) {
                r.sym = __ldg(&t->vals[(__ldg(&t->valoff[l]) + (int)(v16 >> (16 - l))) & 0xff]);
                r.len = l;
                break;
            }
        }
    }
    return r;
}

// RECEIVE + EXTEND (decoder.zig:1115-1134) on the `size` bits that follow a `len`-bit code
__device__ __forceinline__ int receive_extend(uint64_t buf, int len, int size) {
    const uint32_t t = (uint32_t)((buf << len) >> 32);   // value bits, left aligned
    const int v = (int)((t >> 1) >> (31 - size));        // size == 0 -> 0
    const int neg = (size != 0) && !(t >> 31);           // first bit 0 -> negative
    return neg ? v + ((-1) << size) + 1 : v;
}

__device__ __forceinline__ void report(unsigned long long* status, uint32_t img_slot, int scan_index, uint64_t ordinal,
                                       int code) {
    const unsigned long long key =
        ((unsigned long long)(uint32_t)scan_index << 48) | ((ordinal & 0xffffffffffull) << 8) | (unsigned)code;
    atomicMin(&status[img_slot], key);
}

// A coefficient left the int16 range the HBM layout stores (non-conforming stream; the reference keeps int32).
// The image cannot be reproduced, but decoding goes on with the wrapped value: if the reference meets a real
// error further on (MissingFF00, BadHuffmanCode ...), that one is what it returns, so this code sorts after
// every real error key.
__device__ __forceinline__ void report_coef_range(unsigned long long* status, uint32_t img_slot) {
    atomicMin(&status[img_slot], (0x7fffull << 48) | (unsigned)ZPX_E_COEF_RANGE);
}

}  // namespace zpx
