// zpx_entropy.cuh -- device code shared by the entropy kernels (zpx_k1.cu, zpx_k1s.cu, zpx_k3.cu):
// zig-zag table, canonical Huffman symbol decode from the tables in HBM, RECEIVE/EXTEND, error report.
// Reference semantics: src/jpeg/decoder.zig:909-1022 (decodeHuffman, ensureNBits, decodeBit(s)),
// :1115-1134 (receiveExtend).
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
