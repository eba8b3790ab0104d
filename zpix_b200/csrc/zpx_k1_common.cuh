// zpx_k1_common.cuh -- what the sequential entropy kernels share (zpx_k1.cu: write passes, zpx_k1s.cu:
// synchronisation passes of the self-synchronising decoder):
//   * RingReader: a bit position over the UNSTUFFED stream of one restart interval (k0_unstuff, zpx_k0.cu, has
//     removed the FF 00 stuffing, so src/jpeg/decoder.zig:712-749 readByteStuffedByte costs nothing here) read
//     through a small per-lane ring of big-endian words in shared memory.  No bit buffer, no refill branch in the
//     symbol loop: a symbol step is two LDS for the 32-bit window, one for the table entry and straight-line
//     arithmetic, identical on every lane.  The ring is topped up at warp-uniform points (block starts, every
//     K1_TOPUP symbols) with 16-byte global loads issued one chunk ahead.
//   * K1Tables: per CTA, the first-level Huffman LUTs (DC 7 bits, AC 9 bits; 32-bit ZPX_FE entries), the canonical
//     limit / offset / value arrays of the CTA's tables (codes longer than the first level never leave shared
//     memory) and the per-block descriptors of its scans.
//   * the rare-symbol path (long codes, End-Of-Band runs, DC category > 16, invalid codes).
// Reference semantics: decoder.zig:909-970 decodeHuffman, :975-1022 ensureNBits/decodeBits, :1115-1134 receiveExtend.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "zpx_entropy.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

#ifndef ZPX_WINDOW_LOP3
#define ZPX_WINDOW_LOP3 2
#endif

namespace zpx {

constexpr int K1_MAXSLOTS = 256;  // lanes with a stream ("slots") per CTA, at most
constexpr int K1_RW = 16;        // ring words per lane (64 bytes)
constexpr int K1_TOPUP = 8;      // AC symbols between two top-ups: after one, >= 4*K1_RW - 15 = 49 bytes lie ahead; a DC
                                 // symbol and 8 AC symbols take at most 9 * 32 bits = 36 bytes, and the three-word
                                 // window of the write kernels reaches at most 11 bytes further
constexpr int K1_DLB = 7;        // first-level bits of DC tables in shared memory
constexpr int K1_ALB = 9;        //                     AC tables
constexpr int K1_MAXT = 12;      // tables cached per CTA
constexpr int K1_MAXS = 8;       // scans cached per CTA
constexpr int K1_LUTW = 6 * (1 << K1_DLB) + 6 * (1 << K1_ALB);  // LUT pool, 32-bit words (15 KB)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, int v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((short)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0) : "memory");
}

// fields of a ZPX_FE entry: one PRMT / SHF each
__device__ __forceinline__ int fe_tot(uint32_t e) { return (int)__byte_perm(e, 0, 0x4440); }
__device__ __forceinline__ int fe_len(uint32_t e) { return (int)__byte_perm(e, 0, 0x4441); }
__device__ __forceinline__ int fe_s32(uint32_t e) { return (int)__byte_perm(e, 0, 0x4442); }  // 32 - value bits
__device__ __forceinline__ int fe_adv(uint32_t e) { return (int)(e >> 24); }  // special bit is clear where this is used
// RECEIVE + EXTEND (decoder.zig:1115-1134) on the value bits at the top of t, s32 = 32 - their number (32: no value
// bits -> 0): PTX shr clamps a shift by 32 to zero.  A value whose first bit is 0 is negative: v - (2^size - 1).
__device__ __forceinline__ int fe_extend(uint32_t t, int s32) {
    uint32_t v, m;
    asm("shr.u32 %0, %1, %2;" : "=r"(v) : "r"(t), "r"(s32));
    asm("shr.u32 %0, %1, %2;" : "=r"(m) : "r"(0xffffffffu), "r"(s32));
    return (int)t >= 0 ? (int)(v - m) : (int)v;
}

// ---------------------------------------------------------------------------
// reader
// ---------------------------------------------------------------------------
template <int RS, int RW = K1_RW>  // RS: bytes between consecutive ring words of one lane = 4 * (lanes with a stream
struct RingReader {                //     per CTA); RW: ring words per lane
    uint32_t ring;       // shared address of this lane's column: word slot s at ring + s * RS
    const uint8_t* src;  // first byte of the interval in the unstuffed blob (16-byte aligned)
    uint32_t bitpos;     // next unread bit, from src
    uint32_t endbits;    // 8 * unstuffed length: a symbol that ends beyond it needed bits the stream does not have
    uint32_t wbyte;      // bytes handed to the ring so far (multiple of 16)
    uint32_t lim16;      // chunks at or past this offset read as zeros
    uint4 nxt;           // the chunk at wbyte, loaded ahead

    __device__ __forceinline__ uint4 fetch(uint32_t off) const {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (off < lim16) v = __ldg(reinterpret_cast<const uint4*>(src + off));
        return v;
    }
    // chunks can be added as long as the one that holds the read position stays in the ring
    __device__ __forceinline__ void topup() {
        while (wbyte - ((bitpos >> 7) << 4) <= (uint32_t)(4 * RW - 16)) {
            const uint32_t a = ring + ((wbyte >> 2) & (uint32_t)(RW - 1)) * (uint32_t)RS;
            sts_u32(a, __byte_perm(nxt.x, 0, 0x0123));
            sts_u32(a + RS, __byte_perm(nxt.y, 0, 0x0123));
            sts_u32(a + 2 * RS, __byte_perm(nxt.z, 0, 0x0123));
            sts_u32(a + 3 * RS, __byte_perm(nxt.w, 0, 0x0123));
            wbyte += 16;
            nxt = fetch(wbyte);
        }
    }
    __device__ __forceinline__ void init(uint32_t ring_col, const uint8_t* ublob, uint64_t ustart, uint32_t ulen, uint32_t pos) {
        ring = ring_col;
        src = ublob + ustart;
        endbits = ulen * 8u;
        lim16 = (ulen + 15u) & ~15u;
        bitpos = pos;
        wbyte = (pos >> 7) << 4;
        nxt = fetch(wbyte);
        topup();
    }
    // a lane without work: never loads, never tops up
    __device__ __forceinline__ void init_idle(uint32_t ring_col) {
        ring = ring_col;
        src = nullptr;
        endbits = lim16 = bitpos = 0;
        wbyte = 4 * RW;
        nxt = make_uint4(0u, 0u, 0u, 0u);
    }
    // shared address of ring word (bitpos >> 5) + d
    __device__ __forceinline__ uint32_t word_addr(uint32_t d) const {
        return ring + (((bitpos >> 5) + d) & (uint32_t)(RW - 1)) * (uint32_t)RS;
    }
    // the next 32 bits
    __device__ __forceinline__ uint32_t peek() const {
        return __funnelshift_l(lds_u32(word_addr(1)), lds_u32(word_addr(0)), bitpos);
    }
    __device__ __forceinline__ bool overrun() const { return bitpos > endbits; }
};

// The stream window of the write kernels (zpx_k1.cu, zpx_k3l.cu): three ring words in registers -- the word under the
// bit position and the two after it -- so the 32 bits of a step come from one funnel shift; when a step crosses a
// word boundary the registers move up and the third is reloaded from the ring, a load nothing waits for until the
// step after the next.
template <int RS>
struct Window {
    uint32_t w0, w1, w2;
    uint32_t a2;  // shared address of w2's ring word
    __device__ __forceinline__ void load(const RingReader<RS>& rd) {
        w0 = lds_u32(rd.word_addr(0));
        w1 = lds_u32(rd.word_addr(1));
        a2 = rd.word_addr(2);
        w2 = lds_u32(a2);
    }
    __device__ __forceinline__ uint32_t peek(uint32_t bitpos) const { return __funnelshift_l(w1, w0, bitpos); }
    // the position moves from `from` by tot <= 32 bits.  Branch-free: m is all ones when it enters the next word
    // (then the registers move up and w2 is the next ring word; otherwise w2 is simply read again)
    __device__ __forceinline__ void advance(const RingReader<RS>& rd, uint32_t from, uint32_t tot) {
        const uint32_t t = (from & 31u) + tot;
#if ZPX_WINDOW_LOP3 == 2
        // predicate form: compare + selects (one dependent operation less than shift, shift, LOP3)
        const bool cross = t >= 32u;
        w0 = cross ? w1 : w0;
        w1 = cross ? w2 : w1;
        a2 += cross ? (uint32_t)RS : 0u;
        if (a2 == rd.ring + K1_RW * RS) a2 = rd.ring;
        w2 = lds_u32(a2);
        return;
#endif
        const uint32_t m = (uint32_t)((int)(t << 26) >> 31);
#if ZPX_WINDOW_LOP3
        // one three-input logic instruction per word: (next & m) | (this & ~m)  (the compiler builds it from a compare, a
        // select and a two-input LOP3, all three on the position's dependency chain)
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(w0) : "r"(w0), "r"(w1), "r"(m));
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(w1) : "r"(w1), "r"(w2), "r"(m));
#else
        w0 = (w1 & m) | (w0 & ~m);
        w1 = (w2 & m) | (w1 & ~m);
#endif
        a2 += m & (uint32_t)RS;
        if (a2 == rd.ring + K1_RW * RS) a2 = rd.ring;
        w2 = lds_u32(a2);
    }
};

// ---------------------------------------------------------------------------
// per-CTA table cache
// ---------------------------------------------------------------------------
struct K1Tables {
    uint32_t lut[K1_LUTW];
    uint32_t lim[K1_MAXT][16];    // limit[l] at [l - 1]
    int32_t valoff[K1_MAXT][16];  // valoff[l] at [l - 1]
    uint8_t vals[K1_MAXT][256];
    // per scan, per block of its MCU:
    //   x = shared address of the DC table's LUT, y = of the AC table's,
    //   z = comp | hx << 8 | vy << 16 | slot << 24,
    //   w = h | v << 8 | (DC table undefined) << 16 | (AC table undefined) << 17 | superseded << 18 |
    //       DC cache slot << 20 | AC cache slot << 24
    uint4 desc[K1_MAXS][ZPX_MAX_BLK_PER_MCU];
    uint32_t scan_id[K1_MAXS];
    uint32_t tab_id[K1_MAXT];   // device table index of each cache slot
    uint32_t tab_lut[K1_MAXT];  // word offset of the slot's LUT inside lut[]
    uint32_t tab_bits[K1_MAXT]; // first-level bits of the slot: K1_DLB (a DC table) or K1_ALB
    uint32_t lane_scan[K1_MAXSLOTS];
    // zig-zag index -> byte offset of that coefficient inside a lane's block in the write kernels
    // (row * 16 + column * 2); padded: k + run <= 78
    uint16_t unzig[80];
    int nscan, ntab, ok;
};

// Collect the distinct scans and Huffman tables of the CTA's `slots` streams (lane_scan[] filled by the caller,
// followed by __syncthreads) and stage them.  Returns false when they do not fit: the CTA then reads the tables in
// HBM.  Ends with __syncthreads.
__device__ __forceinline__ bool k1_tables_setup(const K1Params& P, K1Tables& T, const int slots) {
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid == 0) {
        int ns = 0, nt = 0, ok = 1;
        uint32_t words = 0;
        for (int l = 0; l < slots && ok; l++) {
            const uint32_t scn = T.lane_scan[l];
            if (l > 0 && scn == T.lane_scan[l - 1]) continue;
            int f = -1;
            for (int i = 0; i < ns; i++)
                if (T.scan_id[i] == scn) f = i;
            if (f >= 0) continue;
            if (ns == K1_MAXS) { ok = 0; break; }
            const ZpxScanDev* sc = &P.scans[scn];
            const int nb = sc->interleaved ? sc->nblk : 1;
            for (int b = 0; b < nb && ok; b++) {
                uint4 d = reinterpret_cast<const uint4*>(sc->blk_pack)[b];
                uint32_t ids[2] = {d.x, d.y};
                uint32_t slots[2] = {0, 0};
                for (int j = 0; j < 2; j++) {
                    int slot = -1;
                    for (int i = 0; i < nt; i++)
                        if (T.tab_id[i] == ids[j]) slot = i;
                    if (slot < 0) {
                        const uint32_t need = 1u << (j == 0 ? K1_DLB : K1_ALB);
                        if (nt == K1_MAXT || words + need > (uint32_t)K1_LUTW) { ok = 0; break; }
                        slot = nt++;
                        T.tab_id[slot] = ids[j];
                        T.tab_lut[slot] = words;
                        T.tab_bits[slot] = j == 0 ? K1_DLB : K1_ALB;
                        words += need;
                    }
                    slots[j] = (uint32_t)slot;
                    ids[j] = smem_addr(T.lut) + T.tab_lut[slot] * 4u;
                }
                d.x = ids[0];
                d.y = ids[1];
                d.w = (d.w & 0x000fffffu) | slots[0] << 20 | slots[1] << 24;
                T.desc[ns][b] = d;
            }
            T.scan_id[ns++] = scn;
        }
        T.nscan = ns;
        T.ntab = nt;
        T.ok = ok;
    }
    if (tid < 80) {
        const int nat = tid < 64 ? c_unzig[tid] : 63;
        T.unzig[tid] = (uint16_t)((nat >> 3) * 16 + (nat & 7) * 2);
    }
    __syncthreads();
    const bool cached = T.ok != 0;
    if (cached) {
        const int nt = T.ntab;
        for (int s = 0; s < nt; s++) {
            const uint32_t w0 = T.tab_lut[s];
            const int bits = (int)T.tab_bits[s];
            const ZpxHuffDev* __restrict__ tab = &P.huff[T.tab_id[s]];
            for (int i = tid; i < (1 << bits); i += nthr) {
                uint32_t e = __ldg(&tab->fast[i << (ZPX_LUT_BITS - bits)]);
                if (((e >> 8) & 0xffu) > (uint32_t)bits) e = 0;
                T.lut[w0 + i] = e;
            }
            if (tid < 16) {
                T.lim[s][tid] = __ldg(&tab->limit[tid + 1]);
                T.valoff[s][tid] = __ldg(&tab->valoff[tid + 1]);
            }
            for (int i = tid; i < 64; i += nthr)
                reinterpret_cast<uint32_t*>(T.vals[s])[i] = __ldg(reinterpret_cast<const uint32_t*>(tab->vals) + i);
        }
    }
    __syncthreads();
    return cached;
}

// ---------------------------------------------------------------------------
// rare path of a symbol step: code longer than the first-level table, or a special entry (AC End-Of-Band
// run, DC category > 16).  Returns the ZPX_FE fields of the symbol in the low word (bit 31 kept for an End-Of-Band
// run, with its r in byte 2) and an error code in the high word.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long k1_pack_symbol(uint32_t sym, int len, bool isdc) {
    // same field packing as zpx_fast_entry (zpx_parse.cpp)
    int err = 0;
    uint32_t size, adv, special = 0, rr = 0;
    if (isdc) {
        size = sym;
        adv = 1;
        if (sym > 16) {  // DC category > 16 (decoder.zig:1370)
            size = 0;
            err = ZPX_E_ExcessiveDCComponent;
        }
    } else {
        const uint32_t r = sym >> 4, s2 = sym & 15;
        if (s2 != 0) { size = s2; adv = r + 1; }
        else if (r == 15) { size = 0; adv = 16; }
        else if (r == 0) { size = 0; adv = 64; }
        else { size = 0; adv = 64; special = 1; rr = r; }
    }
    const uint32_t e = special ? ZPX_FE_RUN((uint32_t)len, len, rr) : ZPX_FE((uint32_t)len + size, len, size, adv, 0);
    return (unsigned long long)e | ((unsigned long long)(uint32_t)err << 32);
}

// tables in HBM (CTAs whose tables did not fit the cache); e = the 10-bit first-level entry
static __device__ __noinline__ unsigned long long k1_slow_symbol(const ZpxHuffDev* __restrict__ tab, uint32_t hi, bool isdc, uint32_t e) {
    uint32_t sym = 0;
    int len = (int)((e >> 8) & 0xffu);
    if (e == 0) {
        const uint32_t v16 = hi >> 16;
        len = 0;
        for (int l = ZPX_LUT_BITS + 1; l <= 16 && len == 0; l++) {
            if (v16 < tab->limit[l]) {
                sym = tab->vals[(tab->valoff[l] + (int)(v16 >> (16 - l))) & 0xff];
                len = l;
            }
        }
        if (len == 0)  // the reference reads 16 bits, then BadHuffmanCode (decoder.zig:947-969)
            return (unsigned long long)ZPX_FE(16, 16, 0, 64, 0) | ((unsigned long long)ZPX_E_BadHuffmanCode << 32);
    } else {
        // special first-level entry: recover the symbol from the 16-bit table
        sym = (uint32_t)tab->lut[hi >> (32 - ZPX_LUT_BITS)] >> 8;
    }
    return k1_pack_symbol(sym, len, isdc);
}

// tables in the CTA's cache: slot `slot` of the K1Tables at shared address tb; e = the first-level entry
// (0: longer than the first level or no code; else special).  The canonical search keeps the reference's
// first-match rule (decoder.zig:946-969: the shortest length l with code <= max_codes[l]).
static __device__ __noinline__ unsigned long long k1_slow_symbol_sm(uint32_t tb, uint32_t slot, uint32_t hi, bool isdc, uint32_t e) {
    const uint32_t v16 = hi >> 16;
    int len = (int)((e >> 8) & 0xffu);
    if (e == 0) {
        const uint32_t la = tb + (uint32_t)offsetof(K1Tables, lim) + slot * 64u;
        const int lb = isdc ? K1_DLB : K1_ALB;
        len = 0;
#pragma unroll
        for (int l = 16; l > K1_DLB; l--)
            if (l > lb && v16 < lds_u32(la + (uint32_t)(l - 1) * 4u)) len = l;
        if (len == 0)
            return (unsigned long long)ZPX_FE(16, 16, 0, 64, 0) | ((unsigned long long)ZPX_E_BadHuffmanCode << 32);
    }
    const int off = (int)lds_u32(tb + (uint32_t)offsetof(K1Tables, valoff) + slot * 64u + (uint32_t)(len - 1) * 4u);
    const uint32_t sym = lds_u8(tb + (uint32_t)offsetof(K1Tables, vals) + slot * 256u + (uint32_t)((off + (int)(v16 >> (16 - len))) & 0xff));
    return k1_pack_symbol(sym, len, isdc);
}

// The common rare symbol, inline: an ordinary AC run/size symbol whose code is longer than the first-level table
// (tables in the CTA's cache).  Canonical search over the cached limits -- first match = the reference's rule,
// decoder.zig:946-969.  Returns its ZPX_FE entry, or 0 for everything else (End-Of-Band runs, values of 13 bits
// or more, invalid codes): those take k1_slow_symbol_sm.
__device__ __forceinline__ uint32_t k1_long_ac_sm(uint32_t tb, uint32_t slot, uint32_t hi) {
    static_assert(K1_ALB == 9, "the search below starts at length 10");
    const uint32_t v16 = hi >> 16;
    const uint32_t la = tb + (uint32_t)offsetof(K1Tables, lim) + slot * 64u;
    const uint4 la9 = lds_u128(la + 32), la13 = lds_u128(la + 48);  // limit[9..12], limit[13..16]
    int len = 0;
    if (v16 < la13.w) len = 16;
    if (v16 < la13.z) len = 15;
    if (v16 < la13.y) len = 14;
    if (v16 < la13.x) len = 13;
    if (v16 < la9.w) len = 12;
    if (v16 < la9.z) len = 11;
    if (v16 < la9.y) len = 10;
    if (len == 0) return 0;
    const int off = (int)lds_u32(tb + (uint32_t)offsetof(K1Tables, valoff) + slot * 64u + (uint32_t)(len - 1) * 4u);
    const uint32_t sym = lds_u8(tb + (uint32_t)offsetof(K1Tables, vals) + slot * 256u + (uint32_t)((off + (int)(v16 >> (16 - len))) & 0xff));
    const uint32_t r = sym >> 4, s2 = sym & 15u;
    if (s2 != 0 && s2 < 13) return ZPX_FE((uint32_t)len + s2, len, s2, r + 1, 0);
    if (sym == 0xf0u) return ZPX_FE(len, len, 0, 16, 0);
    if (sym == 0) return ZPX_FE(len, len, 0, 64, 0);
    return 0;
}

}  // namespace zpx
