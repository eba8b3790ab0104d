// zpx_k1.cu -- Huffman / run-length coefficient decode on the GPU.
//
// Replaces the MCU loop of processSos (src/jpeg/decoder.zig:1294-1452) together with
// decodeHuffman (:909-970), ensureNBits (:975-991), readByteStuffedByte (:712-749),
// receiveExtend (:1115-1134) and decodeBits (:1009-1022) for sequential (SOF0/SOF1) scans.
// Output: int16 coefficient blocks, natural (de-zigzagged) order, absolute DC, in the HBM layout
// k2 consumes (zpx_k2.cu header).
//
// Semantics kept from the reference, bit for bit on every conforming stream:
//   * MSB-first bit reader over the byte-stuffed stream (FF 00 -> FF); a 0xFF followed by anything
//     else is never consumed.  The host hands each restart interval its byte range [start, limit):
//     limit = first such 0xFF (zpx_parse.cpp).  Bits past the limit read as zero here and any
//     symbol that needs them is the reference's MissingFF00 (or UnexpectedEof at end of file).
//   * Huffman codes up to 16 bits, canonical; RECEIVE/EXTEND; DC prediction per component, reset at
//     each restart interval; AC run-length placement b[unzig[zig]].
//   * the End-Of-Band-run quirk of sequential scans inside one interval.
// Error kinds are reported per image through an atomicMin on (scan, block ordinal, code) so the
// first error in the reference's decode order wins.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "zpx_entropy.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

namespace zpx {

// ---------------------------------------------------------------------------
// K1a: one lane per restart interval (32 intervals per warp), serial inside the interval.
//
// The loop body decodes ONE Huffman symbol per lane per iteration, DC or AC alike, from a 32-bit
// table entry that already holds code length, value bits, zig-zag advance and total bits
// (ZpxHuffDev::fast), so the 32 lanes of a warp -- which sit at different symbols of different
// blocks -- share one short instruction stream; the loop head is a warp vote that keeps them in
// lock step.  Divergent: the end of a block (flush of the 128-byte block, next block's descriptor) and
// the rare paths (codes longer than ZPX_LUT_BITS, 0xFF inside a refill word, EOB runs, errors).
// The stream words are loaded two refills ahead so that their latency overlaps decoding.
// ---------------------------------------------------------------------------
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
__device__ __forceinline__ void sts_u16(uint32_t addr, int v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((short)v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stuffed-stream reader.  `nw` always holds the next word of the stream, loaded when the previous one
// was fed (a predicated load straight into nw's register: nothing waits for it until the next refill,
// several symbols later).  feed() is straight-line code for the common case (32 data bits, no 0xFF,
// inside the range); edges, FF 00 pairs and the end of the data take feed_slow().
struct FastReader {
    const uint32_t* base;  // 4-byte aligned base of the interval
    uint32_t off;          // byte offset (from base) of nw
    uint32_t first, end;   // valid byte range
    uint32_t nw;           // base[off / 4]
    uint64_t buf;          // unread bits, left aligned
    int cnt;               // bits in buf (data + zero padding)
    int pad;               // zero padding bits appended after the data ran out (always at the tail)
    uint32_t skip;         // next byte is the 0x00 of an FF 00 pair

    __device__ __forceinline__ void init(const uint8_t* blob, uint64_t start, uint32_t len) {
        const uint64_t a = start & ~(uint64_t)3;
        base = reinterpret_cast<const uint32_t*>(blob + a);
        first = (uint32_t)(start - a);
        end = first + len;
        off = 0;
        nw = __ldg(base);
        buf = 0;
        cnt = 0;
        pad = 0;
        skip = 0;
        feed_slow();  // the first word may start before `first`
    }
    // start at raw bit position `bitpos`, counted from (start & ~3) like the positions of zpx_k1s.cu; it lies in a
    // data byte (never in the 0x00 of an FF 00 pair) or at / past the limit (everything reads as padding)
    __device__ __forceinline__ void init_bits(const uint8_t* blob, uint64_t start, uint32_t len, uint32_t bitpos) {
        const uint64_t a = start & ~(uint64_t)3;
        const uint32_t lim = (uint32_t)(start - a) + len;
        const uint32_t byte = min(bitpos >> 3, lim);
        const uint32_t wofs = byte & ~3u;
        base = reinterpret_cast<const uint32_t*>(blob + a + wofs);
        first = byte - wofs;
        end = lim - wofs;
        off = 0;
        nw = __ldg(base);
        buf = 0;
        cnt = 0;
        pad = 0;
        skip = 0;
        feed_slow();
        const int b = (bitpos >> 3) < lim ? (int)(bitpos & 7u) : 0;
        buf <<= b;
        cnt -= b;
    }
    __device__ __forceinline__ bool overrun() const { return cnt < pad; }

    // make sure more than 32 bits are buffered.  One divergent region, entered every fifth symbol or so; the
    // common refill (32 data bits, no 0xFF, inside the range) is straight-line code that ends by loading the
    // NEXT word straight into nw's register (inline asm: no move waits for it), so nothing stalls on that load
    // until the next refill.
    __device__ __forceinline__ void feed() {
        if (cnt <= 32) {
            const uint32_t raw = nw;
            const uint32_t hasff = (~raw - 0x01010101u) & raw & 0x80808080u;
            if ((hasff | skip) == 0 && off + 4 <= end) {  // off >= first after init
                buf |= ((uint64_t)__byte_perm(raw, 0, 0x0123) << 32) >> cnt;
                cnt += 32;
                off += 4;
                // L1 allocates 32-byte sectors: ask for the sector four ahead when entering a new one
                if ((off & 31u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(base) + off + 128));
                asm volatile("ld.global.nc.u32 %0, [%1];" : "+r"(nw) : "l"(base + (off >> 2)));
            } else {
                feed_slow();
            }
        }
    }

    // byte by byte: range edges, FF 00 pairs, zero padding past the limit; until cnt > 32
    __device__ __forceinline__ void feed_slow() {
        while (cnt <= 32) {
            const uint32_t o = off;
            if (o >= end) {
                cnt += 32;
                pad += 32;
                continue;
            }
            const uint32_t be = __byte_perm(nw, 0, 0x0123);
            uint32_t acc = 0;
            int nb = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t oo = o + j;
                const uint32_t b = (be >> (24 - 8 * j)) & 0xffu;
                if (oo < first || oo >= end) continue;
                if (skip) {
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
            }
            off += 4;
            nw = __ldg(base + (off >> 2));  // at most one word past the limit: the blob is padded
        }
    }
};

// rare path of the symbol step: code longer than the first-level table, or a special entry.
// Returns the fast-entry fields for the symbol in the low word (ZPX_FE; bit 31 kept for an AC End-Of-Band
// run, with its r in byte 2) and an error code in the high word.
__device__ __noinline__ unsigned long long k1_slow_symbol(const ZpxHuffDev* __restrict__ tab, uint32_t hi, bool isdc,
                                                          uint32_t e) {
    int err = 0;
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
    // same field packing as zpx_fast_entry (zpx_parse.cpp)
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
    e = ZPX_FE((uint32_t)len + size, len, special ? rr : size, adv, special);
    return (unsigned long long)e | ((unsigned long long)(uint32_t)err << 32);
}

// ---------------------------------------------------------------------------
// The kernel.  Per CTA (NT lanes = NT consecutive restart intervals, usually 2-3 images):
//   setup   the distinct scans and Huffman tables of the CTA's intervals are collected; the tables'
//           first-level LUTs (K1_SLB bits) and the scans' per-block descriptors are staged in shared
//           memory, so the per-symbol lookup is one LDS and a block end touches no global descriptor.
//           CTAs with more than K1_MAXT tables / K1_MAXS scans fall back to the global LUTs.
//   loop    K1_T symbol steps per warp vote; a lane that finishes a block idles until the vote, then
//           the block-end code (flush + next block) runs once for all lanes that finished.
// ---------------------------------------------------------------------------
constexpr int K1_SLB = 9;     // bits of the shared-memory first-level LUT
constexpr int K1_MAXT = 12;   // Huffman tables cached per CTA (static shared memory stays under 48 KB)
constexpr int K1_MAXS = 8;    // scans cached per CTA

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

// fields of a ZpxHuffDev::fast entry (ZPX_FE): one PRMT / SHF each
__device__ __forceinline__ int fe_tot(uint32_t e) { return (int)__byte_perm(e, 0, 0x4440); }
__device__ __forceinline__ int fe_len(uint32_t e) { return (int)__byte_perm(e, 0, 0x4441); }
__device__ __forceinline__ int fe_size(uint32_t e) { return (int)__byte_perm(e, 0, 0x4442); }
__device__ __forceinline__ int fe_adv(uint32_t e) { return (int)(e >> 24); }  // special bit is clear on this path
// RECEIVE + EXTEND (decoder.zig:1115-1134) on the `size` bits at the top of t (size 0 -> 0): PTX shr clamps a
// shift by 32 to zero
__device__ __forceinline__ int fe_extend(uint32_t t, int size) {
    uint32_t v;
    asm("shr.u32 %0, %1, %2;" : "=r"(v) : "r"(t), "r"(32 - size));
    return (int)v + (((int)t >= 0) ? 1 - (1 << size) : 0);
}

// slow-path wrapper shared by the DC and AC steps: returns the entry fields, sets err / eob_run
template <bool SMEM>
__device__ __forceinline__ uint32_t k1_rare(const K1Params& P, const FastReader& br, uint32_t hi, bool isdc, uint32_t e,
                                            uint32_t lut_addr, const uint32_t* gtab, uint32_t slut, uint32_t& eob_run, int& err) {
    const ZpxHuffDev* __restrict__ tab;
    if (SMEM) {
        const uint32_t slot = (lut_addr - slut) >> (K1_SLB + 2);
        tab = &P.huff[lds_u32(slut + (K1_MAXT << (K1_SLB + 2)) + slot * 4)];
    } else {
        tab = reinterpret_cast<const ZpxHuffDev*>(gtab);
    }
    const uint32_t e10 = SMEM ? __ldg(&tab->fast[hi >> (32 - ZPX_LUT_BITS)]) : e;
    const unsigned long long r = k1_slow_symbol(tab, hi, isdc, e10);
    e = (uint32_t)r;
    err = (int)(r >> 32);
    if (e >> 31) {
        // (r, 0) with 0 < r < 15 (decoder.zig:1399-1407): eob_run = (1 << r | next r bits) - 1
        const int len = (int)((e >> 8) & 0xffu), rr = (int)((e >> 16) & 0xffu);
        eob_run = (1u << rr) | (uint32_t)((br.buf << len) >> (64 - rr));
        eob_run = (eob_run - 1) & 0xffffu;
        e = ZPX_FE(len + rr, len, 0, 64, 0);  // consume code + run bits, end of block
    }
    // an error ends the block: advance past index 63 (the caller's err stays set for the block-end code)
    if (err) e = (e & 0x00ffffffu) | 64u << 24;
    return e;
}

// Block-synchronous main loop: the 32 lanes of a warp decode their k-th block together --
//   DC symbol (all lanes), AC symbols until every lane's block is complete (a warp vote per symbol; a
//   lane whose block ended idles), block end (all lanes: flush the 128-byte block, next descriptor).
// Lanes of a warp sit at the same block phase of the MCU (all on luma or all on chroma), so the number
// of AC iterations is close to the lanes' own symbol counts, and the block-end code runs once per
// block for all lanes instead of once per symbol for a few.
// Where a lane starts inside its interval.  One lane per interval: the interval's first bit, block 0, all of its
// blocks.  Self-synchronising mode (SUB, zpx_k1s.cu): the true decoder state at the start of the lane's
// sub-sequence -- possibly inside a block, whose tail the lane skips because it belongs to the previous lane --
// and the number of blocks that start inside the sub-sequence, both known from the synchronisation passes.
struct K1Start {
    uint32_t bitpos;  // raw bit position, counted from (iv.start & ~3)
    int k;            // 0: at a block start; 1..63: inside a block, next zig-zag index k
    int dc0, dc1, dc2, dc3;  // DC predictors at that point
    uint32_t j0;      // ordinal (inside the interval) of the first block this lane writes
    uint32_t count;   // blocks this lane writes
    // SUB: after its last block the lane also decodes the next DC symbol, writing nothing, and reports the error
    // if there is one.  The synchronisation pass steps over invalid codes bit by bit; when that happens at a block
    // start, the block it finally finds may start in a later sub-sequence, and no lane would ever decode the
    // invalid code itself.  The lane that owns the preceding block does, here (reference: BadHuffmanCode at the
    // block that follows, decoder.zig:947-969).
    bool probe;
};

template <int NT, bool SMEM, bool SUB>
__device__ __forceinline__ void k1_lane_loop(const K1Params& P, const ZpxIntervalDev& iv, const K1Start& st, const uint32_t sb,
                                             const uint32_t su, const uint32_t sdesc /* smem: this lane's scan's blk table */,
                                             const uint32_t slut /* smem: LUT slots base */) {
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];

    FastReader br;
    if (st.count != 0 || (SUB && st.probe)) {
        br.init_bits(P.blob, iv.start, iv.len, st.bitpos);
    } else {  // idle lane: never reads
        br.base = reinterpret_cast<const uint32_t*>(P.blob);
        br.off = br.first = br.end = br.nw = br.skip = 0;
        br.buf = 0;
        br.cnt = br.pad = 0;
    }

    const bool interleaved = sc->interleaved != 0;
    const int nblk = interleaved ? sc->nblk : 1;
    const uint32_t mxx = (uint32_t)im->mxx;
    const bool planar = im->layout == ZPX_LAYOUT_PLANAR || !interleaved;
    const uint4* __restrict__ bpack = reinterpret_cast<const uint4*>(sc->blk_pack);
    const uint32_t cw = (uint32_t)sc->cw;
    const uint64_t coef_base = im->coef_base;
    const uint32_t bpm = (uint32_t)im->bpm;

    // position of block j0
    uint32_t mcu = iv.first_mcu, mx = 0, my = 0, bxn = 0, byn = 0;
    int c = 0;
    if (interleaved) {
        if (SUB) {
            mcu += st.j0 / (uint32_t)nblk;
            c = (int)(st.j0 % (uint32_t)nblk);
        }
        mx = mcu % mxx;
        my = mcu / mxx;
    } else {  // n-th coded block of the component, row-major over the blocks that intersect the image
        const uint32_t o = iv.first_block + (SUB ? st.j0 : 0u);
        byn = o / cw;
        bxn = o - byn * cw;
    }
    // SUB: a lane that starts inside a block first runs that block's remaining AC symbols without storing
    // anything (phase: the one before block j0's)
    bool tail = SUB && st.k != 0 && (st.count != 0 || st.probe);
    const int c_first = c;
    if (tail) c = c == 0 ? nblk - 1 : c - 1;
    // bi: x = DC table (SMEM: shared address of its LUT slot; else table index), y = AC likewise,
    //     z = comp | hx << 8 | vy << 16 | slot << 24, w = h | v << 8 | undefined-table flags
    uint4 bi = SMEM ? lds_u128(sdesc + c * 16) : bpack[c];
    const uint32_t* __restrict__ gdc = SMEM ? nullptr : P.huff[bi.x].fast;
    const uint32_t* __restrict__ gac = SMEM ? nullptr : P.huff[bi.y].fast;

    int dc0 = st.dc0, dc1 = st.dc1, dc2 = st.dc2, dc3 = st.dc3;
    uint32_t eob_run = 0;
    int wide = 0;  // max over the lane's symbols of (AC value bits, DC magnitude >> 8): >= 13 / >= 16 flags the image
    const bool probe = SUB && st.probe;
    const uint32_t total = st.count + (probe ? 1u : 0u);
    uint32_t left = total;  // blocks still to decode (including the current one and the probe)

    while (__any_sync(0xffffffffu, left != 0)) {
        int k = 64;  // > 63: no block in flight on this lane
        int err = 0;
        if (SUB && tail) {
            k = st.k;
        } else if (left != 0) {
            // ---- DC (decoder.zig:1366-1376) ----
            br.feed();
            const uint32_t hi = (uint32_t)(br.buf >> 32);
            uint32_t e;
            if (SMEM) e = lds_u32(bi.x + ((hi >> (32 - K1_SLB)) << 2));
            else e = __ldg(gdc + (hi >> (32 - ZPX_LUT_BITS)));
            if ((int)e <= 0) e = k1_rare<SMEM>(P, br, hi, true, e, bi.x, gdc, slut, eob_run, err);
            if (bi.w & 0x10000u) err = ZPX_E_UninitializedHuffmanTable;
            const int len = fe_len(e), size = fe_size(e);
            const int v = fe_extend(__funnelshift_l((uint32_t)br.buf, (uint32_t)(br.buf >> 32), len), size);
            const int comp = (int)(bi.z & 0xff);
            int dc = comp == 0 ? dc0 : comp == 1 ? dc1 : comp == 2 ? dc2 : dc3;
            dc += v;
            if (comp == 0) dc0 = dc; else if (comp == 1) dc1 = dc; else if (comp == 2) dc2 = dc; else dc3 = dc;
            if (dc < -32768 || dc > 32767) report_coef_range(P.status, im->status_slot);
            wide = max(wide, ((dc ^ (dc >> 31)) >> 12) ? 13 : 0);
            br.buf <<= fe_tot(e);
            br.cnt -= fe_tot(e);
            sts_u16(sb, dc);
            k = 1;
            if (eob_run > 0) {  // decoder.zig:1379-1380 (End-Of-Band run)
                eob_run--;
                k = 64;
            }
            if (err) k = 64;
            else if (k == 1 && (bi.w & 0x20000u)) {
                err = ZPX_E_UninitializedHuffmanTable;
                k = 64;
            }
            if (SUB && probe && left == 1) {
                // probe block: only the DC symbol counts, and only if it is an error (a valid one belongs to the
                // next lane's block); nothing is stored
                k = 64;
                sts_u16(sb, 0);
                if (err == ZPX_E_UninitializedHuffmanTable && !(bi.w & 0x10000u)) err = 0;
                if (!err && !br.overrun()) left = 0;
            }
        }
        // ---- AC (decoder.zig:1383-1411): one symbol per lane per vote ----
        while (__any_sync(0xffffffffu, k <= 63)) {
            if (k <= 63) {
                br.feed();
                const uint32_t hi = (uint32_t)(br.buf >> 32);
                uint32_t e;
                if (SMEM) e = lds_u32(bi.y + ((hi >> (32 - K1_SLB)) << 2));
                else e = __ldg(gac + (hi >> (32 - ZPX_LUT_BITS)));
                if ((int)e <= 0) {
                    e = k1_rare<SMEM>(P, br, hi, false, e, bi.y, gac, slut, eob_run, err);
                    // End-Of-Band RUN inside a sequential scan: the synchronisation passes do not
                    // model that state
                    if (SUB && eob_run != 0) {
                        eob_run = 0;
                        if (!err) err = ZPX_E_UNSUPPORTED_STREAM;
                    }
                }
                const int len = fe_len(e), size = fe_size(e);
                int tot = fe_tot(e);
                const int adv = fe_adv(e);  // 64 after an error: the block ends here
                wide = max(wide, size);
                const int v = fe_extend(__funnelshift_l((uint32_t)br.buf, (uint32_t)(br.buf >> 32), len), size);
                const int kk = k + adv - 1;                    // zig-zag index the value goes to
                bool store = size != 0 && !(SUB && tail);
                if (kk > 63) {  // decoder.zig:1393-1395: run past the block end, the value bits stay unread
                    if (size != 0) tot = len;  // (an End-Of-Band run keeps its r run bits: size == 0 there)
                    store = false;
                }
                k += adv;
                br.buf <<= tot;
                br.cnt -= tot;
                if (store) sts_u16(sb + lds_u16(su + 2 * kk), v);
            }
        }
        // ---- block end ----
        if (SUB && tail) {
            // end of the skipped tail (its errors are the previous lane's to report): block j0 comes next
            tail = false;
            c = c_first;
            if (SMEM) {
                bi = lds_u128(sdesc + c * 16);
            } else {
                bi = bpack[c];
                gdc = P.huff[bi.x].fast;
                gac = P.huff[bi.y].fast;
            }
        } else if (left != 0) {
            if (err || br.overrun()) {
                // a symbol that needed bits past the limit is the reference's MissingFF00 / UnexpectedEof,
                // whatever the garbage decoded to
                if (br.overrun()) err = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;
                report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + st.j0 + (total - left), err);
                left = 0;
                eob_run = 0;
            } else {
                // hand the block to HBM: slot s of the 128-byte line holds row s ^ key
                uint32_t bx;
                uint64_t blk;
                if (interleaved) {
                    bx = (bi.w & 0xff) * mx + ((bi.z >> 8) & 0xff);
                    if (planar) {
                        const int comp = (int)(bi.z & 0xff);
                        const uint32_t by = ((bi.w >> 8) & 0xff) * my + ((bi.z >> 16) & 0xff);
                        blk = im->comp_base[comp] + (uint64_t)by * im->comp_bw[comp] + bx;
                    } else {
                        blk = coef_base + (uint64_t)mcu * bpm + (bi.z >> 24);
                    }
                } else {
                    bx = bxn;
                    const int comp = (int)(bi.z & 0xff);
                    blk = im->comp_base[comp] + (uint64_t)byn * im->comp_bw[comp] + bx;
                }
                uint4* __restrict__ dst = P.coef + blk * 8;
                const uint32_t key = bx & 7;
                const bool keep = !(bi.w & 0x40000u);  // not superseded by a later scan of the same component
#pragma unroll
                for (uint32_t slot = 0; slot < 8; slot++) {
                    const uint32_t a = sb + (slot ^ key) * (NT * 16);
                    if (keep) dst[slot] = lds_v4(a);
                    sts_zero16(a);
                }
                left--;
                if (interleaved) {
                    if (++c == nblk) {
                        c = 0;
                        mcu++;
                        if (++mx == mxx) { mx = 0; my++; }
                    }
                    if (SMEM) {
                        bi = lds_u128(sdesc + c * 16);
                    } else {
                        bi = bpack[c];
                        gdc = P.huff[bi.x].fast;
                        gac = P.huff[bi.y].fast;
                    }
                } else if (++bxn == cw) {
                    bxn = 0;
                    byn++;
                }
            }
        }
    }
    // The reference keeps its End-Of-Band run across scans (decoder.zig:144, reset only at RSTn :1451); here
    // every scan starts from zero, so a run that is still open when a scan ends (corrupt streams only) would
    // make the next scan differ: refuse the image instead.
    if (wide >= 13) atomicOr(&P.img_flags[im->status_slot], 1u);  // a coefficient outside [-4096, 4095]
    if (!SUB && st.count != 0 && (iv.flags & 2u) && eob_run != 0)
        report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + iv.n_blocks, ZPX_E_UNSUPPORTED_STREAM);
}

// One CTA: NT lanes, each with an interval and a start inside it (K1Start).  Collects the distinct scans and
// Huffman tables of the CTA's lanes, stages their first-level LUTs and block descriptors in shared memory and
// runs the block-synchronous loop.
template <int NT, bool SUB>
__device__ __forceinline__ void k1_cta_run(const K1Params& P, const ZpxIntervalDev& iv, const K1Start& st) {
    __shared__ uint4 sblk[8 * NT];   // per-lane block: [8 rows][NT lanes] x 16 bytes
    // zig-zag index -> byte offset of that coefficient inside the lane's block (row * NT*16 + column * 2);
    // lane-divergent index: shared, not constant, memory (padded: k + run <= 78)
    __shared__ uint16_t s_unzig[80];
    __shared__ __align__(16) uint32_t s_lut[(K1_MAXT << K1_SLB) + K1_MAXT];  // LUT slots, then the slots' table indices
    __shared__ uint4 s_desc[K1_MAXS][ZPX_MAX_BLK_PER_MCU];
    __shared__ uint32_t s_scan[K1_MAXS];
    __shared__ uint32_t s_lane_scan[NT];
    __shared__ int s_nscan, s_ntab, s_ok;

    const int tid = threadIdx.x;
    for (int r = 0; r < 8; r++) sblk[r * NT + tid] = make_uint4(0, 0, 0, 0);
    if (tid < 80) {
        const int nat = tid < 64 ? c_unzig[tid] : 63;
        s_unzig[tid] = (uint16_t)((nat >> 3) * (NT * 16) + (nat & 7) * 2);
    }
    s_lane_scan[tid] = iv.scan;
    __syncthreads();

    // ---- collect the CTA's scans and tables (one thread; a few dozen steps) ----
    if (tid == 0) {
        int ns = 0, nt = 0, ok = 1;
        uint32_t* tab_ids = s_lut + (K1_MAXT << K1_SLB);
        for (int l = 0; l < NT && ok; l++) {
            const uint32_t scn = s_lane_scan[l];
            if (l > 0 && scn == s_lane_scan[l - 1]) continue;
            int f = -1;
            for (int i = 0; i < ns; i++)
                if (s_scan[i] == scn) f = i;
            if (f >= 0) continue;
            if (ns == K1_MAXS) { ok = 0; break; }
            const ZpxScanDev* sc = &P.scans[scn];
            const int nb = sc->interleaved ? sc->nblk : 1;
            for (int b = 0; b < nb && ok; b++) {
                uint4 d = reinterpret_cast<const uint4*>(sc->blk_pack)[b];
                uint32_t ids[2] = {d.x, d.y};
                for (int j = 0; j < 2; j++) {
                    int slot = -1;
                    for (int i = 0; i < nt; i++)
                        if (tab_ids[i] == ids[j]) slot = i;
                    if (slot < 0) {
                        if (nt == K1_MAXT) { ok = 0; break; }
                        slot = nt++;
                        tab_ids[slot] = ids[j];
                    }
                    ids[j] = (uint32_t)__cvta_generic_to_shared(s_lut) + ((uint32_t)slot << (K1_SLB + 2));
                }
                d.x = ids[0];
                d.y = ids[1];
                s_desc[ns][b] = d;
            }
            s_scan[ns++] = scn;
        }
        s_nscan = ns;
        s_ntab = nt;
        s_ok = ok;
    }
    __syncthreads();
    const bool cached = s_ok != 0;
    uint32_t sdesc = 0;
    if (cached) {
        // stage the LUTs: entry i of the K1_SLB-bit table = entry 2^(10-SLB)*i of the 10-bit one if its code fits
        const int nt = s_ntab;
        const uint32_t* tab_ids = s_lut + (K1_MAXT << K1_SLB);
        for (int i = tid; i < (nt << K1_SLB); i += NT) {
            const int slot = i >> K1_SLB, ix = i & ((1 << K1_SLB) - 1);
            uint32_t e = __ldg(&P.huff[tab_ids[slot]].fast[ix << (ZPX_LUT_BITS - K1_SLB)]);
            if (((e >> 8) & 0xffu) > (uint32_t)K1_SLB) e = 0;
            s_lut[i] = e;
        }
        for (int i = 0; i < s_nscan; i++)
            if (s_scan[i] == iv.scan) sdesc = (uint32_t)__cvta_generic_to_shared(&s_desc[i][0]);
    }
    __syncthreads();
    uint32_t sb = smem_addr(sblk) + tid * 16;  // this lane's row 0
    uint32_t su = smem_addr(s_unzig);
    uint32_t slut = smem_addr(s_lut);
    // opaque copies: keeps the shared-window address arithmetic out of the symbol loop
    asm volatile("mov.u32 %0, %0;" : "+r"(sb));
    asm volatile("mov.u32 %0, %0;" : "+r"(su));
    asm volatile("mov.u32 %0, %0;" : "+r"(slut));
    if (cached) k1_lane_loop<NT, true, SUB>(P, iv, st, sb, su, sdesc, slut);
    else k1_lane_loop<NT, false, SUB>(P, iv, st, sb, su, 0, 0);
}

// K1a: one lane per restart interval.
// LPW = lanes of each warp that carry an interval.  The kernel is bound by the latency of a warp's
// serial step, not by issue slots; with LPW = 16 a warp has half the divergent work per step (fewer
// block ends, refills and cache misses to wait for) and twice as many warps fill the idle issue slots.
template <int NT, int LPW>
__global__ void __launch_bounds__(NT) k1_lane_per_interval(const K1Params P) {
    const int tid = threadIdx.x;
    const int gid = (blockIdx.x * (NT / 32) + (tid >> 5)) * LPW + (tid & 31);
    // lanes past the end of the interval list (or beyond LPW) idle through the loop (its head is a warp vote)
    const bool live = gid < P.n_iv && (tid & 31) < LPW;
    const ZpxIntervalDev iv = P.ivs[live ? gid : P.n_iv - 1];
    K1Start st;
    st.bitpos = ((uint32_t)iv.start & 3u) * 8u;
    st.k = 0;
    st.dc0 = st.dc1 = st.dc2 = st.dc3 = 0;
    st.j0 = 0;
    st.count = live ? iv.n_blocks : 0;
    st.probe = false;
    k1_cta_run<NT, false>(P, iv, st);
}

// K1b, last pass (zpx_k1s.cu): one lane per sub-sequence, 32 consecutive sub-sequences of one segment per warp.
// Every lane starts from its true state (found by k1s_sync), skips the tail of a block begun in the previous
// sub-sequence and writes the blocks that START inside its own (their number and the DC predictors at that
// point come from k1s_scan), running past its boundary to finish the last one.
template <int NT>
__global__ void __launch_bounds__(NT) k1s_write(const K1SParams P) {
    const int wid = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const bool wv = wid < P.n_warps;
    const ZpxWarpDev w = P.warps[wv ? wid : P.n_warps - 1];
    const ZpxIntervalDev iv = P.k1.ivs[w.iv];
    const uint32_t li = w.first + lane;
    const bool valid = wv && li < iv.nsub;
    const uint32_t t = iv.sub_first + (valid ? li : 0);
    const unsigned long long in = P.s_in[t];
    const uint32_t excl = min((uint32_t)max(P.s_n[t], 0), iv.n_blocks);
    const uint32_t next = (valid && li + 1 < iv.nsub) ? min((uint32_t)max(P.s_n[t + 1], 0), iv.n_blocks) : iv.n_blocks;
    const int4 dc = P.s_dc[t];
    K1Start st;
    st.bitpos = (uint32_t)in;
    st.k = (int)((in >> 40) & 0xff);
    st.dc0 = dc.x;
    st.dc1 = dc.y;
    st.dc2 = dc.z;
    st.dc3 = dc.w;
    st.j0 = excl;
    st.count = valid && next > excl ? next - excl : 0;
    // probe the symbol after the lane's last block if the lane skipped an invalid code and a block follows
    st.probe = valid && P.s_bad[t] != 0 && st.j0 + st.count < iv.n_blocks;
    k1_cta_run<NT, true>(P.k1, iv, st);
}

// The kernel keeps 44 KB of shared memory per CTA; at the register-limited 5 CTAs per SM the driver carves
// 228 KB out of the 256 KB unified array and leaves the 640 per-lane streams of an SM some 28 KB of L1.
// Measured on cfg2 (544 CTAs): carve-out 100 / 86 % 5.81 ms, 72 % (164 KB, 3 CTAs per SM, 92 KB of L1) 5.64 ms,
// 57 % 8.43 ms.  ZPX_K1_CARVEOUT overrides the percentage (-1: leave it to the driver).
template <typename K>
static void k1_prefer_l1(K kernel) {
    static bool done = false;
    if (!done) {
        const char* e = getenv("ZPX_K1_CARVEOUT");
        const int pct = e ? atoi(e) : 72;
        if (pct >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        done = true;
    }
}

cudaError_t k1_launch_lane_per_interval(const K1Params& P, cudaStream_t s) {
    if (P.n_iv <= 0) return cudaSuccess;
    constexpr int NT = 128;
    k1_prefer_l1(k1_lane_per_interval<NT, 32>);
    if (P.lanes_per_warp == 16) {
        constexpr int per_cta = (NT / 32) * 16;
        k1_lane_per_interval<NT, 16><<<(P.n_iv + per_cta - 1) / per_cta, NT, 0, s>>>(P);
    } else {
        k1_lane_per_interval<NT, 32><<<(P.n_iv + NT - 1) / NT, NT, 0, s>>>(P);
    }
    return cudaGetLastError();
}

cudaError_t k1s_launch_write(const K1SParams& P, cudaStream_t s) {
    if (P.n_warps <= 0) return cudaSuccess;
    constexpr int NT = 128;
    const int wpc = NT / 32;
    k1s_write<NT><<<(P.n_warps + wpc - 1) / wpc, NT, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace zpx
