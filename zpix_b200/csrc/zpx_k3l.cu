// zpx_k3l.cu -- progressive (SOF2) scans, one LANE per restart interval (per scan when DRI = 0).
//
// Replaces the progressive branches of processSos (src/jpeg/decoder.zig:1268-1283, 1340-1425), refine (:1459-1518)
// and refineNonZeroes (:1522-1549) for every frame whose scan script is an ordinary successive approximation (each
// band of each component coded once, then refined with strictly falling Al -- zpx_api.cu checks it; anything else
// takes the warp-per-interval kernel of zpx_k3.cu, which reads the coefficients back).
//
// A scan is a serial stream: what a block consumes depends on everything before it, and in an AC refinement pass
// also on which coefficients earlier scans left non-zero.  zpx_k3.cu spends a whole warp on one such stream (lane 0
// decodes, 31 lanes feed it) and is bound by instruction issue at one useful lane in 32.  Here every lane owns a
// stream of its own -- the 32 lanes of a warp run the same pass of 32 different images -- and the serial part of a
// pass is cut down to what really is serial:
//   * stream: the unstuffed copy k0_unstuff made (FF 00 -> FF), read through the per-lane ring of zpx_k1_common.cuh
//     and a three-word register window: a symbol is one funnel shift, one table look-up and straight-line arithmetic;
//   * tables: every lane has its OWN first-level table in shared memory, [entry][lane] (progressive files carry
//     optimised tables, one set per scan, so the lanes of a warp share nothing): 9 bits for an AC pass, 7 bits for
//     each of the four DC tables of an interleaved DC pass; 16-bit entries that hold the symbol's bit counts ready
//     made.  Longer AC codes: the lane's canonical limit / offset / value arrays, also in shared memory;
//   * first passes only WRITE coefficients (one 2-byte store per non-zero coefficient into the zeroed grid) and
//     record, per block, which coefficients of the band they set and their signs: two 64-bit maps per block
//     (zig-zag order), maintained with atomic ORs;
//   * an AC refinement pass is three kernels.  k3l_refine_prep (parallel, one warp per block) turns every block's
//     non-zero map into the ascending list of its ZERO positions.  With that list the serial part needs no
//     coefficient and no bit operation: the target of a run of r zeros is list[zi + r], and the number of correction
//     bits passed on the way is (target - zig - r) -- they are skipped by count.  The serial lanes (k3l_level) only
//     find where every block starts in the stream.  k3l_refine_apply (parallel, one thread per block) then parses each
//     block again from its start and does the writes: a store for a new +-2^Al, a 32-bit atomic add on the word that
//     holds it for a correction bit, the map update.
// DC refinement is one bit per block: k3l_dc_refine reads them 32 at a time.
//
// Exactness of the add: a correction adds +-2^Al to the 16-bit half of a 32-bit word.  The script check guarantees
// bit Al of the magnitude is clear and the value stays inside int16, so the add neither carries nor borrows across
// the half (negative values are >= 0x8000 as unsigned halves).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "zpx_k1_common.cuh"

namespace zpx {

namespace {

constexpr int L_ALB = 9;             // first-level bits of a lane's AC table (refinement passes: 32 lanes per warp)
constexpr int L_FLB = 10;            // ... in an AC first pass, whose warps carry 16 lanes: optimised tables of full
                                     // run/size alphabets have enough 10-bit codes for one lane in 32 to miss a
                                     // 9-bit table at nearly every step (and the canonical search then stalls all)
constexpr int L_FLANES = 16;
constexpr int L_DLB = 7;             //                    of each of its DC tables
constexpr int L_RS = 128;            // bytes between ring words of a lane: 32 lanes x 4
constexpr int L_ZLB = 64;            // bytes of a zero-position list: up to 63 entries, [63] = their number
constexpr uint32_t SM_RING = 0;                             // [ring words][32] u32: 16 words, 32 in the refinement pass
constexpr uint32_t SM_LUT = SM_RING + 32 * 128;             // [512][32] u16 (AC refinement) / [1024][16] u16 (AC first) /
                                                            // [4][128][32] u16 (DC)
constexpr uint32_t SM_AUX = SM_LUT + (1u << L_ALB) * 64;    // AC: lim [32][8] u32, valoff [32][8] i32, vals [32][256] u8
constexpr uint32_t SM_LIM = SM_AUX, SM_VOFF = SM_AUX + 1024, SM_VALS = SM_AUX + 2048;
constexpr uint32_t SM_DESC = SM_AUX;                        // DC: [3][ZPX_MAX_BLK_PER_MCU][32] u32
constexpr uint32_t SM_UNZIG = SM_AUX + 2048 + 8192;         // [64] u8
constexpr uint32_t SM_ZL = SM_UNZIG + 64;                   // AC refinement: [3][32][L_ZLB] u8, the lists of the lanes' current
                                                            // block and of the two after it
constexpr uint32_t SM_BYTES = SM_ZL + 3 * 32 * L_ZLB;
static_assert(3 * ZPX_MAX_BLK_PER_MCU * 128 <= 2048 + 8192, "DC descriptors fit the AC tables' space");
static_assert(SM_BYTES <= 56 * 1024, "four CTAs per SM");

enum { T_DCF = 0, T_ACF = 1, T_ACR = 2 };

__device__ __forceinline__ int scan_type(const ZpxScanDev* sc) { return sc->ss == 0 ? T_DCF : sc->ah == 0 ? T_ACF : T_ACR; }

// stored slot (in shorts) of natural coefficient `nat` in a block whose rows are XOR-swizzled by key
__device__ __forceinline__ uint32_t cslot(uint32_t key, uint32_t nat) { return (((nat >> 3) ^ key) << 3) + (nat & 7u); }

__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t s) {  // v >> s, 0 for s >= 32
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

// ---- table entries -----------------------------------------------------------------------------------------------
// AC symbol (r, s) with an len-bit code, 16 bits:
//   bits 0-4   bits of the whole symbol: code + value bits (first pass), code + sign bit (refinement), code + run
//              bits (End-Of-Band run)
//   bits 5-9   code length
//   bits 10-13 r (zero run)
//   bit 14     a coefficient follows (s != 0; refinement: s == 1)
//   bit 15     End-Of-Band run (s == 0, r < 15)
//   neither flag: ZRL (r == 15); refinement only: r == 0 marks a symbol with s > 1 (UnexpectedHuffmanCode)
// 0 = code longer than the first level, or none
__device__ __forceinline__ uint32_t ac_entry(bool refine, uint32_t sym, uint32_t len) {
    const uint32_t r = sym >> 4, s = sym & 15u;
    if (s == 0) return r == 15 ? (len | len << 5 | 15u << 10) : ((len + r) | len << 5 | r << 10 | 0x8000u);
    if (refine) return s == 1 ? ((len + 1u) | len << 5 | r << 10 | 0x4000u) : (len | len << 5);
    return (len + s) | len << 5 | r << 10 | 0x4000u;
}
// DC symbol t: bits 0-5 code + value bits, bits 6-10 code length, bit 11: t > 16 (ExcessiveDCComponent)
__device__ __forceinline__ uint32_t dc_entry(uint32_t sym, uint32_t len) {
    return sym > 16 ? (len | len << 6 | 0x800u) : ((len + sym) | len << 6);
}

struct Lane {
    RingReader<L_RS> rd;
    Window<L_RS> win;
    __device__ __forceinline__ void init(bool active, uint32_t sm, int lane, const K1Params& P, const ZpxIntervalDev& iv) {
        if (active) rd.init(sm + SM_RING + (uint32_t)lane * 4u, P.ublob, iv.ustart, iv.ulen, 0);
        else rd.init_idle(sm + SM_RING + (uint32_t)lane * 4u);
        win.load(rd);
    }
    __device__ __forceinline__ uint32_t peek() const { return win.peek(rd.bitpos); }
    __device__ __forceinline__ void consume(uint32_t tot) {  // tot <= 32
        win.advance(rd, rd.bitpos, tot);
        rd.bitpos += tot;
    }
    __device__ __forceinline__ void seek(uint32_t pos) {
        rd.bitpos = pos;
        win.load(rd);
    }
};

// the lane's AC symbol at the top of hi when the first level has no entry: canonical search over the lane's limits,
// first match = the reference's rule (decoder.zig:946-969).  0 = no code matches
__device__ __forceinline__ uint32_t ac_long(bool refine, uint32_t sm, int lane, uint32_t hi) {  // (lengths 10..16: a
    // 10-bit code that a 10-bit first level would have held is simply never asked for)
    const uint32_t v16 = hi >> 16;
    const uint4 a = lds_u128(sm + SM_LIM + (uint32_t)lane * 32u), b = lds_u128(sm + SM_LIM + (uint32_t)lane * 32u + 16u);
    int len = 0;
    if (v16 < b.z) len = 16;
    if (v16 < b.y) len = 15;
    if (v16 < b.x) len = 14;
    if (v16 < a.w) len = 13;
    if (v16 < a.z) len = 12;
    if (v16 < a.y) len = 11;
    if (v16 < a.x) len = 10;
    if (len == 0) return 0;
    const int off = (int)lds_u32(sm + SM_VOFF + (uint32_t)lane * 32u + (uint32_t)(len - 10) * 4u);
    const uint32_t sym = lds_u8(sm + SM_VALS + (uint32_t)lane * 256u + (uint32_t)((off + (int)(v16 >> (16 - len))) & 0xff));
    return ac_entry(refine, sym, (uint32_t)len);
}

// DC code longer than the first level: canonical search in the table in HBM
__device__ __noinline__ uint32_t dc_long(const ZpxHuffDev* __restrict__ t, uint32_t hi) {
    const uint32_t v16 = hi >> 16;
    for (int l = L_DLB + 1; l <= 16; l++) {
        if (v16 < __ldg(&t->limit[l])) {
            const uint32_t sym = __ldg(&t->vals[(__ldg(&t->valoff[l]) + (int)(v16 >> (16 - l))) & 0xff]);
            return dc_entry(sym, (uint32_t)l);
        }
    }
    return 0;
}

struct Work {
    bool active;
    ZpxIntervalDev iv;
    const ZpxScanDev* sc;
    const ZpxImageDev* im;
};

__device__ __forceinline__ int eof_code(const ZpxIntervalDev& iv) { return (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00; }

// ---------------------------------------------------------------------------------------------------------------
// DC first pass (decoder.zig:1366-1376): Ss = Se = 0, Ah = 0; interleaved or not
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void run_dc_first(const K1Params& P, const Work& w, const uint32_t sm, const int lane) {
    // per block of the lane's MCU, three words [word][c][lane]: the block's index (relative to the image) in MCU (0, 0);
    // the step of an MCU row (v * bw); h | hx << 8 | comp << 16 | undefined table << 24
    const ZpxScanDev* sc = w.sc;
    const ZpxImageDev* im = w.im;
    uint32_t nblk = 1, mxx = 1, mx = 0, my = 0;
    if (w.active) {
        const bool inter = sc->interleaved != 0;
        nblk = inter ? (uint32_t)sc->nblk : 1u;
        for (uint32_t c = 0; c < nblk; c++) {
            const int comp = sc->blk_comp[c];
            const uint32_t bw = (uint32_t)im->comp_bw[comp];
            const uint32_t h = inter ? im->h[comp] : 1u, v = inter ? im->v[comp] : 1u;
            const uint32_t hx = inter ? sc->blk_hx[c] : 0u, vy = inter ? sc->blk_vy[c] : 0u;
            const uint32_t undef = (sc->blk_pack[c][3] >> 16) & 1u;
            const uint32_t a = sm + SM_DESC + (c * 32u + (uint32_t)lane) * 4u;
            sts_u32(a, (uint32_t)(im->comp_base[comp] - im->coef_base) + vy * bw + hx);
            sts_u32(a + ZPX_MAX_BLK_PER_MCU * 128u, v * bw);
            sts_u32(a + 2u * ZPX_MAX_BLK_PER_MCU * 128u, h | hx << 8 | (uint32_t)comp << 16 | undef << 24);
        }
        mxx = inter ? (uint32_t)im->mxx : (uint32_t)sc->cw;
        const uint32_t first = inter ? w.iv.first_mcu : w.iv.first_block;
        my = first / mxx;
        mx = first - my * mxx;
    }
    Lane L;
    L.init(w.active, sm, lane, P, w.iv);
    const uint32_t n = w.active ? w.iv.n_blocks : 0u;
    const int al = w.active ? sc->al : 0;
    short* const cimg = reinterpret_cast<short*>(P.coef) + (w.active ? im->coef_base * 64ull : 0ull);
    const uint32_t lut = sm + SM_LUT + (uint32_t)lane * 2u;
    uint32_t j = 0, c = 0;
    int d0 = 0, d1 = 0, d2 = 0, d3 = 0, err = 0;
    while (j < n) {
        L.rd.topup();
        const uint32_t a = sm + SM_DESC + (c * 32u + (uint32_t)lane) * 4u;
        const uint32_t dz = lds_u32(a + 2u * ZPX_MAX_BLK_PER_MCU * 128u);
        const uint32_t comp = (dz >> 16) & 3u;
        if (dz >> 24) { err = ZPX_E_UninitializedHuffmanTable; break; }
        const uint32_t hi = L.peek();
        uint32_t e = lds_u16(lut + (comp * (1u << L_DLB) + (hi >> (32 - L_DLB))) * 64u);
        if (e == 0) {
            e = dc_long(&P.huff[sc->blk_dc[c]], hi);
            if (e == 0) { L.rd.bitpos += 16; err = ZPX_E_BadHuffmanCode; break; }
        }
        const uint32_t tot = e & 63u, len = (e >> 6) & 31u;
        if (e & 0x800u) { L.rd.bitpos += len; err = ZPX_E_ExcessiveDCComponent; break; }
        // RECEIVE + EXTEND (decoder.zig:1115-1134) on the tot - len bits after the code: all inside hi
        const uint32_t t = hi << len, s32 = 32u - (tot - len);
        const uint32_t raw = shr_clamp(t, s32);
        const int diff = (int)t < 0 ? (int)raw : (int)(raw - shr_clamp(0xffffffffu, s32));
        int d = comp == 0 ? d0 : comp == 1 ? d1 : comp == 2 ? d2 : d3;
        d += diff;
        if (comp == 0) d0 = d; else if (comp == 1) d1 = d; else if (comp == 2) d2 = d; else d3 = d;
        L.consume(tot);
        if (L.rd.overrun()) break;
        const int v = (int)((uint32_t)d << al);
        // hard error, as in zpx_k3.cu: later scans parse according to which coefficients are non-zero
        if (v < -32768 || v > 32767) { err = ZPX_E_COEF_RANGE; break; }
        const uint32_t h = dz & 0xffu, hx = (dz >> 8) & 0xffu;
        const uint32_t blk = lds_u32(a) + my * lds_u32(a + ZPX_MAX_BLK_PER_MCU * 128u) + mx * h;
        cimg[(uint64_t)blk * 64u + ((mx * h + hx) & 7u) * 8u] = (short)v;
        j++;
        if (++c == nblk) {
            c = 0;
            if (++mx == mxx) { mx = 0; my++; }
        }
    }
    if (w.active) {
        if (L.rd.overrun()) err = eof_code(w.iv);
        if (err) report(P.status, im->status_slot, sc->scan_index, (uint64_t)w.iv.first_block + j, err);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// AC first pass of a band (decoder.zig:1378-1411): Ss > 0, Ah = 0, one component.
// Warp-synchronous: the 32 lanes step through their symbols together, L_STEPS symbols between two top-ups of the
// rings; a lane that is done (or failed) idles on null symbols.
// ---------------------------------------------------------------------------------------------------------------
constexpr int L_STEPS = 8;  // a symbol takes at most 31 bits: 8 of them 31 bytes, the window reaches 12 further, a
                            // topped-up ring holds 49

__device__ __forceinline__ uint32_t* scan_pos(const K1Params& P, const ZpxImageDev* im, const ZpxScanDev* sc) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(P.coef) + im->ppos_base * 128ull) + sc->pos_off;
}
__device__ __forceinline__ uint8_t* scan_lists(const K1Params& P, const ZpxImageDev* im, const ZpxScanDev* sc) {
    return reinterpret_cast<uint8_t*>(P.coef) + im->pzl_base * 128ull + (uint64_t)sc->zl_off * L_ZLB;
}

// The serial part: where every block that has bits starts (pos[b] = bit position + 1; blocks inside an End-Of-Band run
// keep the 0 the area was cleared to).  k3l_first_apply decodes the blocks.
__device__ __forceinline__ void run_ac_first(const K1Params& P, const Work& w, const uint32_t sm, const int lane) {
    const ZpxScanDev* sc = w.sc;
    const ZpxImageDev* im = w.im;
    Lane L;
    L.init(w.active, sm, lane, P, w.iv);
    const uint32_t n = w.active ? w.iv.n_blocks : 0u;
    int ss = 1, se = 63;
    bool undef = false;
    uint32_t* pos = nullptr;
    uint32_t* done = nullptr;
    if (w.active) {
        ss = sc->ss;
        se = sc->se;
        undef = ((sc->blk_pack[0][3] >> 17) & 1u) != 0;
        pos = scan_pos(P, im, sc);
        done = pos + (uint32_t)sc->cw * (uint32_t)sc->ch + w.iv.ordinal;
        pos += w.iv.first_block;
    }
    const uint32_t lut = sm + SM_LUT + (uint32_t)(lane & (L_FLANES - 1)) * 2u;  // (lanes 16..31 never have a stream)
    uint32_t j = 0, eob = 0;
    int zig = ss, err = 0;
    if (undef && n) err = ZPX_E_UninitializedHuffmanTable;
    bool act = n != 0 && !err;
    if (act) pos[0] = 1u;
    while (__any_sync(0xffffffffu, act)) {
        L.rd.topup();
#pragma unroll
        for (int u = 0; u < L_STEPS; u++) {
            const uint32_t hi = L.peek();
            uint32_t e = lds_u16(lut + (hi >> (32 - L_FLB)) * (2u * L_FLANES));
            if (act && e == 0) {
                e = ac_long(false, sm, lane, hi);
                if (e == 0) { L.rd.bitpos += 16; err = ZPX_E_BadHuffmanCode; act = false; }
            }
            if (!act) e = 0;  // null symbol: no bits, no block end
            const uint32_t tot = e & 31u, len = (e >> 5) & 31u, r = (e >> 10) & 15u, nx = tot - len;
            const bool iseob = (e & 0x8000u) != 0, iscoef = (e & 0x4000u) != 0;
            const int zc = zig + (int)r;
            // a coefficient past the band's end: its value bits stay unread and the block is over (decoder.zig:1392-1394)
            L.consume(iscoef && zc > se ? len : tot);
            // EOBn: the rest of the band of this block and of the next eob blocks (decoder.zig:1399-1407)
            if (iseob) eob = (1u << nx) + shr_clamp(hi << len, 32u - nx) - 1u;
            zig = iseob ? 64 : zc + (act ? 1 : 0);
            if (act && zig > se) {
                if (L.rd.overrun()) {
                    act = false;
                } else {
                    const uint32_t skip = min(eob, n - j - 1u);
                    eob -= skip;
                    j += 1u + skip;
                    act = j < n;
                    if (act) pos[j] = L.rd.bitpos + 1u;
                    zig = ss;
                }
            }
        }
    }
    if (w.active) {
        *done = j;
        if (L.rd.overrun()) err = eof_code(w.iv);
        if (err) report(P.status, im->status_slot, sc->scan_index, (uint64_t)w.iv.first_block + j, err);
        // an End-Of-Band run still open at the end of a scan: see zpx_k3.cu (the reference carries it into the next scan)
        else if ((w.iv.flags & 2u) && eob != 0)
            report(P.status, im->status_slot, sc->scan_index, (uint64_t)w.iv.first_block + w.iv.n_blocks, ZPX_E_UNSUPPORTED_STREAM);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// AC refinement of a band (decoder.zig:1468-1517, refineNonZeroes :1522-1549): Ss > 0, Ah > 0, one component.
// Per-scan scratch in the image's part of the coefficient buffer (b = ordinal of a coded block in the scan):
//   pos[b]     bit position of the block's first bit in its interval's stream, bit 31 = the block lies inside an
//              End-Of-Band run (it has no symbols)
//   pos[B + i] B = coded blocks of the scan, i = interval ordinal: blocks of interval i the serial pass got through
//   list[b]    L_ZLB bytes: zig-zag positions of the band's zero coefficients, ascending; [63] = their number
// ---------------------------------------------------------------------------------------------------------------
// asynchronous copy of one list into a lane's shared-memory slot (no registers, nothing waits until wait_lists)
__device__ __forceinline__ void copy_list(uint32_t dst, const uint8_t* src) {
#pragma unroll
    for (int i = 0; i < L_ZLB / 16; i++)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * i) : "memory");
}
__device__ __forceinline__ void commit_lists() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_lists() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }  // all but the last two copies

// The serial part: where every block of the interval starts.  Warp-synchronous like the first pass; a step is a
// symbol (or a whole block inside an End-Of-Band run) and moves by up to 30 + 63 bits, so the ring is twice as long
// and read directly (two words per step).
constexpr int R_RW = 32;     // ring words per lane
constexpr int R_STEPS = 8;   // 8 x 93 bits = 93 bytes, a read reaches 8 further, a topped-up ring holds 113

__device__ __forceinline__ void run_ac_refine(const K1Params& P, const Work& w, const uint32_t sm, const int lane) {
    const ZpxScanDev* sc = w.sc;
    const ZpxImageDev* im = w.im;
    RingReader<L_RS, R_RW> rd;
    if (w.active) rd.init(sm + SM_RING + (uint32_t)lane * 4u, P.ublob, w.iv.ustart, w.iv.ulen, 0);
    else rd.init_idle(sm + SM_RING + (uint32_t)lane * 4u);
    const uint32_t n = w.active ? w.iv.n_blocks : 0u;
    int ss = 1, se = 63;
    bool undef = false;
    uint32_t* pos = nullptr;
    const uint8_t* lp = nullptr;   // list of the block three after the current one: the next to be copied in
    uint32_t* done = nullptr;
    if (w.active) {
        ss = sc->ss;
        se = sc->se;
        undef = ((sc->blk_pack[0][3] >> 17) & 1u) != 0;
        pos = scan_pos(P, im, sc);
        done = pos + (uint32_t)sc->cw * (uint32_t)sc->ch + w.iv.ordinal;
        pos += w.iv.first_block;
        lp = scan_lists(P, im, sc) + (uint64_t)w.iv.first_block * L_ZLB;
    }
    const uint32_t lut = sm + SM_LUT + (uint32_t)lane * 2u;
    uint32_t j = 0, eob = 0, zi = 0;
    int zig = ss, err = 0;
    // the lists of blocks j, j + 1, j + 2 sit in (or are on their way to) the lane's three slots
    const uint32_t zl0 = sm + SM_ZL + (uint32_t)lane * L_ZLB, zl_end = zl0 + 3u * 32u * L_ZLB;
    uint32_t zl = zl0;  // slot of the current block
    for (uint32_t k = 0; k < 3; k++) {
        if (k < n) copy_list(zl0 + k * 32u * L_ZLB, lp);
        commit_lists();
        lp += L_ZLB;
    }
    wait_lists();
    uint32_t nzeros = lds_u8(zl + L_ZLB - 1);
    if (n) pos[0] = 0;
    if (undef && n) err = ZPX_E_UninitializedHuffmanTable;  // (the first block of a scan starts with a symbol)
    bool act = n != 0 && !err;
    const uint32_t endbits = rd.endbits;
    while (__any_sync(0xffffffffu, act)) {
        rd.topup();
        // One step = one symbol, or one whole block inside an End-Of-Band run.  Straight-line code: everything a step
        // may do (the block end included) is predicated, only the rare cases branch.
#pragma unroll
        for (int u = 0; u < R_STEPS; u++) {
            const uint32_t hi = rd.peek();
            uint32_t e = lds_u16(lut + (hi >> (32 - L_ALB)) * 64u);
            if (act && eob == 0 && e == 0) {
                e = ac_long(true, sm, lane, hi);
                if (e == 0) { rd.bitpos += 16; err = ZPX_E_BadHuffmanCode; act = false; }
            }
            if (!act || eob != 0) e = 0;  // no symbol: a lane that is done, or a block inside a run
            const uint32_t tot = e & 31u, len = (e >> 5) & 31u, r = (e >> 10) & 15u, nx = tot - len;
            const bool iseob = (e & 0x8000u) != 0, iscoef = (e & 0x4000u) != 0;
            // correction bits of the block's rest: its non-zero coefficients at or after zig
            const uint32_t tail = (uint32_t)(se + 1 - zig) - (nzeros - zi);
            // a run of r zeros, then a new coefficient (or ZRL's 16th zero): the (r + 1)-th zero coefficient at or after
            // zig; the non-zero ones passed on the way take a bit each
            bool target = e != 0 && !iseob;
            const uint32_t zr = zi + r;
            const uint32_t t = lds_u8(zl + min(zr, (uint32_t)L_ZLB - 2u));
            if (target && ((!iscoef && r != 15u) || zr >= nzeros)) {
                if (!iscoef && r != 15u) {
                    rd.bitpos += len;
                    err = ZPX_E_UnexpectedHuffmanCode;
                } else {
                    rd.bitpos += tot + tail;  // (after the block's remaining correction bits, as the reference)
                    err = ZPX_E_TooManyCoefficients;
                }
                act = false;
                target = false;
            }
            // EOBn (decoder.zig:1480-1488): this block's rest and the next eob - 1 blocks
            {
                const uint32_t run = (1u << nx) + shr_clamp(hi << len, 32u - nx);
                asm("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %2, %0, p;}" : "+r"(eob) : "r"((uint32_t)iseob), "r"(run));
            }
            rd.bitpos += !act ? 0u : target ? tot + (t - (uint32_t)zig - r) : tot + tail;
            zi = target ? zr + 1u : zi;
            zig = target ? (int)t + 1 : zig;
            const bool fin = act && (!target || zig > se);
            eob -= (fin && !target) ? 1u : 0u;
            const bool be = fin && rd.bitpos <= endbits;  // block end (a block that ran past the stream's end fails below)
            j += be ? 1u : 0u;
            act = act && (!fin || be) && j < n;
            // ---- block end: the only common branch of a step ----
            if (be) {
                if (act) pos[j] = rd.bitpos | (eob ? 0x80000000u : 0u);
                // the slot just finished takes the list of block j + 2; block j's own copy was issued two blocks ago
                if (j + 2 < n) copy_list(zl, lp);
                commit_lists();
                lp += L_ZLB;
                zl += 32u * L_ZLB;
                if (zl == zl_end) zl = zl0;
                wait_lists();
                nzeros = lds_u8(zl + (uint32_t)L_ZLB - 1u);
                zig = ss;
                zi = 0;
            }
        }
    }
    if (w.active) {
        *done = j;
        if (rd.overrun()) err = eof_code(w.iv);
        if (err) report(P.status, im->status_slot, sc->scan_index, (uint64_t)w.iv.first_block + j, err);
        else if ((w.iv.flags & 2u) && eob != 0)
            report(P.status, im->status_slot, sc->scan_index, (uint64_t)w.iv.first_block + w.iv.n_blocks, ZPX_E_UNSUPPORTED_STREAM);
    }
}

// one lane's first-level table into shared memory, written by the whole warp.  Source: the 10-bit table of ZpxHuffDev
template <typename F>
__device__ __forceinline__ void stage_lut(uint32_t dst /* shared address of entry 0 of the lane's column */, const ZpxHuffDev* __restrict__ t,
                                          int bits, uint32_t stride /* bytes between entries */, int lane, F entry) {
    for (int i = lane; i < (1 << bits); i += 32) {
        const uint32_t e = __ldg(&t->lut[i << (ZPX_LUT_BITS - bits)]);
        const uint32_t len = e & 0xffu;
        sts_u16(dst + (uint32_t)i * stride, (len == 0 || len > (uint32_t)bits) ? 0 : (int)entry(e >> 8, len));
    }
}

// 32 bits of an unstuffed stream from bit position p (plain loads: the parallel kernels)
__device__ __forceinline__ uint32_t bits_at(const uint8_t* __restrict__ src, uint32_t p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src) + (p >> 5);
    return __funnelshift_l(__byte_perm(__ldg(w + 1), 0, 0x0123), __byte_perm(__ldg(w), 0, 0x0123), p);
}

}  // namespace

// one warp per CTA; list: interval indices grouped by pass type, every group padded to a multiple of 32 with ~0 (AC
// first passes: 16 intervals and 16 x ~0 per warp)
__global__ void __launch_bounds__(32) k3l_level(const K1Params P, const uint32_t* __restrict__ list) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    const uint32_t sm = smem_addr(s_raw);
    const int lane = threadIdx.x;
    const uint32_t ix = list[blockIdx.x * 32 + lane];
    Work w;
    w.active = ix != 0xffffffffu;
    w.iv = P.ivs[w.active ? ix : list[blockIdx.x * 32]];  // (a group's first entry is always real)
    w.sc = &P.scans[w.iv.scan];
    w.im = &P.imgs[w.sc->img];
    const int type = scan_type(&P.scans[P.ivs[list[blockIdx.x * 32]].scan]);
    s_raw[SM_UNZIG + lane] = c_unzig[lane];
    s_raw[SM_UNZIG + 32 + lane] = c_unzig[32 + lane];
    // ---- every lane's tables, staged by the whole warp ----
    for (int l = 0; l < 32; l++) {
        const uint32_t lix = __shfl_sync(0xffffffffu, ix, l);
        if (lix == 0xffffffffu) continue;
        const ZpxScanDev* __restrict__ sc = &P.scans[P.ivs[lix].scan];
        if (type == T_DCF) {
            const int nblk = sc->interleaved ? sc->nblk : 1;
            uint32_t done = 0;
            for (int c = 0; c < nblk; c++) {
                const uint32_t comp = sc->blk_comp[c];
                if (done >> comp & 1u) continue;
                done |= 1u << comp;
                stage_lut(sm + SM_LUT + (comp * (1u << L_DLB) * 32u + (uint32_t)l) * 2u, &P.huff[sc->blk_dc[c]], L_DLB, 64u, lane,
                          [](uint32_t sym, uint32_t len) { return dc_entry(sym, len); });
            }
        } else {
            const ZpxHuffDev* __restrict__ t = &P.huff[sc->blk_ac[0]];
            const bool refine = type == T_ACR;
            // (an AC first pass keeps streams in lanes 0..15 only: the host pads its list that way)
            stage_lut(sm + SM_LUT + (uint32_t)l * 2u, t, refine ? L_ALB : L_FLB, refine ? 64u : 2u * L_FLANES, lane,
                      [refine](uint32_t sym, uint32_t len) { return ac_entry(refine, sym, len); });
            if (lane < 7) {
                sts_u32(sm + SM_LIM + (uint32_t)l * 32u + (uint32_t)lane * 4u, __ldg(&t->limit[10 + lane]));
                sts_u32(sm + SM_VOFF + (uint32_t)l * 32u + (uint32_t)lane * 4u, (uint32_t)__ldg(&t->valoff[10 + lane]));
            }
            for (int i = lane; i < 64; i += 32)
                sts_u32(sm + SM_VALS + (uint32_t)l * 256u + (uint32_t)i * 4u, __ldg(reinterpret_cast<const uint32_t*>(t->vals) + i));
        }
    }
    __syncwarp();
    if ((P.dbg & 2) && type == T_DCF) return;  // experiments only (ZPX_K1_DBG): time a level without one pass type
    if ((P.dbg & 4) && type == T_ACF) return;
    if (type == T_DCF) run_dc_first(P, w, sm, lane);
    else if (type == T_ACF) run_ac_first(P, w, sm, lane);
    else run_ac_refine(P, w, sm, lane);
}

// AC refinement, before the serial pass: the zero-position list of every block.  blockIdx.y = entry of the list of
// intervals, blockIdx.x = group of 128 blocks of the interval; one warp per block at a time, lane p looks at zig-zag
// positions p and p + 32
__global__ void __launch_bounds__(128) k3l_refine_prep(const K1Params P, const uint32_t* __restrict__ list) {
    const uint32_t ix = list[blockIdx.y];
    if (ix == 0xffffffffu) return;
    const ZpxIntervalDev& iv = P.ivs[ix];
    const uint32_t n = iv.n_blocks, j0 = blockIdx.x * 128u + (threadIdx.x >> 5) * 32u;
    if (j0 >= n) return;
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const uint32_t lane = threadIdx.x & 31u;
    const int ss = sc->ss, se = sc->se, comp = sc->blk_comp[0];
    const uint32_t cw = (uint32_t)sc->cw, bw = (uint32_t)im->comp_bw[comp];
    const uint2* __restrict__ maps = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(P.coef) + im->pmask_base * 128ull +
                                                                   (im->comp_base[comp] - im->coef_base) * 16ull);
    uint8_t* out = scan_lists(P, im, sc) + (uint64_t)(iv.first_block + j0) * L_ZLB;
    const unsigned long long band = (se == 63 ? ~0ull : (1ull << (se + 1)) - 1ull) & ~((1ull << ss) - 1ull);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t q = iv.first_block + j0, by = q / cw, bx = q - by * cw;
    const uint32_t jn = min(32u, n - j0);
    for (uint32_t k = 0; k < jn; k++, out += L_ZLB) {
        const uint2 nz = __ldg(maps + 2ull * ((uint64_t)by * bw + bx));
        const uint32_t zlo = ~nz.x & (uint32_t)band, zhi = ~nz.y & (uint32_t)(band >> 32);
        const uint32_t clo = (uint32_t)__popc(zlo), nzeros = clo + (uint32_t)__popc(zhi);
        if (zlo >> lane & 1u) out[__popc(zlo & lt)] = (uint8_t)lane;
        if (zhi >> lane & 1u) out[clo + (uint32_t)__popc(zhi & lt)] = (uint8_t)(lane + 32u);
        if (lane == 0) out[L_ZLB - 1] = (uint8_t)nzeros;  // (entries past the last zero are never read)
        if (++bx == cw) { bx = 0; by++; }
    }
}

// AC first pass, after the serial pass: every block that has bits decoded from its start, one thread per block:
// coefficients and the non-zero / sign maps (decoder.zig:1378-1411).  Grid as k3l_refine_prep
__global__ void __launch_bounds__(128) k3l_first_apply(const K1Params P, const uint32_t* __restrict__ list) {
    __shared__ uint16_t s_lut[1 << L_ALB];  // sym << 8 | len, 0 = longer
    __shared__ uint8_t s_unzig[64];
    const uint32_t ix = list[blockIdx.y];
    if (ix == 0xffffffffu) return;
    const ZpxIntervalDev& iv = P.ivs[ix];
    if (blockIdx.x * 128u >= iv.n_blocks) return;
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const ZpxHuffDev* __restrict__ tab = &P.huff[sc->blk_ac[0]];
    for (int i = threadIdx.x; i < (1 << L_ALB); i += 128) {
        const uint32_t e = __ldg(&tab->lut[i << (ZPX_LUT_BITS - L_ALB)]);
        s_lut[i] = (e & 0xffu) <= (uint32_t)L_ALB ? (uint16_t)e : (uint16_t)0;
    }
    if (threadIdx.x < 64) s_unzig[threadIdx.x] = c_unzig[threadIdx.x];
    __syncthreads();
    const uint32_t* pos = scan_pos(P, im, sc);
    const uint32_t done = pos[(uint32_t)sc->cw * (uint32_t)sc->ch + iv.ordinal];
    const uint32_t j = blockIdx.x * 128u + threadIdx.x;
    if (j >= done) return;
    const uint32_t p0 = pos[iv.first_block + j];
    if (p0 == 0) return;  // inside an End-Of-Band run: nothing coded
    const int se = sc->se, al = sc->al, comp = sc->blk_comp[0];
    const uint32_t cw = (uint32_t)sc->cw, bw = (uint32_t)im->comp_bw[comp];
    const uint32_t q = iv.first_block + j, by = q / cw, bx = q - by * cw;
    const uint64_t blkix = (uint64_t)by * bw + bx;
    short* const blk = reinterpret_cast<short*>(P.coef) + (im->comp_base[comp] + blkix) * 64ull;
    const uint8_t* __restrict__ src = P.ublob + iv.ustart;
    const uint32_t key = bx & 7u;
    unsigned long long nz = 0, sg = 0;
    uint32_t p = p0 - 1u;
    int zig = sc->ss;
    while (zig <= se) {
        const uint32_t hi = bits_at(src, p);
        uint32_t e = s_lut[hi >> (32 - L_ALB)];
        if (e == 0) {
            const HuffSym hs = huff_decode(tab, hi);
            if (hs.len == 0) break;
            e = hs.sym << 8 | (uint32_t)hs.len;
        }
        const uint32_t len = e & 0xffu, r = e >> 12, s = (e >> 8) & 15u;
        if (s == 0) {
            if (r != 15u) break;  // EOBn
            zig += 16;
            p += len;
            continue;
        }
        zig += (int)r;
        if (zig > se) break;
        const uint32_t t = hi << len;
        const int v = (int)((uint32_t)((int)t < 0 ? (int)(t >> (32 - s)) : (int)(t >> (32 - s)) - (int)((1u << s) - 1u)) << al);
        if (v < -32768 || v > 32767) {
            // hard error, as in zpx_k3.cu: later scans parse according to which coefficients are non-zero
            report(P.status, im->status_slot, sc->scan_index, (uint64_t)q, ZPX_E_COEF_RANGE);
            break;
        }
        blk[cslot(key, s_unzig[zig])] = (short)v;
        nz |= 1ull << zig;
        if (v < 0) sg |= 1ull << zig;
        p += len + s;
        zig++;
    }
    if (nz) {
        unsigned long long* m = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(P.coef) + im->pmask_base * 128ull +
                                                                      (im->comp_base[comp] - im->coef_base + blkix) * 16ull);
        atomicOr(m, nz);
        if (sg) atomicOr(m + 1, sg);
    }
}

// AC refinement, after the serial pass: every block parsed again from its start, one thread per block, and written:
// new coefficients, correction bits, map update (decoder.zig:1468-1549).  Grid as k3l_refine_prep
__global__ void __launch_bounds__(128) k3l_refine_apply(const K1Params P, const uint32_t* __restrict__ list) {
    __shared__ uint16_t s_lut[1 << L_ALB];  // sym << 8 | len, 0 = longer
    __shared__ uint8_t s_unzig[64];
    const uint32_t ix = list[blockIdx.y];
    if (ix == 0xffffffffu) return;
    const ZpxIntervalDev& iv = P.ivs[ix];
    if (blockIdx.x * 128u >= iv.n_blocks) return;
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const ZpxHuffDev* __restrict__ tab = &P.huff[sc->blk_ac[0]];
    for (int i = threadIdx.x; i < (1 << L_ALB); i += 128) {
        const uint32_t e = __ldg(&tab->lut[i << (ZPX_LUT_BITS - L_ALB)]);
        s_lut[i] = (e & 0xffu) <= (uint32_t)L_ALB ? (uint16_t)e : (uint16_t)0;
    }
    if (threadIdx.x < 64) s_unzig[threadIdx.x] = c_unzig[threadIdx.x];
    __syncthreads();
    const uint32_t* pos = scan_pos(P, im, sc);
    const uint32_t done = pos[(uint32_t)sc->cw * (uint32_t)sc->ch + iv.ordinal];
    const uint32_t j = blockIdx.x * 128u + threadIdx.x;
    if (j >= done) return;
    const uint32_t p0 = pos[iv.first_block + j];
    const int ss = sc->ss, se = sc->se, comp = sc->blk_comp[0];
    const int delta = 1 << sc->al;
    const uint32_t cw = (uint32_t)sc->cw, bw = (uint32_t)im->comp_bw[comp];
    const uint32_t q = iv.first_block + j, by = q / cw, bx = q - by * cw;
    const uint64_t blkix = (uint64_t)by * bw + bx;
    short* const blk = reinterpret_cast<short*>(P.coef) + (im->comp_base[comp] + blkix) * 64ull;
    uint4* const mp = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(P.coef) + im->pmask_base * 128ull + (im->comp_base[comp] - im->coef_base) * 16ull) + blkix;
    const uint4 m = *mp;
    const unsigned long long band = (se == 63 ? ~0ull : (1ull << (se + 1)) - 1ull) & ~((1ull << ss) - 1ull);
    const unsigned long long nz = ((unsigned long long)m.y << 32 | m.x) & band, sg = (unsigned long long)m.w << 32 | m.z;
    const uint8_t* __restrict__ src = P.ublob + iv.ustart;
    const uint32_t key = bx & 7u;
    unsigned long long newnz = 0, newsg = 0;
    uint32_t p = p0 & 0x7fffffffu;
    bool inrun = (p0 >> 31) != 0;
    int zig = ss;
    // correction bits (decoder.zig:1538-1547) of the non-zero coefficients cm, taken from bit position cp on
    auto correct = [&](unsigned long long cm, uint32_t cp) {
        uint32_t c = 0, left = 0;
        while (cm) {
            if (left == 0) { c = bits_at(src, cp); cp += 32; left = 32; }
            const int z = __ffsll((long long)cm) - 1;
            cm &= cm - 1ull;
            if ((int)c < 0) {
                const uint32_t slot = cslot(key, s_unzig[z]);
                const int d = (sg >> z & 1ull) ? -delta : delta;
                atomicAdd(reinterpret_cast<unsigned int*>(blk + (slot & ~1u)), (unsigned int)d << ((slot & 1u) * 16u));
            }
            c <<= 1;
            left--;
        }
    };
    for (int guard = 0; guard < 64 && zig <= se; guard++) {
        const unsigned long long from = ~((1ull << zig) - 1ull);
        if (inrun) {
            correct(nz & from, p);
            break;
        }
        const uint32_t hi = bits_at(src, p);
        uint32_t e = s_lut[hi >> (32 - L_ALB)];
        if (e == 0) {
            const HuffSym hs = huff_decode(tab, hi);
            if (hs.len == 0) break;
            e = hs.sym << 8 | (uint32_t)hs.len;
        }
        const uint32_t len = e & 0xffu, r = e >> 12, s = (e >> 8) & 15u;
        if (s == 0 && r != 15u) {
            p += len + r;
            inrun = true;  // the block's rest: next iteration
            continue;
        }
        if (s > 1) break;
        int z = 0;
        if (s == 1) z = ((hi << len) >> 31) ? delta : -delta;
        p += len + s;
        // the (r + 1)-th zero at or after zig
        unsigned long long zm = ~nz & band & from;
        for (uint32_t i = 0; i < r && zm; i++) zm &= zm - 1ull;
        if (zm == 0) break;
        const int t = __ffsll((long long)zm) - 1;
        const unsigned long long cm = nz & from & ((1ull << t) - 1ull);
        correct(cm, p);
        p += (uint32_t)__popcll(cm);
        if (z != 0) {
            blk[cslot(key, s_unzig[t])] = (short)z;
            newnz |= 1ull << t;
            if (z < 0) newsg |= 1ull << t;
        }
        zig = t + 1;
    }
    if (newnz) {
        unsigned long long* mm = reinterpret_cast<unsigned long long*>(mp);
        atomicOr(mm, newnz);
        if (newsg) atomicOr(mm + 1, newsg);
    }
}

// DC refinement (decoder.zig:1461-1468): block j of the interval takes bit j of its stream.  blockIdx.y = entry of the
// list of intervals, blockIdx.x = group of DCR_CHUNK blocks of the interval, a warp takes 32 blocks at a time.
constexpr uint32_t DCR_CHUNK = 2048;
__global__ void __launch_bounds__(128) k3l_dc_refine(const K1Params P, const uint32_t* __restrict__ list) {
    const int lane = threadIdx.x & 31;
    const ZpxIntervalDev iv = P.ivs[list[blockIdx.y]];
    if (blockIdx.x * DCR_CHUNK >= iv.n_blocks) return;
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const bool inter = sc->interleaved != 0;
    const uint32_t nblk = inter ? (uint32_t)sc->nblk : 1u;
    const uint32_t mxx = (uint32_t)im->mxx, cw = (uint32_t)sc->cw;
    const uint32_t bits = iv.ulen * 8u, n = min(iv.n_blocks, bits);
    const uint8_t* __restrict__ src = P.ublob + iv.ustart;
    const unsigned int orv = (unsigned int)((1 << sc->al) & 0xffff);
    short* const cbase = reinterpret_cast<short*>(P.coef);
    const uint32_t c0 = blockIdx.x * DCR_CHUNK, c1 = min(n, c0 + DCR_CHUNK);
    for (uint32_t j0 = c0 + (threadIdx.x >> 5) * 32u; j0 < c1; j0 += 128) {
        const uint32_t j = j0 + (uint32_t)lane;
        // 32 bits of the stream: one word per warp iteration, read by every lane (j0 is a multiple of 32)
        const uint32_t word = __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(src + (j0 >> 3))), 0, 0x0123);
        if (j < n && (word >> (31 - lane) & 1u)) {
            uint32_t comp, bx, by;
            if (inter) {
                const uint32_t m = iv.first_mcu + j / nblk, c = j % nblk;
                const uint32_t my = m / mxx, mx = m - my * mxx;
                comp = sc->blk_comp[c];
                bx = im->h[comp] * mx + sc->blk_hx[c];
                by = im->v[comp] * my + sc->blk_vy[c];
            } else {
                const uint32_t q = iv.first_block + j;
                comp = sc->blk_comp[0];
                by = q / cw;
                bx = q - by * cw;
            }
            short* p = cbase + (im->comp_base[comp] + (uint64_t)by * (uint32_t)im->comp_bw[comp] + bx) * 64ull + (bx & 7u) * 8u;
            const uintptr_t a = reinterpret_cast<uintptr_t>(p);
            atomicOr(reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3), orv << ((a & 2) ? 16 : 0));
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && iv.n_blocks > bits)  // block `bits` is the first one without a bit
        report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + bits, eof_code(iv));
}

cudaError_t k3l_launch_level(const K1Params& P, const uint32_t* list, int n_padded, cudaStream_t s) {
    if (n_padded <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k3l_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES);
    if (e != cudaSuccess) return e;
    k3l_level<<<n_padded / 32, 32, SM_BYTES, s>>>(P, list);
    return cudaGetLastError();
}

cudaError_t k3l_launch_refine_prep(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s) {
    for (int y0 = 0; y0 < n_list; y0 += 65535) {
        const dim3 grid((max_blocks + 127) / 128, (unsigned)std::min(65535, n_list - y0));
        k3l_refine_prep<<<grid, 128, 0, s>>>(P, list + y0);
    }
    return cudaGetLastError();
}

cudaError_t k3l_launch_first_apply(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s) {
    for (int y0 = 0; y0 < n_list; y0 += 65535) {
        const dim3 grid((max_blocks + 127) / 128, (unsigned)std::min(65535, n_list - y0));
        k3l_first_apply<<<grid, 128, 0, s>>>(P, list + y0);
    }
    return cudaGetLastError();
}

cudaError_t k3l_launch_refine_apply(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s) {
    for (int y0 = 0; y0 < n_list; y0 += 65535) {
        const dim3 grid((max_blocks + 127) / 128, (unsigned)std::min(65535, n_list - y0));
        k3l_refine_apply<<<grid, 128, 0, s>>>(P, list + y0);
    }
    return cudaGetLastError();
}

cudaError_t k3l_launch_dc_refine(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s) {
    for (int y0 = 0; y0 < n_list; y0 += 65535) {
        const dim3 grid((max_blocks + DCR_CHUNK - 1) / DCR_CHUNK, (unsigned)std::min(65535, n_list - y0));
        k3l_dc_refine<<<grid, 128, 0, s>>>(P, list + y0);
    }
    return cudaGetLastError();
}

}  // namespace zpx
