// zpx_k1.cu -- Huffman / run-length coefficient decode on the GPU: the passes that WRITE coefficients.
//
// Replaces the MCU loop of processSos (src/jpeg/decoder.zig:1294-1452) together with
// decodeHuffman (:909-970), ensureNBits (:975-991), receiveExtend (:1115-1134) and decodeBits (:1009-1022)
// for sequential (SOF0/SOF1) scans; readByteStuffedByte (:712-749) is k0_unstuff's (zpx_k0.cu).
// Output: int16 coefficient blocks, natural (de-zigzagged) order, absolute DC, in the HBM layout
// k2 consumes (zpx_k2.cu header).
//
// Semantics kept from the reference, bit for bit on every conforming stream:
//   * MSB-first bit reader; a 0xFF followed by anything but 0x00 is never consumed.  The host hands each
//     restart interval its byte range [start, limit): limit = first such 0xFF (zpx_parse.cpp).  Bits past the
//     limit read as zero here and any symbol that needs them is the reference's MissingFF00 (or UnexpectedEof
//     at end of file).
//   * Huffman codes up to 16 bits, canonical; RECEIVE/EXTEND; DC prediction per component, reset at
//     each restart interval; AC run-length placement b[unzig[zig]].
//   * the End-Of-Band-run quirk of sequential scans inside one interval.
// Error kinds are reported per image through an atomicMin on (scan, block ordinal, code) so the
// first error in the reference's decode order wins.
//
//   k1_lane_per_interval   one lane per restart interval (32 intervals per warp), serial inside the interval
//   k1s_write              last pass of the self-synchronising decoder (zpx_k1s.cu): one lane per sub-sequence,
//                          started in its true state
// Both run the same block-synchronous loop over a RingReader (zpx_k1_common.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "zpx_k1_common.cuh"

#ifndef ZPX_K1_ONE_SELECT
#define ZPX_K1_ONE_SELECT 1
#endif
#ifndef ZPX_K1_VOTE_EVERY
#define ZPX_K1_VOTE_EVERY 2
#endif

namespace zpx {

// The run a scan's first interval starts with in the serial re-decode (never in the batch decode).  Out of line on
// purpose: inlined, the load shared a scoreboard with the ring's prefetch load and every block's first test of eob_run
// waited for that prefetch (long-scoreboard stalls 0.20 -> 0.51 per issue, ncu; 0.31 with this) although the load itself
// never ran.
static __device__ __noinline__ uint32_t k1_carry_in(const uint32_t* p) { return *p & 0xffffu; }

// slow-path wrapper shared by the DC and AC steps: returns the entry fields, sets err / eob_run / wide
template <bool SMEM>
__device__ __forceinline__ uint32_t k1_rare(const K1Params& P, uint32_t hi, bool isdc, uint32_t e, const uint4& bi, uint32_t tb,
                                            uint32_t& eob_run, int& err, int& wide) {
    unsigned long long r;
    if (SMEM) {
        r = k1_slow_symbol_sm(tb, isdc ? (bi.w >> 20) & 15u : (bi.w >> 24) & 15u, hi, isdc, e);
    } else {
        r = k1_slow_symbol(&P.huff[isdc ? bi.x : bi.y], hi, isdc, e);
    }
    e = (uint32_t)r;
    err = (int)(r >> 32);
    if (e >> 31) {
        // (r, 0) with 0 < r < 15 (decoder.zig:1399-1407): eob_run = (1 << r | next r bits) - 1
        const int len = (int)((e >> 8) & 0xffu), rr = (int)((e >> 16) & 0xffu);
        eob_run = (1u << rr) | ((hi << len) >> (32 - rr));
        eob_run = (eob_run - 1) & 0xffffu;
        e = ZPX_FE(len + rr, len, 0, 64, 0);  // consume code + run bits, end of block
    } else if (!isdc) {
        wide = max(wide, 32 - fe_s32(e));  // AC values of 13 bits or more only come this way (special entries)
    }
    // an error ends the block: advance past index 63 (the caller's err stays set for the block-end code)
    if (err) e = (e & 0x00ffffffu) | 64u << 24;
    return e;
}

// Where a lane starts inside its interval.  One lane per interval: the interval's first bit, block 0, all of its
// blocks.  Self-synchronising mode (SUB, zpx_k1s.cu): the true decoder state at the start of the lane's
// sub-sequence -- possibly inside a block, whose tail the lane skips because it belongs to the previous lane --
// and the number of blocks that start inside the sub-sequence, both known from the synchronisation passes.
struct K1Start {
    uint32_t bitpos;  // bit position in the interval's unstuffed stream
    int k;            // 0: at a block start; 1..63: inside a block, next zig-zag index k
    int dc0, dc1, dc2, dc3;  // DC predictors at that point
    uint32_t j0;      // ordinal (inside the interval) of the first block this lane writes
    uint32_t count;   // blocks this lane writes.  (SUB: the synchronisation pass steps over invalid codes bit by bit
                      // and counts a block that fails to start as started: the lane that reaches it in the true
                      // state decodes it here and reports the reference's error, BadHuffmanCode at the block that
                      // follows, decoder.zig:947-969)
};

// Block-synchronous main loop: the 32 lanes of a warp decode their k-th block together --
//   DC symbol (all lanes), AC symbols until every lane's block is complete (a warp vote per symbol; a
//   lane whose block ended idles), block end (all lanes: flush the 128-byte block, next descriptor).
// Lanes of a warp sit at the same block phase of the MCU (all on luma or all on chroma), so the number
// of AC iterations is close to the lanes' own symbol counts, and the block-end code runs once per
// block for all lanes instead of once per symbol for a few.
//
// The AC step is straight-line code that every lane executes (a lane whose block is complete advances by zero
// bits and stores nothing): the only branches are the rare-symbol call and the loop's vote.  The stream window is
// three ring words in registers -- the word under the bit position and the two after it -- so the 32 bits of a
// step come from one funnel shift; when a step crosses a word boundary the registers move up and the third is
// reloaded from the ring, a load nothing waits for until the step after the next.
// (Window<RS>: zpx_k1_common.cuh)

// what a lane that idles (its block is complete) or takes the rare path sees in the common code of a step:
// no bits, no value, no advance
constexpr uint32_t K1_NULL_E = ZPX_FE(0, 0, 0, 0, 0);

// shared-memory bytes per lane of the block under assembly: 8 rows x 16 bytes, + 16 so that the lanes' rows start in
// different banks
constexpr int K1_BLKB = 144;

template <int SLOTS, int LPW, bool SMEM, bool SUB>
__device__ __forceinline__ void k1_lane_loop(const K1Params& P, const ZpxIntervalDev& iv, const K1Start& st, const uint32_t sb,
                                             const uint32_t su, const uint32_t sdesc /* smem: this lane's scan's blk table */,
                                             const uint32_t tb /* smem: K1Tables */, const uint32_t ring_col,
                                             const uint32_t sbw /* smem: block of the warp's first lane */,
                                             const uint32_t siw /* smem: flush record of the warp's first lane */) {
    constexpr int RS = SLOTS * 4;
    const int lane = threadIdx.x & 31;
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];

    RingReader<RS> rd;
    if (st.count != 0) rd.init(ring_col, P.ublob, iv.ustart, iv.ulen, st.bitpos);
    else rd.init_idle(ring_col);
    Window<RS> win;
    win.load(rd);  // (top-ups never touch the ring words under the window: it stays valid from block to block)

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
    bool tail = SUB && st.k != 0 && st.count != 0;
    const int c_first = c;
    if (tail) c = c == 0 ? nblk - 1 : c - 1;
    // bi: x = DC table (SMEM: shared address of its LUT; else table index), y = AC likewise,
    //     z = comp | hx << 8 | vy << 16 | slot << 24, w = h | v << 8 | flags (| cache slots)
    uint4 bi = SMEM ? lds_u128(sdesc + c * 16) : bpack[c];
    const uint32_t* __restrict__ gdc = SMEM ? nullptr : P.huff[bi.x].fast;
    const uint32_t* __restrict__ gac = SMEM ? nullptr : P.huff[bi.y].fast;

    int dc0 = st.dc0, dc1 = st.dc1, dc2 = st.dc2, dc3 = st.dc3;
    uint32_t eob_run = 0;
    // (serial re-decode, zpx_api.cu rescue_eob_carry: the scan's first interval starts with the End-Of-Band run the scan
    // before it left open, decoder.zig:144)
    if (!SUB && P.eob_in != nullptr && iv.ordinal == 0 && st.count != 0) eob_run = k1_carry_in(P.eob_in + im->status_slot);
    int wide = 0;  // >= 13: some coefficient of the lane lies outside [-4096, 4095]
    const uint32_t total = st.count;
    uint32_t left = total;  // blocks still to decode (including the current one)

    // block flush (end of the loop body): this lane stores row (lane & 7) of the blocks of lanes 4 i + (lane >> 3)
    const uint32_t r16 = (uint32_t)(lane & 7) << 4;
    const uint32_t sbr = sbw + (uint32_t)(lane >> 3) * K1_BLKB, sir = siw + (uint32_t)(lane >> 3) * 8;

    while (__any_sync(0xffffffffu, left != 0)) {
        int k = 64;  // > 63: no block in flight on this lane
        int err = 0;
        rd.topup();  // the ring now holds what the DC symbol and K1_TOPUP AC symbols can take
        if (SUB && tail) {
            k = st.k;
        } else if (left != 0) {
            // ---- DC (decoder.zig:1366-1376) ----
            const uint32_t hi = win.peek(rd.bitpos);
            uint32_t e;
            if (SMEM) e = lds_u32(bi.x + ((hi >> (32 - K1_DLB)) << 2));
            else e = __ldg(gdc + (hi >> (32 - ZPX_LUT_BITS)));
            if ((int)e <= 0) e = k1_rare<SMEM>(P, hi, true, e, bi, tb, eob_run, err, wide);
            if (bi.w & 0x10000u) err = ZPX_E_UninitializedHuffmanTable;
            const int v = fe_extend(hi << fe_len(e), fe_s32(e));
            const int comp = (int)(bi.z & 0xff);
            int dc = comp == 0 ? dc0 : comp == 1 ? dc1 : comp == 2 ? dc2 : dc3;
            dc += v;
            if (comp == 0) dc0 = dc; else if (comp == 1) dc1 = dc; else if (comp == 2) dc2 = dc; else dc3 = dc;
            if (dc < -32768 || dc > 32767) report_coef_range(P.status, im->status_slot);
            wide = max(wide, ((dc ^ (dc >> 31)) >> 12) ? 13 : 0);
            win.advance(rd, rd.bitpos, (uint32_t)fe_tot(e));
            rd.bitpos += (uint32_t)fe_tot(e);
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
        }
        // ---- AC (decoder.zig:1383-1411): one symbol per lane per vote ----
        // K1_TOPUP steps, unrolled, between two top-ups.  The common code of a step is branch-free; lanes that idle
        // or meet a rare symbol see a null entry there, and the rare ones are dealt with at the end of the step.
        const bool nostore = SUB && tail;
        bool more = __any_sync(0xffffffffu, k <= 63);
        while (more) {
#pragma unroll
            for (int u = 0; u < K1_TOPUP; u++) {
                const uint32_t hi = win.peek(rd.bitpos);
                uint32_t e0;
                if (SMEM) e0 = lds_u32(bi.y + ((hi >> (32 - K1_ALB)) << 2));
                else e0 = __ldg(gac + (hi >> (32 - ZPX_LUT_BITS)));
                const bool act = k <= 63;
                const bool rare = act && (int)e0 <= 0;
#if ZPX_K1_ONE_SELECT
                // (one select on the chain between the table look-up and the entry's fields: "block open" is known early)
                uint32_t e;
                asm("{ .reg .pred q, p; setp.ne.s32 q, %2, 0; setp.gt.and.s32 p, %1, 0, q; selp.b32 %0, %1, %3, p; }"
                    : "=r"(e) : "r"(e0), "r"((int)act), "r"(K1_NULL_E));
#else
                const uint32_t e = (act && (int)e0 > 0) ? e0 : K1_NULL_E;
#endif
                {
                    const int len = fe_len(e), s32 = fe_s32(e);
                    int tot = fe_tot(e);
                    const int v = fe_extend(hi << len, s32);
                    const int kn = k + fe_adv(e);  // the value goes to zig-zag index kn - 1
                    // decoder.zig:1393-1395: a run past the block end leaves the value bits unread.  (Taking this select off
                    // the position's dependency chain -- advance by tot, step back in a rare branch -- measured 3.7 % slower.)
                    if (kn > 64 && s32 != 32) tot = len;
                    win.advance(rd, rd.bitpos, (uint32_t)tot);
                    rd.bitpos += (uint32_t)tot;
                    if (s32 != 32 && kn <= 64 && !nostore) sts_u16(sb + lds_u16(su + 2 * kn - 2), v);
                    k = kn;
                }
                if (rare) {
                    // Almost every rare symbol is an ordinary run/size symbol whose code is longer than the first-level
                    // table: canonical search over the cached limits (first match = the reference's rule,
                    // decoder.zig:946-969), inline.  Everything else (End-Of-Band runs, values of 13 bits or more,
                    // invalid codes, tables in HBM) takes the general path.
                    uint32_t e2 = 0;
                    if (SMEM && e0 == 0) e2 = k1_long_ac_sm(tb, (bi.w >> 24) & 15u, hi);
                    if (e2 == 0) e2 = k1_rare<SMEM>(P, hi, false, e0, bi, tb, eob_run, err, wide);
                    // End-Of-Band RUN inside a sequential scan: the synchronisation passes do not
                    // model that state
                    if (SUB && eob_run != 0) {
                        eob_run = 0;
                        if (!err) err = ZPX_E_UNSUPPORTED_STREAM;
                    }
                    const int len = fe_len(e2), s32 = fe_s32(e2);
                    int tot = fe_tot(e2);
                    const int v = fe_extend(hi << len, s32);
                    const int kn = k + fe_adv(e2);  // 64 after an error: the block ends here
                    if (kn > 64 && s32 != 32) tot = len;  // (an End-Of-Band run keeps its r run bits: no value bits)
                    win.advance(rd, rd.bitpos, (uint32_t)tot);
                    rd.bitpos += (uint32_t)tot;
                    if (s32 != 32 && kn <= 64 && !nostore) sts_u16(sb + lds_u16(su + 2 * kn - 2), v);
                    k = kn;
                }
#if ZPX_K1_VOTE_EVERY > 1
                // the warp votes on "any block still open" after every second step only (lanes whose block is complete idle
                // on null entries in between): -1.2 % entropy time; every fourth step: +0.7 %
                if ((u % ZPX_K1_VOTE_EVERY) == ZPX_K1_VOTE_EVERY - 1) {
                    more = __any_sync(0xffffffffu, k <= 63);
                    if (!more) break;
                }
#else
                more = __any_sync(0xffffffffu, k <= 63);
                if (!more) break;
#endif
            }
            if (more) rd.topup();
        }
        // ---- block end ----
        // flush record of this lane's block: x = destination address (low 32 bits), y = high 16 bits | key << 16 |
        // (store it) << 24 | (zero it) << 25; 0 = nothing to do
        uint32_t fx = 0, fy = 0;
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
            if (err || rd.overrun()) {
                // a symbol that needed bits past the limit is the reference's MissingFF00 / UnexpectedEof,
                // whatever the garbage decoded to
                if (rd.overrun()) err = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;
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
                const bool keep = !(bi.w & 0x40000u) && !(P.dbg & 1);  // not superseded by a later scan of the same component
                const unsigned long long dst = (unsigned long long)(P.coef + blk * 8);  // (device addresses fit 48 bits)
                fx = (uint32_t)dst;
                fy = (uint32_t)(dst >> 32) | (bx & 7u) << 16 | (keep ? 1u << 24 : 0u) | 1u << 25;
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
        // The warp stores its lanes' blocks together: eight lanes per block, one 16-byte row each, so that a store
        // instruction writes four whole 128-byte lines (a lane storing its own block row by row touches 32 lines
        // per instruction).  The rows are cleared for the next block on the way.
        if (__any_sync(0xffffffffu, fy != 0)) {
            if (lane < LPW) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(siw + lane * 8), "r"(fx), "r"(fy) : "memory");
            __syncwarp();
            // all records first, then all rows, then the stores: independent loads back to back instead of eight
            // dependent load -> load -> store chains
            uint32_t ix[LPW / 4], iy[LPW / 4];  // record of lane 4 i + (lane >> 3): x = address low, y = high | key << 16 | flags << 24
#pragma unroll
            for (int i = 0; i < LPW / 4; i++)
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ix[i]), "=r"(iy[i]) : "r"(sir + i * 32));
            uint4 row[LPW / 4];
            uint32_t ra[LPW / 4];
#pragma unroll
            for (int i = 0; i < LPW / 4; i++) {
                ra[i] = sbr + i * (4 * K1_BLKB) + ((r16 ^ (iy[i] >> 12)) & 0x70u);  // row (r ^ key) of that lane's block
                row[i] = lds_u128(ra[i]);  // (a valid shared address whatever the record holds)
            }
#pragma unroll
            for (int i = 0; i < LPW / 4; i++) {
                if (iy[i] & (1u << 24))
                    *reinterpret_cast<uint4*>(((unsigned long long)(iy[i] & 0xffffu) << 32 | ix[i]) + r16) = row[i];
                if (iy[i] >> 24) sts_zero16(ra[i]);
            }
            __syncwarp();
        }
    }
    // The reference keeps its End-Of-Band run across scans (decoder.zig:144, reset only at RSTn :1451); here
    // every scan starts from zero, so a run that is still open when a scan ends (corrupt streams only) would
    // make the next scan differ: refuse the image instead.
    if (wide >= 13) atomicOr(&P.img_flags[im->status_slot], 1u);  // a coefficient outside [-4096, 4095]
    if (!SUB && st.count != 0 && (iv.flags & 2u)) {
        if (P.eob_out != nullptr) {
            if (left == 0) P.eob_out[im->status_slot] = eob_run;  // (a lane that stopped at an error hands nothing on)
        } else if (eob_run != 0) {
            report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + iv.n_blocks, ZPX_E_UNSUPPORTED_STREAM);
        }
    }
}

// Dynamic shared memory of the write kernels
template <int SLOTS>
struct K1WriteSmem {
    uint32_t ring[K1_RW * SLOTS];  // per-lane stream rings: [word slot][lane]
    uint4 sblk[(K1_BLKB / 16) * SLOTS];  // per-lane block under assembly: [lanes][8 rows + 1] x 16 bytes
    uint2 sinfo[SLOTS];            // per-lane flush record (see the block end of k1_lane_loop)
    K1Tables tab;
};

// One CTA: WARPS warps whose first LPW lanes each carry a stream ("slot" = warp * LPW + lane): an interval and a
// start inside it (K1Start).  Collects the distinct scans and Huffman tables of the CTA's streams, stages their
// first-level LUTs and block descriptors in shared memory and runs the block-synchronous loop.
// Why fewer than 32 streams per warp: the loop is a chain of dependent instructions (a warp issues one every 4-5
// cycles), so its throughput comes from the number of resident warps.  A batch with fewer streams than
// 32 x resident warps is spread over narrower warps: the same work per stream, more warps to hide the latency.
template <int LPW, int WARPS, bool SUB>
__device__ __forceinline__ void k1_cta_run(const K1Params& P, const ZpxIntervalDev& iv, const K1Start& st) {
    constexpr int SLOTS = WARPS * LPW;
    extern __shared__ __align__(16) uint8_t k1_smem[];
    K1WriteSmem<SLOTS>& S = *reinterpret_cast<K1WriteSmem<SLOTS>*>(k1_smem);
    const int lane = threadIdx.x & 31;
    const bool owner = lane < LPW;
    // lanes without a stream run the same straight-line code on a neighbour's (never written) addresses
    const int slot = (threadIdx.x >> 5) * LPW + (owner ? lane : lane - LPW);
    if (owner) {
        for (int r = 0; r < K1_BLKB / 16; r++) S.sblk[slot * (K1_BLKB / 16) + r] = make_uint4(0, 0, 0, 0);
        S.tab.lane_scan[slot] = iv.scan;
    }
    __syncthreads();
    const bool cached = k1_tables_setup(P, S.tab, SLOTS);
    uint32_t sdesc = 0;
    if (cached)
        for (int i = 0; i < S.tab.nscan; i++)
            if (S.tab.scan_id[i] == iv.scan) sdesc = smem_addr(&S.tab.desc[i][0]);
    uint32_t sb = smem_addr(S.sblk) + slot * K1_BLKB;  // this lane's row 0
    const int wslot0 = (threadIdx.x >> 5) * LPW;
    uint32_t sbw = smem_addr(S.sblk) + wslot0 * K1_BLKB, siw = smem_addr(S.sinfo) + wslot0 * 8;
    uint32_t su = smem_addr(S.tab.unzig);
    uint32_t tb = smem_addr(&S.tab);
    uint32_t rc = smem_addr(S.ring) + slot * 4;
    // opaque copies: keeps the shared-window address arithmetic out of the symbol loop
    asm volatile("mov.u32 %0, %0;" : "+r"(sb));
    asm volatile("mov.u32 %0, %0;" : "+r"(su));
    asm volatile("mov.u32 %0, %0;" : "+r"(tb));
    asm volatile("mov.u32 %0, %0;" : "+r"(rc));
    asm volatile("mov.u32 %0, %0;" : "+r"(sbw));
    asm volatile("mov.u32 %0, %0;" : "+r"(siw));
    if (cached) k1_lane_loop<SLOTS, LPW, true, SUB>(P, iv, st, sb, su, sdesc, tb, rc, sbw, siw);
    else k1_lane_loop<SLOTS, LPW, false, SUB>(P, iv, st, sb, su, 0, tb, rc, sbw, siw);
}

// K1a: one lane per restart interval, LPW intervals per warp.
template <int LPW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS == 8 ? (LPW <= 20 ? 4 : 3) : 4)) k1_lane_per_interval(const K1Params P) {
    const int lane = threadIdx.x & 31;
    // lanes past LPW shadow the stream of lane - LPW (same interval, so every table address they form is valid)
    // and lanes past the end of the interval list the last interval; both idle through the loop
    const int gid = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * LPW + (lane < LPW ? lane : lane - LPW);
    const bool live = lane < LPW && gid < P.n_iv;
    const ZpxIntervalDev iv = P.ivs[gid < P.n_iv ? gid : P.n_iv - 1];
    K1Start st;
    st.bitpos = 0;
    st.k = 0;
    st.dc0 = st.dc1 = st.dc2 = st.dc3 = 0;
    st.j0 = 0;
    st.count = live ? iv.n_blocks : 0;
    k1_cta_run<LPW, WARPS, false>(P, iv, st);
}

// K1b, last pass (zpx_k1s.cu): one lane per sub-sequence, 32 consecutive sub-sequences of one segment per warp.
// Every lane starts from its true state (found by k1s_sync), skips the tail of a block begun in the previous
// sub-sequence and writes the blocks that START inside its own (their number and the DC predictors at that
// point come from k1s_scan), running past its boundary to finish the last one.
constexpr int K1S_WWARPS = 4;  // warps per CTA of k1s_write
__global__ void __launch_bounds__(K1S_WWARPS * 32, 4) k1s_write(const K1SParams P) {
    const int wid = blockIdx.x * K1S_WWARPS + (threadIdx.x >> 5);
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
    k1_cta_run<32, K1S_WWARPS, true>(P.k1, iv, st);
}

// Streams per warp for a batch of n intervals: the narrowest warps whose single wave still holds the whole batch
// (capacity = SMs x resident CTAs x warps x LPW); 32 when no width does (several waves: full warps do the most
// work per issued instruction).  ZPX_K1_LPW / ZPX_K1_WARPS override (experiments).
static int k1_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int LPW, int WARPS>
static cudaError_t k1_launch_lpw(const K1Params& P, cudaStream_t s) {
    constexpr int per_cta = WARPS * LPW;
    const size_t smem = sizeof(K1WriteSmem<per_cta>);
    cudaError_t e = cudaFuncSetAttribute(k1_lane_per_interval<LPW, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k1_lane_per_interval<LPW, WARPS><<<(P.n_iv + per_cta - 1) / per_cta, WARPS * 32, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t k1_launch_lane_per_interval(const K1Params& P, int sm_count, cudaStream_t s) {
    if (P.n_iv <= 0) return cudaSuccess;
    (void)sm_count;
    const int lpw = k1_env("ZPX_K1_LPW", 32), warps = k1_env("ZPX_K1_WARPS", 4);
    if (lpw == 16) return warps == 8 ? k1_launch_lpw<16, 8>(P, s) : k1_launch_lpw<16, 4>(P, s);
    return warps == 8 ? k1_launch_lpw<32, 8>(P, s) : k1_launch_lpw<32, 4>(P, s);
}

cudaError_t k1s_launch_write(const K1SParams& P, cudaStream_t s) {
    if (P.n_warps <= 0) return cudaSuccess;
    const size_t smem = sizeof(K1WriteSmem<K1S_WWARPS * 32>);
    cudaError_t e = cudaFuncSetAttribute(k1s_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k1s_write<<<(P.n_warps + K1S_WWARPS - 1) / K1S_WWARPS, K1S_WWARPS * 32, smem, s>>>(P);
    return cudaGetLastError();
}

}  // namespace zpx
