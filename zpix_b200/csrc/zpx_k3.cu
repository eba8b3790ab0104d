// zpx_k3.cu -- progressive (SOF2) scans on the GPU.
//
// Replaces the progressive branches of processSos (src/jpeg/decoder.zig:1268-1283, 1340-1425),
// refine (:1459-1518) and refineNonZeroes (:1522-1549).  The coefficients of a progressive frame
// live in per-component planar grids of int16 blocks in HBM (the reference's
// progressive_coefficients, :1341/:1415), zeroed before the first scan; every scan is one launch
// that reads-modifies-writes them.  Scans of one image are launched by dependency level (a scan waits
// for the earlier scans that touch the same component and an overlapping band; zpx_api.cu); scans of
// all images with the same level share a launch.  After the last scan the unfused k2g kernels
// reconstruct only the blocks that intersect the image (:1636-1661).
//
// Spectral selection / successive approximation make the bits a block consumes depend on the
// coefficients earlier scans left in it, so an interval is decoded serially; the unit of parallelism
// is one WARP per restart interval (per scan when DRI = 0).  Lane 0 runs the serial symbol loop out of
// shared memory and registers only; the 32 lanes together keep it fed and do everything that is not
// serial:
//   * stream: 512 raw bytes per refill, loaded coalesced one refill ahead (registers), FF 00 pairs
//     removed in parallel (a 0x00 after a 0xFF is dropped: inside an interval every 0xFF is followed
//     by 0x00, the host's limit search guarantees it), bytes compacted into a 1 KB ring of big-endian
//     words.  The reader is a bit position plus the two ring words under it;
//   * Huffman first-level tables (9 bits, 512 x u16) are copied to shared memory per warp;
//   * AC refinement (the only pass that reads coefficients): batches of 8 blocks are loaded one batch
//     ahead and transposed to zig-zag order in shared memory.  For each block the warp builds the list
//     of zero positions and the running count of non-zero ones, so that lane 0 only decodes the Huffman
//     symbols: the target of a run is a table lookup, and the correction bits in between are *skipped*
//     by count (their number is known from the non-zero mask).  Afterwards all lanes apply the
//     correction bits in parallel: coefficient z finds its bit at
//         block start + bits of the symbols read before z was passed + non-zero coefficients before z.
//   * all other passes only write (or OR one bit): stores straight to HBM.
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_entropy.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

namespace zpx {

namespace {

constexpr int K3_WARPS = 4;
constexpr int K3_NT = 32 * K3_WARPS;
constexpr int K3_RING = 1024;        // bytes, power of two
constexpr int K3_RW = K3_RING / 4;
constexpr int K3_CHUNK = 512;        // raw bytes per refill
constexpr uint32_t K3_LOW = 320 * 8; // lane 0 starts a block only with this many bits buffered (a block
                                     // consumes at most 63 x 32 bits)
constexpr int K3_BATCH = 8;          // AC refinement: blocks per shared-memory batch
constexpr int K3_LB = 9;             // first-level table bits in shared memory
constexpr int K3_LS = 1 << K3_LB;

struct __align__(16) BlkPos {  // one block position of the scan's MCU
    uint32_t base;             // component grid origin (block units)
    uint16_t bw;               // grid width in blocks
    uint8_t comp, hx, vy, h, v;
    uint8_t flags;             // bit0 DC table undefined, bit1 AC table undefined
    uint16_t tdc, tac;         // device table indices
};
static_assert(sizeof(BlkPos) == 16, "BlkPos is one 16-byte shared-memory load");

struct WarpSm {
    uint32_t ring[K3_RW];
    uint16_t lut[4][K3_LS];             // first-level tables, slot = frame component: sym << 8 | len, 0 = longer
    short zz[K3_BATCH][64];             // refinement batch, zig-zag order (zero outside the scan's band)
    uint16_t cum[K3_BATCH][64];         // [z]: bits of the symbols read so far, written when a symbol starts at z
    uint8_t zl[K3_BATCH][80];           // zero positions of the band, ascending, 0xff-terminated
    uint8_t ncnt[K3_BATCH][64];         // [z]: non-zero coefficients of the band below z
    uint32_t nzlo[K3_BATCH], nzhi[K3_BATCH];  // non-zero masks (before this scan)
    uint32_t p0[K3_BATCH];              // bit position of the block's first bit
    BlkPos pos[ZPX_MAX_BLK_PER_MCU];
};

// stored slot of natural coefficient `nat` in a block whose rows are XOR-swizzled by key
__device__ __forceinline__ int cslot(int key, int nat) { return (((nat >> 3) ^ key) << 3) + (nat & 7); }

// bit reader over the de-stuffed ring: a bit position and the two ring words under it
struct Rd {
    const uint32_t* ring;
    uint32_t pos, wi, w0, w1;
    __device__ __forceinline__ void init(const uint32_t* r, uint32_t p) {
        ring = r;
        pos = p;
        wi = p >> 5;
        w0 = ring[wi & (K3_RW - 1)];
        w1 = ring[(wi + 1) & (K3_RW - 1)];
    }
    __device__ __forceinline__ void seek(uint32_t p) {
        pos = p;
        const uint32_t i = p >> 5;
        if (i != wi) {
            wi = i;
            w0 = ring[i & (K3_RW - 1)];
            w1 = ring[(i + 1) & (K3_RW - 1)];
        }
    }
    __device__ __forceinline__ void skip(uint32_t n) { seek(pos + n); }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(w1, w0, pos); }
};

// decodeHuffman (decoder.zig:909-970): first level from shared memory, longer codes from the table in HBM.
// Returns sym | len << 8; len == 0: no code matches.
__device__ __noinline__ uint32_t huff_long(const ZpxHuffDev* __restrict__ t, uint32_t hi) {
    const uint32_t v16 = hi >> 16;
    for (int l = K3_LB + 1; l <= 16; l++) {
        if (v16 < __ldg(&t->limit[l])) {
            const uint32_t sym = __ldg(&t->vals[(__ldg(&t->valoff[l]) + (int)(v16 >> (16 - l))) & 0xff]);
            return sym | (uint32_t)l << 8;
        }
    }
    return 0;
}
__device__ __forceinline__ uint32_t huff_sm(const uint16_t* lut, const ZpxHuffDev* __restrict__ t, uint32_t hi) {
    const uint32_t e = lut[hi >> (32 - K3_LB)];  // sym << 8 | len
    if ((e & 0xffu) != 0) return __byte_perm(e, 0, 0x4401);
    return huff_long(t, hi);
}
// table kept in HBM (the kind this scan does not cache)
__device__ __noinline__ uint32_t huff_global(const ZpxHuffDev* __restrict__ t, uint32_t hi) {
    const HuffSym q = huff_decode(t, hi);
    return q.sym | (uint32_t)q.len << 8;
}

// RECEIVE + EXTEND (decoder.zig:1115-1134) on the `size` bits after the first `len` bits of hi; len + size <= 32
__device__ __forceinline__ int extend32(uint32_t hi, int len, int size) {
    if (size == 0) return 0;
    const uint32_t t = hi << len;
    const int v = (int)(t >> (32 - size));
    return (t >> 31) ? v : v + ((-1) << size) + 1;
}

}  // namespace

// (see k1_carry_in, zpx_k1.cu: an inlined load here would share a scoreboard with the stream loads)
static __device__ __noinline__ uint32_t k3_carry_in(const uint32_t* p) { return *p; }

__global__ void __launch_bounds__(K3_NT, 5) k3_progressive(const K1Params P, const uint32_t* __restrict__ list, const int n_list) {
    __shared__ __align__(16) WarpSm s_w[K3_WARPS];
    __shared__ uint8_t s_unzig[64];
    if (threadIdx.x < 64) s_unzig[threadIdx.x] = c_unzig[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = blockIdx.x * K3_WARPS + warp;
    if (gid >= n_list) return;  // whole warps
    WarpSm& W = s_w[warp];
    const ZpxIntervalDev iv = P.ivs[list[gid]];
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const int err_eof = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;
    const int ss = sc->ss, se = sc->se, ah = sc->ah, al = sc->al;
    const bool interleaved = sc->interleaved != 0;
    const int nblk = interleaved ? sc->nblk : 1;
    const uint32_t mxx = (uint32_t)im->mxx, cw = (uint32_t)sc->cw;
    const bool acref = ah != 0 && ss != 0;   // the only pass that reads coefficients
    const bool lut_is_dc = ss == 0;          // which kind of table sits in shared memory
    const uint32_t lt = (1u << lane) - 1;

    // ---- per-warp tables ----
    if (lane < nblk) {
        const int comp = sc->blk_comp[lane];
        BlkPos b;
        b.base = (uint32_t)im->comp_base[comp];
        b.bw = (uint16_t)im->comp_bw[comp];
        b.comp = (uint8_t)comp;
        b.hx = sc->blk_hx[lane];
        b.vy = sc->blk_vy[lane];
        b.h = im->h[comp];
        b.v = im->v[comp];
        b.flags = (uint8_t)((sc->blk_pack[lane][3] >> 16) & 3u);
        b.tdc = sc->blk_dc[lane];
        b.tac = sc->blk_ac[lane];
        W.pos[lane] = b;
    }
    __syncwarp();
    {
        uint32_t loaded = 0;
        for (int c = 0; c < nblk; c++) {
            const BlkPos b = W.pos[c];
            if (loaded >> b.comp & 1u) continue;
            loaded |= 1u << b.comp;
            const ZpxHuffDev* __restrict__ t = &P.huff[lut_is_dc ? b.tdc : b.tac];
            // 9-bit table from the 10-bit one: a code of at most 9 bits fills both entries 2i and 2i + 1
            for (int i = lane; i < K3_LS; i += 32) {
                const uint32_t e = __ldg(&t->lut[2 * i]);
                W.lut[b.comp][i] = (e & 0xffu) <= (uint32_t)K3_LB ? (uint16_t)e : (uint16_t)0;
            }
        }
    }

    // ---- stream state (warp-uniform) ----
    const uint64_t a0 = iv.start & ~(uint64_t)15;
    const uint4* __restrict__ raw = reinterpret_cast<const uint4*>(P.blob + a0);
    const uint32_t first = (uint32_t)(iv.start - a0), end = first + iv.len;  // byte offsets relative to a0
    uint32_t chunk = 0;      // byte offset (relative to a0) of the chunk held in registers
    uint32_t wpos = 0;       // de-stuffed bytes written to the ring (zero padding after the data included)
    uint32_t dbits = 0;      // data bits among them
    uint32_t last_raw = 0;   // last raw byte of the previous chunk
    uint4 creg = make_uint4(0, 0, 0, 0);
    {
        const uint32_t o = chunk + 16 * lane;
        if (o < end) creg = __ldg(raw + (o >> 4));
    }

    // ---- block / decode state ----
    uint32_t j = 0;            // blocks done (uniform after each round)
    uint32_t bitpos = 0;       // reader position (uniform after each round)
    int err = 0;
    // lane 0 only:
    int dc0 = 0, dc1 = 0, dc2 = 0, dc3 = 0;
    uint32_t eob_run = 0;
    // (serial re-decode: the reference keeps the run across scans, decoder.zig:144 -- the scan's first interval starts
    // with what the previous scan left; a restart marker resets it, :1451)
    if (P.eob_in != nullptr && iv.ordinal == 0) eob_run = k3_carry_in(P.eob_in + im->status_slot);  // (out of line: zpx_k1.cu)
    // refinement batch
    uint32_t bj0 = 0, bn = 0, bk = 0;   // batch = blocks [bj0, bj0+bn), bk done
    short pre[2 * K3_BATCH];             // prefetched coefficients of the next batch: zig-zag lane, lane+32
    uint32_t pj0 = 0, pn = 0;            // the prefetched batch
    short* const cbase = reinterpret_cast<short*>(P.coef);
    const int uz0 = s_unzig[lane], uz1 = s_unzig[lane + 32];
    // Only the scan's own band [ss, se] is read and written back: scans of the same level may be working
    // on other bands of the same blocks at the same time.
    const bool inb0 = lane >= ss && lane <= se, inb1 = lane + 32 >= ss && lane + 32 <= se;

    // address (in shorts) and swizzle key of block ordinal jb of this interval
    auto block_addr = [&](uint32_t jb, int& key) -> uint64_t {
        int bx, by, c;
        if (interleaved) {
            const uint32_t m = iv.first_mcu + jb / (uint32_t)nblk;
            c = (int)(jb % (uint32_t)nblk);
            const uint32_t my = m / mxx, mx = m - my * mxx;
            const BlkPos b = W.pos[c];
            bx = b.h * (int)mx + b.hx;
            by = b.v * (int)my + b.vy;
        } else {
            c = 0;
            const uint32_t q = iv.first_block + jb;
            by = (int)(q / cw);
            bx = (int)(q - (uint32_t)by * cw);
        }
        const BlkPos b = W.pos[c];
        key = bx & 7;
        return ((uint64_t)b.base + (uint64_t)by * b.bw + (uint64_t)bx) * 64;
    };
    // the same for the k-th block after (bx0, by0) of a non-interleaved scan: blocks follow each other along the
    // component's row (one division per batch instead of one per block)
    auto block_addr_seq = [&](uint32_t bx0, uint32_t by0, uint32_t k, int& key) -> uint64_t {
        uint32_t bx = bx0 + k, by = by0;
        while (bx >= cw) {
            bx -= cw;
            by++;
        }
        const BlkPos b = W.pos[0];
        key = (int)(bx & 7u);
        return ((uint64_t)b.base + (uint64_t)by * b.bw + (uint64_t)bx) * 64;
    };
    auto prefetch = [&](uint32_t j0) {
        pj0 = j0;
        pn = j0 < iv.n_blocks ? min((uint32_t)K3_BATCH, iv.n_blocks - j0) : 0;
        const uint32_t q0 = iv.first_block + j0, by0 = interleaved ? 0 : q0 / cw, bx0 = interleaved ? 0 : q0 - by0 * cw;
#pragma unroll
        for (int k = 0; k < K3_BATCH; k++) {
            if ((uint32_t)k < pn) {
                int key;
                const short* g = cbase + (interleaved ? block_addr(j0 + k, key) : block_addr_seq(bx0, by0, (uint32_t)k, key));
                pre[2 * k] = inb0 ? g[cslot(key, uz0)] : (short)0;
                pre[2 * k + 1] = inb1 ? g[cslot(key, uz1)] : (short)0;
            }
        }
    };
    auto store_batch = [&]() {
        const uint32_t q0 = iv.first_block + bj0, by0 = interleaved ? 0 : q0 / cw, bx0 = interleaved ? 0 : q0 - by0 * cw;
        for (uint32_t k = 0; k < bn; k++) {
            int key;
            short* g = cbase + (interleaved ? block_addr(bj0 + k, key) : block_addr_seq(bx0, by0, k, key));
            if (inb0) g[cslot(key, uz0)] = W.zz[k][lane];
            if (inb1) g[cslot(key, uz1)] = W.zz[k][lane + 32];
        }
    };
    if (acref) prefetch(0);

    for (;;) {
        // ---- A. refill the ring (all lanes) ----
        if (wpos - (bitpos >> 3) <= (uint32_t)(K3_RING - K3_CHUNK - 8)) {
            const uint32_t o0 = chunk + 16 * lane;
            const uint32_t wd[4] = {creg.x, creg.y, creg.z, creg.w};
            uint32_t prev = __shfl_up_sync(0xffffffffu, creg.w >> 24, 1);
            if (lane == 0) prev = last_raw;
            last_raw = __shfl_sync(0xffffffffu, creg.w >> 24, 31);
            uint32_t keep = 0, nreal = 0;
            uint8_t by[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t o = o0 + k;
                uint32_t b = (wd[k >> 2] >> (8 * (k & 3))) & 0xffu;
                bool kp;
                if (o >= end) {         // past the limit: zero padding
                    b = 0;
                    kp = true;
                } else if (o < first) {
                    kp = false;
                } else {
                    kp = !(b == 0 && prev == 0xffu && o > first);
                    nreal += kp;
                }
                by[k] = (uint8_t)b;
                keep |= (uint32_t)kp << k;
                prev = (wd[k >> 2] >> (8 * (k & 3))) & 0xffu;
            }
            const uint32_t v = (uint32_t)__popc(keep) | nreal << 16;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t p = wpos + ((incl - v) & 0xffffu);
            uint8_t* rb = reinterpret_cast<uint8_t*>(W.ring);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (keep >> k & 1u) {
                    rb[(p & (K3_RING - 1)) ^ 3u] = by[k];
                    p++;
                }
            }
            wpos += tot & 0xffffu;
            dbits += (tot >> 16) * 8;
            chunk += K3_CHUNK;
            creg = make_uint4(0, 0, 0, 0);
            const uint32_t o = chunk + 16 * lane;
            if (o < end) creg = __ldg(raw + (o >> 4));
        }
        // ---- B. refinement batch turnover (all lanes) ----
        if (acref && bk == bn) {
            if (bn) store_batch();
            bj0 = pj0;
            bn = pn;
            bk = 0;
            const uint32_t band_lo = __ballot_sync(0xffffffffu, inb0), band_hi = __ballot_sync(0xffffffffu, inb1);
#pragma unroll
            for (int k = 0; k < K3_BATCH; k++) {
                if ((uint32_t)k < bn) {
                    W.zz[k][lane] = pre[2 * k];
                    W.zz[k][lane + 32] = pre[2 * k + 1];
                    const uint32_t nlo = __ballot_sync(0xffffffffu, pre[2 * k] != 0);
                    const uint32_t nhi = __ballot_sync(0xffffffffu, pre[2 * k + 1] != 0);
                    const uint32_t zlo = band_lo & ~nlo, zhi = band_hi & ~nhi;
                    const int nzl = __popc(nlo), nzero_lo = __popc(zlo), nzero = nzero_lo + __popc(zhi);
                    W.ncnt[k][lane] = (uint8_t)__popc(nlo & lt);
                    W.ncnt[k][lane + 32] = (uint8_t)(nzl + __popc(nhi & lt));
                    if (zlo >> lane & 1u) W.zl[k][__popc(zlo & lt)] = (uint8_t)lane;
                    if (zhi >> lane & 1u) W.zl[k][nzero_lo + __popc(zhi & lt)] = (uint8_t)(lane + 32);
                    if (lane < 16) W.zl[k][nzero + lane] = 0xff;  // zl[zi + r] reads at most 15 past the last zero
                    W.cum[k][lane] = 0;
                    W.cum[k][lane + 32] = 0;
                    if (lane == 0) {
                        W.nzlo[k] = nlo;
                        W.nzhi[k] = nhi;
                    }
                }
            }
            prefetch(bj0 + bn);
        }
        __syncwarp();

        // ---- C. lane 0 decodes until the ring runs low, the batch ends or the interval is done ----
        const uint32_t bk_before = bk;
        if (lane == 0) {
            const uint32_t wbits = wpos * 8;
            if (acref) {
                // ---- AC successive-approximation refinement (decoder.zig:1468-1517) ----
                Rd rd;
                rd.init(W.ring, bitpos);
                const int delta = 1 << al;
                const BlkPos bp = W.pos[0];
                const uint16_t* lut = W.lut[bp.comp];
                const ZpxHuffDev* __restrict__ tac = &P.huff[bp.tac];
                const bool ac_undef = (bp.flags & 2) != 0;
                uint32_t pos = bitpos;
                while (bk < bn && wbits - pos >= K3_LOW) {
                    const uint32_t p0 = pos;
                    const uint8_t* zl = W.zl[bk];
                    const uint8_t* ncnt = W.ncnt[bk];
                    uint16_t* cum = W.cum[bk];
                    short* zz = W.zz[bk];
                    const uint32_t nzc = (uint32_t)(__popc(W.nzlo[bk]) + __popc(W.nzhi[bk]));
                    W.p0[bk] = p0;
                    uint32_t sbits = 0;
                    int zig = ss, zi = 0;
                    if (eob_run == 0) {
                        while (zig <= se) {
                            if (ac_undef) { err = ZPX_E_UninitializedHuffmanTable; break; }
                            pos = p0 + sbits + ncnt[zig];
                            rd.seek(pos);
                            const uint32_t hi = rd.peek();
                            const uint32_t hs = huff_sm(lut, tac, hi);
                            const int len = (int)(hs >> 8);
                            if (len == 0) { pos += 16; err = ZPX_E_BadHuffmanCode; break; }
                            const int r = (int)(hs >> 4) & 15, s = (int)hs & 15;
                            int z = 0;
                            if (s == 0) {
                                if (r != 15) {  // EOBn: the rest of the band of this and the next eob_run-1 blocks
                                    eob_run = 1u << r;
                                    if (r != 0) eob_run |= (hi << len) >> (32 - r);
                                    sbits += len + r;
                                    cum[zig] = (uint16_t)sbits;
                                    break;
                                }
                                sbits += len;
                            } else if (s == 1) {
                                z = ((hi << len) >> 31) ? delta : -delta;
                                sbits += len + 1;
                            } else {
                                pos += len;
                                err = ZPX_E_UnexpectedHuffmanCode;
                                break;
                            }
                            cum[zig] = (uint16_t)sbits;
                            // the (r+1)-th zero coefficient at or after zig; the non-zero ones passed on the way
                            // take one correction bit each (refineNonZeroes :1522-1549): skipped here by count
                            const int t = zl[zi + r];
                            zi += r + 1;
                            if (t > se) {
                                pos = p0 + sbits + nzc;
                                err = ZPX_E_TooManyCoefficients;
                                break;
                            }
                            if (z != 0) zz[t] = (short)z;
                            zig = t + 1;
                        }
                    }
                    if (err) {
                        if (pos > dbits) err = err_eof;
                        break;
                    }
                    if (eob_run > 0) eob_run--;
                    pos = p0 + sbits + nzc;
                    if (pos > dbits) { err = err_eof; break; }
                    bk++;
                    j++;
                }
                bitpos = pos;
            } else {
                Rd rd;
                rd.init(W.ring, bitpos);
                const uint32_t jlim = iv.n_blocks;
                // position of block j
                uint32_t c = 0, mx = 0, my = 0, bxn = 0, byn = 0;
                if (interleaved) {
                    const uint32_t m = iv.first_mcu + j / (uint32_t)nblk;
                    c = j % (uint32_t)nblk;
                    my = m / mxx;
                    mx = m - my * mxx;
                } else {
                    const uint32_t q = iv.first_block + j;
                    byn = q / cw;
                    bxn = q - byn * cw;
                }
                while (j < jlim && wbits - rd.pos >= K3_LOW) {
                    const BlkPos bp = W.pos[c];
                    const int comp = bp.comp;
                    int bx, by;
                    if (interleaved) {
                        bx = bp.h * (int)mx + bp.hx;
                        by = bp.v * (int)my + bp.vy;
                    } else {
                        bx = (int)bxn;
                        by = (int)byn;
                    }
                    const int key = bx & 7;
                    short* gblk = cbase + ((uint64_t)(bp.base + (uint32_t)by * bp.bw + (uint32_t)bx) << 6);
                    const uint16_t* lut = W.lut[comp];
                    uint32_t nskip = 1;  // blocks this step covers

                    if (ah != 0) {
                        // ---- DC refinement (decoder.zig:1462-1467): b[0] |= bit << al, as a one-way atomic OR
                        // on the word that holds it ----
                        const uint32_t hi = rd.peek();
                        rd.skip(1);
                        if (hi >> 31) {
                            short* p = gblk + cslot(key, 0);
                            const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                            atomicOr(reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3), (unsigned int)((1 << al) & 0xffff) << ((a & 2) ? 16 : 0));
                        }
                    } else {
                        // ---- first pass of a band (decoder.zig:1362-1411) ----
                        int zig = ss;
                        bool stop = false;
                        if (zig == 0) {
                            zig++;
                            if (bp.flags & 1) { err = ZPX_E_UninitializedHuffmanTable; stop = true; }
                            if (!stop) {
                                const ZpxHuffDev* __restrict__ tdc = &P.huff[bp.tdc];
                                const uint32_t hi = rd.peek();
                                const uint32_t hs = huff_sm(lut, tdc, hi);  // ss == 0: the DC tables are the cached ones
                                const int len = (int)(hs >> 8), t = (int)(hs & 0xff);
                                if (len == 0) { rd.pos += 16; err = ZPX_E_BadHuffmanCode; stop = true; }
                                else if (t > 16) { rd.pos += len; err = ZPX_E_ExcessiveDCComponent; stop = true; }
                                else {
                                    int d = comp == 0 ? dc0 : comp == 1 ? dc1 : comp == 2 ? dc2 : dc3;
                                    d += extend32(hi, len, t);
                                    if (comp == 0) dc0 = d; else if (comp == 1) dc1 = d; else if (comp == 2) dc2 = d; else dc3 = d;
                                    rd.skip(len + t);
                                    const int v = (int)((uint32_t)d << al);
                                    // hard error here, unlike the sequential kernels: later scans parse according to
                                    // which coefficients are non-zero, so nothing after a wrapped value can be trusted
                                    if (v < -32768 || v > 32767) { err = ZPX_E_COEF_RANGE; stop = true; }
                                    else gblk[cslot(key, 0)] = (short)v;
                                }
                            }
                        }
                        if (!stop && zig <= se) {
                            if (eob_run > 0) {
                                // End-Of-Band run: this block and the following eob_run - 1 have nothing in the band
                                if (ss > 0 && !interleaved) {
                                    nskip = min(eob_run, jlim - j);
                                    eob_run -= nskip;
                                } else {
                                    eob_run--;
                                }
                            } else {
                                const ZpxHuffDev* __restrict__ tac = &P.huff[bp.tac];
                                const bool ac_undef = (bp.flags & 2) != 0;
                                while (zig <= se) {
                                    if (ac_undef) { err = ZPX_E_UninitializedHuffmanTable; break; }
                                    const uint32_t hi = rd.peek();
                                    // AC symbols inside a scan that starts at DC: that table stays in HBM
                                    const uint32_t hs = lut_is_dc ? huff_global(tac, hi) : huff_sm(lut, tac, hi);
                                    const int len = (int)(hs >> 8);
                                    if (len == 0) { rd.pos += 16; err = ZPX_E_BadHuffmanCode; break; }
                                    const int r = (int)(hs >> 4) & 15, s = (int)hs & 15;
                                    if (s != 0) {
                                        zig += r;
                                        if (zig > se) { rd.skip(len); break; }  // decoding goes on: keep the reader's words in step
                                        const int v = (int)((uint32_t)extend32(hi, len, s) << al);
                                        rd.skip(len + s);
                                        if (v < -32768 || v > 32767) { err = ZPX_E_COEF_RANGE; break; }
                                        gblk[cslot(key, s_unzig[zig])] = (short)v;
                                    } else {
                                        if (r != 15) {
                                            eob_run = 1u << r;
                                            if (r != 0) eob_run |= (hi << len) >> (32 - r);
                                            eob_run = (eob_run - 1) & 0xffffu;
                                            rd.skip(len + r);
                                            break;
                                        }
                                        rd.skip(len);
                                        zig += 15;
                                    }
                                    zig++;
                                }
                            }
                        }
                    }
                    if (rd.pos > dbits) err = err_eof;   // a symbol needed bits the interval does not have
                    if (err) break;

                    j += nskip;
                    if (interleaved) {
                        if (++c == (uint32_t)nblk) {
                            c = 0;
                            if (++mx == mxx) { mx = 0; my++; }
                        }
                    } else {
                        bxn += nskip;
                        if (bxn >= cw) {
                            byn += bxn / cw;
                            bxn %= cw;
                        }
                    }
                }
                bitpos = rd.pos;
            }
        }
        __syncwarp();
        j = __shfl_sync(0xffffffffu, j, 0);
        bitpos = __shfl_sync(0xffffffffu, bitpos, 0);
        bk = __shfl_sync(0xffffffffu, bk, 0);
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) break;
        // ---- D. correction bits of the blocks lane 0 just parsed (all lanes) ----
        if (acref) {
            const int delta = 1 << al;
            for (uint32_t k = bk_before; k < bk; k++) {
                const uint32_t nlo = W.nzlo[k], nhi = W.nzhi[k], p0 = W.p0[k];
                // bits of the symbols read before position z was passed = running maximum of cum[0..z]
                uint32_t c0 = W.cum[k][lane], c1 = W.cum[k][lane + 32];
                if (__any_sync(0xffffffffu, (c0 | c1) != 0)) {  // (a block inside an End-Of-Band run has no symbols)
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        c0 = max(c0, __shfl_up_sync(0xffffffffu, c0, d));   // lanes below d read themselves: no-op
                        c1 = max(c1, __shfl_up_sync(0xffffffffu, c1, d));
                    }
                    c1 = max(c1, __shfl_sync(0xffffffffu, c0, 31));
                }
                if (nlo >> lane & 1u) {
                    const uint32_t bp = p0 + c0 + (uint32_t)__popc(nlo & lt);
                    if ((W.ring[(bp >> 5) & (K3_RW - 1)] << (bp & 31)) >> 31) {
                        const int v = W.zz[k][lane], nv = v >= 0 ? v + delta : v - delta;
                        if (nv < -32768 || nv > 32767) report_coef_range(P.status, im->status_slot);
                        W.zz[k][lane] = (short)nv;
                    }
                }
                if (nhi >> lane & 1u) {
                    const uint32_t bp = p0 + c1 + (uint32_t)(__popc(nlo) + __popc(nhi & lt));
                    if ((W.ring[(bp >> 5) & (K3_RW - 1)] << (bp & 31)) >> 31) {
                        const int v = W.zz[k][lane + 32], nv = v >= 0 ? v + delta : v - delta;
                        if (nv < -32768 || nv > 32767) report_coef_range(P.status, im->status_slot);
                        W.zz[k][lane + 32] = (short)nv;
                    }
                }
            }
            __syncwarp();
        }
        if (j >= iv.n_blocks) break;
    }
    if (acref && !err && bn) store_batch();
    if (err && lane == 0) report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + j, err);
    // The reference keeps its End-Of-Band run across scans (decoder.zig:144, reset only at RSTn :1451); here
    // every scan starts from zero (scans of one level run side by side), so a run that is still open when a
    // scan ends (corrupt streams only) would make the next scan differ: refuse the image instead.
    // (the host then decodes the image again with its scans one after the other and the run handed on: zpx_api.cu)
    if (!err && lane == 0 && (iv.flags & 2u)) {
        if (P.eob_out != nullptr) P.eob_out[im->status_slot] = eob_run;
        else if (eob_run != 0)
            report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + iv.n_blocks, ZPX_E_UNSUPPORTED_STREAM);
    }
}

cudaError_t k3_launch_progressive(const K1Params& P, const uint32_t* list, int n_list, cudaStream_t s) {
    if (n_list <= 0) return cudaSuccess;
    k3_progressive<<<(n_list + K3_WARPS - 1) / K3_WARPS, K3_NT, 0, s>>>(P, list, n_list);
    return cudaGetLastError();
}

}  // namespace zpx
