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

#include "zpx_entropy.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

namespace zpx {

// ---------------------------------------------------------------------------
// K1a: one lane per restart interval (32 intervals per warp), serial inside the interval.
//
// The loop body decodes ONE Huffman symbol per lane per iteration, DC or AC alike (the table, the
// run/size split and the destination index are selects), so the 32 lanes of a warp -- which sit at
// different symbols of different blocks -- execute one common instruction stream.  Only the end of a
// block (flush of the 128-byte block, next block's descriptor) and the rare paths (codes longer than
// ZPX_LUT_BITS, FF 00 inside a refill word, errors) are divergent.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k1_lane_per_interval(const K1Params P) {
    __shared__ uint4 sblk[8 * NT];
    __shared__ uint8_t s_unzig[80];  // lane-divergent index: shared, not constant, memory (padded: k+run <= 78)
    const int gid = blockIdx.x * NT + threadIdx.x;
    LaneBlock<NT> lb;
    lb.base = sblk + threadIdx.x;
    lb.clear();
    if (threadIdx.x < 80) s_unzig[threadIdx.x] = threadIdx.x < 64 ? c_unzig[threadIdx.x] : 63;
    __syncthreads();

    // lanes past the end of the interval list idle through the loop (the loop head is a warp vote)
    const ZpxIntervalDev iv = P.ivs[gid < P.n_iv ? gid : P.n_iv - 1];
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const int err_eof = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;

    BitReader br;
    br.init(P.blob, iv.start, iv.len);

    const bool interleaved = sc->interleaved != 0;
    const int nblk = interleaved ? sc->nblk : 1;
    const uint32_t mxx = (uint32_t)im->mxx;
    const bool planar = im->layout == ZPX_LAYOUT_PLANAR || !interleaved;
    const uint4* __restrict__ bpack = reinterpret_cast<const uint4*>(sc->blk_pack);
    const uint32_t cw = (uint32_t)sc->cw;

    // position of the current block
    uint32_t mcu = iv.first_mcu, mx = 0, my = 0, bxn = 0, byn = 0;
    if (interleaved) {
        mx = mcu % mxx;
        my = mcu / mxx;
    } else {  // n-th coded block of the component, row-major over the blocks that intersect the image
        byn = iv.first_block / cw;
        bxn = iv.first_block - byn * cw;
    }
    int c = 0;
    uint4 bi = bpack[0];
    const ZpxHuffDev* __restrict__ tdc = &P.huff[bi.x];
    const ZpxHuffDev* __restrict__ tac = &P.huff[bi.y];

    int dc0 = 0, dc1 = 0, dc2 = 0, dc3 = 0;
    uint32_t eob_run = 0;
    int k = 0;                       // 0: the next symbol is the block's DC; 1..63: next AC index
    uint32_t left = gid < P.n_iv ? iv.n_blocks : 0;  // blocks still to decode (including the current one)
    int err = 0;

    // One Huffman symbol per lane per iteration; the vote keeps the warp in lock step.
    while (__any_sync(0xffffffffu, left != 0)) {
        if (left != 0) {
            SymOut so;
            err = symbol_step<false>(br, tdc, tac, bi.w, (int)(bi.z & 0xff), k, eob_run, dc0, dc1, dc2, dc3, so);
            const bool done = so.done;
            if (so.store) lb.put(s_unzig[so.kk], so.v);
            if (err) {
                // a symbol that needed bits past the limit is the reference's MissingFF00 / UnexpectedEof,
                // whatever the garbage decoded to
                if (br.overrun()) err = err_eof;
                report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + (iv.n_blocks - left), err);
                left = 0;
            } else if (done) {
                if (br.overrun()) {
                    report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + (iv.n_blocks - left), err_eof);
                    left = 0;
                } else {
                    // ---- hand the block to HBM ----
                    const int comp = (int)(bi.z & 0xff);
                    int bx, by;
                    if (interleaved) {
                        bx = (int)(bi.w & 0xff) * (int)mx + (int)((bi.z >> 8) & 0xff);
                        by = (int)((bi.w >> 8) & 0xff) * (int)my + (int)((bi.z >> 16) & 0xff);
                    } else {
                        bx = (int)bxn;
                        by = (int)byn;
                    }
                    uint64_t blk;
                    if (planar) blk = im->comp_base[comp] + (uint64_t)by * im->comp_bw[comp] + bx;
                    else blk = im->coef_base + (uint64_t)mcu * im->bpm + (bi.z >> 24);
                    lb.flush(P.coef + blk * 8, bx & 7);
                    // ---- next block ----
                    left--;
                    k = 0;
                    if (interleaved) {
                        if (++c == nblk) {
                            c = 0;
                            mcu++;
                            if (++mx == mxx) { mx = 0; my++; }
                        }
                        bi = bpack[c];
                        tdc = &P.huff[bi.x];
                        tac = &P.huff[bi.y];
                    } else if (++bxn == cw) {
                        bxn = 0;
                        byn++;
                    }
                }
            }
        }
    }
}

cudaError_t k1_launch_lane_per_interval(const K1Params& P, cudaStream_t s) {
    if (P.n_iv <= 0) return cudaSuccess;
    constexpr int NT = 128;
    k1_lane_per_interval<NT><<<(P.n_iv + NT - 1) / NT, NT, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace zpx
