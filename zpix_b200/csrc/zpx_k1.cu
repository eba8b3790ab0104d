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
// K1a: one lane per restart interval (32 intervals per warp), serial inside the interval
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k1_lane_per_interval(const K1Params P) {
    __shared__ uint4 sblk[8 * NT];
    __shared__ uint8_t s_unzig[64];  // lane-divergent index: shared, not constant, memory
    const int gid = blockIdx.x * NT + threadIdx.x;
    LaneBlock<NT> lb;
    lb.base = sblk + threadIdx.x;
    lb.clear();
    if (threadIdx.x < 64) s_unzig[threadIdx.x] = c_unzig[threadIdx.x];
    __syncthreads();
    if (gid >= P.n_iv) return;

    const ZpxIntervalDev iv = P.ivs[gid];
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const int err_eof = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;

    BitReader br;
    br.init(P.blob, iv.start, iv.len);

    const int nblk = sc->nblk;
    const int mxx = im->mxx;
    const bool interleaved = sc->interleaved != 0;
    const bool planar = im->layout == ZPX_LAYOUT_PLANAR;
    int dc0 = 0, dc1 = 0, dc2 = 0, dc3 = 0;
    uint32_t eob_run = 0;
    int err = 0;
    uint64_t ordinal = (uint64_t)iv.first_mcu * nblk;  // block ordinal inside the scan (error ordering)

    uint32_t mcu = iv.first_mcu;
    int mx = (int)(mcu % (uint32_t)mxx), my = (int)(mcu / (uint32_t)mxx);
    // non-interleaved scans: linear block counter over the component's MCU-padded grid
    // (decoder.zig:1331-1336); blocks outside the image carry no data.
    const int c0 = sc->blk_comp[0];
    const int ni_h = im->h[c0], ni_v = im->v[c0];
    uint32_t block_count = interleaved ? 0 : iv.first_mcu * (uint32_t)(ni_h * ni_v);
    const int ni_bw = mxx * ni_h;

    for (uint32_t m = 0; m < iv.n_mcu && !err; m++) {
        for (int b = 0; b < nblk && !err; b++, ordinal++) {
            const int comp = interleaved ? sc->blk_comp[b] : c0;
            int bx, by;
            if (interleaved) {
                bx = im->h[comp] * mx + sc->blk_hx[b];
                by = im->v[comp] * my + sc->blk_vy[b];
            } else {
                bx = (int)(block_count % (uint32_t)ni_bw);
                by = (int)(block_count / (uint32_t)ni_bw);
                block_count++;
                if (bx * 8 >= im->width || by * 8 >= im->height) continue;
            }
            const ZpxHuffDev* __restrict__ tdc = &P.huff[sc->blk_dc[interleaved ? b : 0]];
            const ZpxHuffDev* __restrict__ tac = &P.huff[sc->blk_ac[interleaved ? b : 0]];

            // ---- DC (decoder.zig:1366-1376) ----
            br.fill();
            if (!tdc->defined) { err = ZPX_E_UninitializedHuffmanTable; break; }
            HuffSym hs = huff_decode(tdc, br.peek32());
            if (hs.len == 0) {
                br.consume(16);
                err = br.overrun() ? err_eof : ZPX_E_BadHuffmanCode;
                break;
            }
            if (hs.sym > 16) {
                br.consume(hs.len);
                err = br.overrun() ? err_eof : ZPX_E_ExcessiveDCComponent;
                break;
            }
            const int diff = receive_extend(br.buf, hs.len, (int)hs.sym);
            br.consume(hs.len + (int)hs.sym);
            int dc = comp == 0 ? dc0 : comp == 1 ? dc1 : comp == 2 ? dc2 : dc3;
            dc += diff;
            if (comp == 0) dc0 = dc; else if (comp == 1) dc1 = dc; else if (comp == 2) dc2 = dc; else dc3 = dc;
            if (dc < -32768 || dc > 32767) { err = br.overrun() ? err_eof : ZPX_E_COEF_RANGE; break; }
            lb.put(0, dc);

            // ---- AC (decoder.zig:1378-1411) ----
            if (eob_run > 0) {
                eob_run--;
            } else {
                if (!tac->defined) { err = ZPX_E_UninitializedHuffmanTable; break; }
                int k = 1;
                while (k <= 63) {
                    br.fill();
                    hs = huff_decode(tac, br.peek32());
                    if (hs.len == 0) {
                        br.consume(16);
                        err = br.overrun() ? err_eof : ZPX_E_BadHuffmanCode;
                        break;
                    }
                    const int r = (int)(hs.sym >> 4), s = (int)(hs.sym & 15);
                    if (s != 0) {
                        k += r;
                        if (k > 63) {
                            br.consume(hs.len);
                            break;
                        }
                        const int ac = receive_extend(br.buf, hs.len, s);
                        br.consume(hs.len + s);
                        lb.put(s_unzig[k], ac);
                        k++;
                    } else if (r != 15) {
                        // EOB, or the EOB-run form that the reference also honours in sequential scans
                        eob_run = 1u << r;
                        if (r != 0) eob_run |= (uint32_t)((br.buf << hs.len) >> (64 - r));
                        eob_run = (eob_run - 1) & 0xffffu;
                        br.consume(hs.len + r);
                        break;
                    } else {
                        br.consume(hs.len);
                        k += 16;
                    }
                }
                if (err) break;
            }
            if (br.overrun()) { err = err_eof; break; }

            // ---- hand the block to HBM ----
            uint64_t blk;
            if (planar) blk = im->comp_base[comp] + (uint64_t)by * im->comp_bw[comp] + bx;
            else blk = im->coef_base + (uint64_t)mcu * im->bpm + sc->blk_slot[b];
            lb.flush(P.coef + blk * 8, bx & 7);
        }
        mcu++;
        if (++mx == mxx) { mx = 0; my++; }
    }
    if (err) report(P.status, im->status_slot, sc->scan_index, ordinal, err);
}

cudaError_t k1_launch_lane_per_interval(const K1Params& P, cudaStream_t s) {
    if (P.n_iv <= 0) return cudaSuccess;
    constexpr int NT = 128;
    k1_lane_per_interval<NT><<<(P.n_iv + NT - 1) / NT, NT, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace zpx
