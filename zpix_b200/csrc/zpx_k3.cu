// zpx_k3.cu -- progressive (SOF2) scans on the GPU.
//
// Replaces the progressive branches of processSos (src/jpeg/decoder.zig:1268-1283, 1340-1425),
// refine (:1459-1518) and refineNonZeroes (:1522-1549).  The coefficients of a progressive frame
// live in per-component planar grids of int16 blocks in HBM (the reference's
// progressive_coefficients, :1341/:1415), zeroed before the first scan; every scan is one launch
// that reads-modifies-writes them, one lane per restart interval (or per scan when DRI = 0).
// Scans of one image are launched in file order (scan k+1 refines what scan k wrote); scans of
// different images with the same ordinal share a launch.  After the last scan the unfused
// k2g kernels reconstruct only the blocks that intersect the image (decoder.zig:1636-1661).
//
// This first version is serial inside an interval (correctness first): spectral selection and
// successive approximation make the bit consumption depend on the coefficient state, which the
// speculative decoder of zpx_k1s.cu does not model yet.
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_entropy.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

namespace zpx {

namespace {

struct Bits {
    BitReader br;
    __device__ __forceinline__ int bit() {
        br.fill();
        const int b = (int)(br.peek32() >> 31);
        br.consume(1);
        return b;
    }
    __device__ __forceinline__ uint32_t bits(int n) {  // 1 <= n <= 16
        br.fill();
        const uint32_t v = br.peek32() >> (32 - n);
        br.consume(n);
        return v;
    }
    // decodeHuffman; -1 = BadHuffmanCode
    __device__ __forceinline__ int huff(const ZpxHuffDev* __restrict__ t) {
        br.fill();
        const HuffSym hs = huff_decode(t, br.peek32());
        if (hs.len == 0) {
            br.consume(16);
            return -1;
        }
        br.consume(hs.len);
        return (int)hs.sym;
    }
    __device__ __forceinline__ int extend(int size) {  // RECEIVE + EXTEND on the next `size` bits
        if (size == 0) return 0;
        br.fill();
        const int v = receive_extend(br.buf, 0, size);
        br.consume(size);
        return v;
    }
};

// coefficient `nat` (natural index) of a block whose rows are stored XOR-swizzled by key
__device__ __forceinline__ short* cptr(short* blk, int key, int nat) { return blk + (((nat >> 3) ^ key) << 3) + (nat & 7); }

// Refinement passes read the coefficients they refine.  The block is copied to shared memory once
// (eight 16-byte loads in flight together: one memory latency per block instead of one per
// coefficient), refined there, and written back.  Layout in shared memory: plain, stored-slot order.
constexpr int K3_NT = 64;
__device__ __forceinline__ void block_load(short* sm, const short* blk) {
    const uint4* g = reinterpret_cast<const uint4*>(blk);
    uint4* d = reinterpret_cast<uint4*>(sm);
    uint4 r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = g[i];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = r[i];
}
__device__ __forceinline__ void block_store(short* blk, const short* sm) {
    uint4* g = reinterpret_cast<uint4*>(blk);
    const uint4* d = reinterpret_cast<const uint4*>(sm);
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] = d[i];
}

// refineNonZeroes (decoder.zig:1522-1549)
__device__ int refine_non_zeroes(Bits& bs, short* blk, int key, const uint8_t* unzig, int zig, int zig_end, int nz, int delta) {
    for (; zig <= zig_end; zig++) {
        short* p = cptr(blk, key, unzig[zig]);
        const int v = *p;
        if (v == 0) {
            if (nz == 0) break;
            nz--;
            continue;
        }
        if (!bs.bit()) continue;
        *p = (short)(v >= 0 ? v + delta : v - delta);
    }
    return zig;
}

}  // namespace

__global__ void __launch_bounds__(K3_NT) k3_progressive(const K1Params P, const uint32_t* __restrict__ list, const int n_list) {
    __shared__ uint8_t s_unzig[64];
    __shared__ __align__(16) short s_blk[K3_NT / 32][64];  // one block per interval for the refinement passes
    if (threadIdx.x < 64) s_unzig[threadIdx.x] = c_unzig[threadIdx.x];
    __syncthreads();
    // One interval per WARP, decoded by lane 0: the nested, data-dependent loops of spectral selection /
    // refinement make lanes of one warp diverge completely (32-fold serialisation when every lane carries
    // its own interval), and the kernel is latency-bound anyway.
    const int gid = blockIdx.x * (K3_NT / 32) + (threadIdx.x >> 5);
    if (gid >= n_list || (threadIdx.x & 31) != 0) return;
    const ZpxIntervalDev iv = P.ivs[list[gid]];
    const ZpxScanDev* __restrict__ sc = &P.scans[iv.scan];
    const ZpxImageDev* __restrict__ im = &P.imgs[sc->img];
    const int err_eof = (iv.flags & 1) ? ZPX_E_UnexpectedEof : ZPX_E_MissingFF00;
    const int ss = sc->ss, se = sc->se, ah = sc->ah, al = sc->al;
    const bool interleaved = sc->interleaved != 0;
    const int nblk = interleaved ? sc->nblk : 1;
    const uint32_t mxx = (uint32_t)im->mxx, cw = (uint32_t)sc->cw;

    Bits bs;
    bs.br.init(P.blob, iv.start, iv.len);

    uint32_t mcu = iv.first_mcu, mx = 0, my = 0, bxn = 0, byn = 0;
    if (interleaved) {
        mx = mcu % mxx;
        my = mcu / mxx;
    } else {
        byn = iv.first_block / cw;
        bxn = iv.first_block - byn * cw;
    }
    int c = 0;
    int dc[4] = {0, 0, 0, 0};
    uint32_t eob_run = 0;
    int err = 0;
    uint32_t j = 0;
    for (; j < iv.n_blocks && !err; j++) {
        const int comp = sc->blk_comp[c];
        int bx, by;
        if (interleaved) {
            bx = im->h[comp] * (int)mx + sc->blk_hx[c];
            by = im->v[comp] * (int)my + sc->blk_vy[c];
        } else {
            bx = (int)bxn;
            by = (int)byn;
        }
        short* gblk = reinterpret_cast<short*>(P.coef) + (im->comp_base[comp] + (uint64_t)by * im->comp_bw[comp] + bx) * 64;
        // AC refinement works on a shared-memory copy; every other pass only writes (or ORs one bit)
        const bool cached = ah != 0 && ss != 0;
        short* blk = cached ? s_blk[threadIdx.x >> 5] : gblk;
        if (cached) block_load(blk, gblk);
        const int key = bx & 7;
        const ZpxHuffDev* __restrict__ tdc = &P.huff[sc->blk_dc[c]];
        const ZpxHuffDev* __restrict__ tac = &P.huff[sc->blk_ac[c]];
        const uint32_t tflags = sc->blk_pack[c][3];  // bit 16: DC table undefined, bit 17: AC table undefined
        const bool dc_undef = (tflags & 0x10000u) != 0, ac_undef = (tflags & 0x20000u) != 0;

        if (ah != 0) {
            // ---- successive-approximation refinement (decoder.zig:1459-1518) ----
            const int delta = 1 << al;
            if (ss == 0) {
                if (bs.bit()) {
                    // b[0] |= delta (decoder.zig:1464-1467) as a one-way atomic OR on the 32-bit word that holds
                    // the coefficient: no read latency on the lane's serial path
                    short* p = cptr(blk, key, 0);
                    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                    atomicOr(reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3), (unsigned int)(delta & 0xffff) << ((a & 2) ? 16 : 0));
                }
            } else {
                int zig = ss;
                if (eob_run == 0) {
                    while (zig <= se) {
                        if (ac_undef) { err = ZPX_E_UninitializedHuffmanTable; break; }
                        const int sym = bs.huff(tac);
                        if (sym < 0) { err = ZPX_E_BadHuffmanCode; break; }
                        const int r = sym >> 4, s = sym & 15;
                        int z = 0;
                        bool stop = false;
                        if (s == 0) {
                            if (r != 15) {
                                eob_run = 1u << r;
                                if (r != 0) eob_run |= bs.bits(r);
                                stop = true;
                            }
                        } else if (s == 1) {
                            z = bs.bit() ? delta : -delta;
                        } else {
                            err = ZPX_E_UnexpectedHuffmanCode;
                            break;
                        }
                        if (stop) break;
                        zig = refine_non_zeroes(bs, blk, key, s_unzig, zig, se, r, delta);
                        if (zig > se) { err = ZPX_E_TooManyCoefficients; break; }
                        if (z != 0) *cptr(blk, key, s_unzig[zig]) = (short)z;
                        zig++;
                    }
                }
                if (!err && eob_run > 0) {
                    eob_run--;
                    refine_non_zeroes(bs, blk, key, s_unzig, zig, se, -1, delta);
                }
            }
        } else {
            // ---- first pass of a band (decoder.zig:1362-1411) ----
            int zig = ss;
            if (zig == 0) {
                zig++;
                if (dc_undef) { err = ZPX_E_UninitializedHuffmanTable; break; }
                const int t = bs.huff(tdc);
                if (t < 0) { err = ZPX_E_BadHuffmanCode; break; }
                if (t > 16) { err = ZPX_E_ExcessiveDCComponent; break; }
                dc[comp] += bs.extend(t);
                const int v = (int)((uint32_t)dc[comp] << al);
                if (v < -32768 || v > 32767) { err = ZPX_E_COEF_RANGE; break; }
                *cptr(blk, key, 0) = (short)v;
            }
            if (zig <= se && eob_run > 0) {
                eob_run--;
            } else {
                while (zig <= se) {
                    if (ac_undef) { err = ZPX_E_UninitializedHuffmanTable; break; }
                    const int sym = bs.huff(tac);
                    if (sym < 0) { err = ZPX_E_BadHuffmanCode; break; }
                    const int r = sym >> 4, s = sym & 15;
                    if (s != 0) {
                        zig += r;
                        if (zig > se) break;
                        const int v = (int)((uint32_t)bs.extend(s) << al);
                        if (v < -32768 || v > 32767) { err = ZPX_E_COEF_RANGE; break; }
                        *cptr(blk, key, s_unzig[zig]) = (short)v;
                    } else {
                        if (r != 15) {
                            eob_run = 1u << r;
                            if (r != 0) eob_run |= bs.bits(r);
                            eob_run = (eob_run - 1) & 0xffffu;
                            break;
                        }
                        zig += 15;
                    }
                    zig++;
                }
            }
        }
        if (cached) block_store(gblk, blk);
        if (bs.br.overrun()) err = err_eof;
        if (err) break;

        if (interleaved) {
            if (++c == nblk) {
                c = 0;
                mcu++;
                if (++mx == mxx) { mx = 0; my++; }
            }
        } else if (++bxn == cw) {
            bxn = 0;
            byn++;
        }
    }
    if (err) {
        if (bs.br.overrun()) err = err_eof;
        report(P.status, im->status_slot, sc->scan_index, (uint64_t)iv.first_block + j, err);
    }
}

cudaError_t k3_launch_progressive(const K1Params& P, const uint32_t* list, int n_list, cudaStream_t s) {
    if (n_list <= 0) return cudaSuccess;
    const int per_cta = K3_NT / 32;
    k3_progressive<<<(n_list + per_cta - 1) / per_cta, K3_NT, 0, s>>>(P, list, n_list);
    return cudaGetLastError();
}

}  // namespace zpx
