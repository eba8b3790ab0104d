// zpx_k2.cu -- block reconstruction and colour: the part of the reference after entropy decode.
//
//   k2_fused<H,V,NC>   dequantise + integer IDCT + chroma replication + YCbCr->RGBA + clamp + store,
//                      one pass over HBM (coefficients in, RGBA out).  Replaces, per image,
//                      reconstructBlock (decoder.zig:1553-1634), idct.transform (idct.zig:77-201),
//                      the YCbCr planes of makeImg (:1708-1783), Image.rgbaPixels (image.zig:103-130),
//                      YCbCrImage.cOffset (image.zig:594-605) and Color.toRGBA (color.zig:90-126).
//   k2g_idct_planes /  the same work unfused (IDCT -> native planes -> colour), for every stream shape
//   k2g_colour         the fused kernel does not take (CMYK/YCCK, RGB-tagged, Y 2x2 + C 1x2, multi-scan
//                      and progressive frames) and for the native-variant output (jpeg.load's planes).
//
// Data layout in HBM: coefficients are int16, natural order, 128 B per block,
// blocks in MCU-interleaved scan order; the eight 16-byte rows of a block are stored XOR-swizzled by
// (bx & 7), bx = the block's x index inside its component, so that the linear bulk copy of a tile
// lands in shared memory bank-conflict-free for one-thread-per-block row loads.
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_idct.cuh"
#include "zpx_internal.h"
#include "zpx_kernels.h"

// resident threads per SM the fused kernel is compiled for (register cap = 65536 / that).  Measured on one B200
// (tools/k2_variants.sh, K2 time, 512 / 640 / 768): 4:2:0 1.72 / 1.72 / 1.87 ms, 4:4:4 1.54 / 1.63 / 1.59, 4:2:2 4.19 /
// 4.34 / 4.78, gray 0.600 / 0.573 / 0.868 -- colour stays at 512 (126 registers, no spills), gray takes 640 (102).
#ifndef ZPX_K2_TILE_QIDX
#define ZPX_K2_TILE_QIDX 0
#endif
#ifndef ZPX_K2_420_ROWS6
#define ZPX_K2_420_ROWS6 1
#endif
#ifndef ZPX_K2_422_ROWS6
#define ZPX_K2_422_ROWS6 0
#endif
#ifndef ZPX_K2_GRAY_ROWS7
#define ZPX_K2_GRAY_ROWS7 1
#endif
#ifndef ZPX_K2_THREADS_PER_SM
#define ZPX_K2_THREADS_PER_SM 512
#endif
#ifndef ZPX_K2_THREADS_PER_SM_GRAY
#define ZPX_K2_THREADS_PER_SM_GRAY 640
#endif

namespace zpx {

__host__ __device__ constexpr int k2_threads_per_sm(int nc) { return nc == 1 ? ZPX_K2_THREADS_PER_SM_GRAY : ZPX_K2_THREADS_PER_SM; }

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA unit, non-tensor form) helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------------------
// fused kernel
// ---------------------------------------------------------------------------
// Shared memory (dynamic):
//   [0,32)     the stages' mbarriers
//   PL         Y plane 8V rows x PY bytes, then Cb, Cr: 8 rows x PC bytes each
//   ST0..      NS coefficient stages, tmax * BPM * 128 bytes each (ring fed by bulk copies)
template <int H, int V, int NC>
struct K2Cfg {
    static constexpr int BPM = NC == 1 ? 1 : H * V + 2;
    static constexpr int YROWS = 8 * V;
    static constexpr int MCU_W = 8 * H;
    __host__ __device__ static int pitch_y(int tmax) { return tmax * MCU_W + 16; }
    __host__ __device__ static int pitch_c(int tmax) { return tmax * 8 + 16; }
    __host__ __device__ static size_t smem_bytes(int tmax, int ns) {
        size_t s = 32;
        s += (size_t)YROWS * pitch_y(tmax);
        if (NC == 1) s += 2048;  // gray tiles may span up to 16 MCU rows, each with 16 bytes of row padding
        if (NC == 3) s += (size_t)2 * 8 * pitch_c(tmax);
        s = (s + 127) & ~(size_t)127;
        s += (size_t)ns * tmax * BPM * 128;
        return s;
    }
};

// NTMAX bounds the registers (512 threads per SM); the launch uses as many threads as the group's largest tile has
// blocks (K2Params::nt, a multiple of 32), so that phase 1 leaves no thread without a block.
template <int H, int V, int NC, int NS, int NTMAX>
__global__ void __launch_bounds__(NTMAX, k2_threads_per_sm(NC) / NTMAX) k2_fused(const K2Params P) {
    const int NT = (int)blockDim.x;
    using Cfg = K2Cfg<H, V, NC>;
    constexpr int BPM = Cfg::BPM;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    const int PY0 = Cfg::pitch_y(P.tmax), PC = Cfg::pitch_c(P.tmax);
    const int PY = PY0;  // layout constant; gray tiles use a per-tile pitch inside the same area
    uint8_t* planeY = smem + 32;
    uint8_t* planeCb = planeY + Cfg::YROWS * PY;
    uint8_t* planeCr = planeCb + 8 * PC;
    size_t st_off = 32 + (size_t)Cfg::YROWS * PY + (NC == 3 ? (size_t)2 * 8 * PC : 2048);
    st_off = (st_off + 127) & ~(size_t)127;
    uint8_t* stage0 = smem + st_off;
    const uint32_t stage_bytes = (uint32_t)P.tmax * BPM * 128;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NS; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint4* __restrict__ coef = reinterpret_cast<const uint4*>(P.coef);
    int tile = blockIdx.x;
    if (tile >= P.ntiles) return;

    // Per-stage tile context and quantisers live in shared memory: written when the tile's bulk copy
    // is issued (NS-1 iterations ahead), read after the mbarrier wait -- no global-load latency and no
    // extra barrier on the per-tile critical path.
    struct TileCtx {
        uint32_t n, mx0, my, img;
        int32_t width, height;
        uint64_t out_off;
        uint32_t nr, wt;  // gray: the tile is nr whole MCU rows of wt MCUs (n = nr * wt); else nr = 1, wt = n
        uint32_t wide;    // the image has coefficients outside [-4096, 4095]: exact all-AC-zero rows (zpx_idct.cuh)
        // native planes (makeImg's layout, decoder.zig:1708-1783), when the launch writes them
        uint64_t poff[3];
        int32_t ystride, cstride;
        uint32_t magic;   // phase 2: 2^32 / (items per row) + 1, for the exact division of an item index by it
        uint32_t interior;  // every pixel of the tile lies inside the image and rows are 16-byte multiples: no bounds tests
        uint32_t planar;    // the image keeps one block grid per component (multi-scan and progressive frames): the stage
                            // then holds the tile's blocks component by component, in phase 1's own thread order
    };
    __shared__ TileCtx ctx[NS];
    __shared__ __align__(16) uint32_t qsm[NS][3 * 32];  // per row four words q[2j] | q[2j+1] << 24 (dp2a operands)

    // issue the bulk copy of one tile into a stage, publish its context and stage the image's quantisers when the
    // image changes.  One warp does all of it: the image descriptor, its flags and the quantisers are dependent
    // global loads (about a thousand cycles before the copy can be issued), and the warp that waits for them reaches
    // the barrier after phase 1 that much later, with every other warp waiting there (ncu: 16 % of the warp samples
    // sit at that barrier).  A tile's threads are ordered luma, Cb, Cr, so the LAST warp holds chroma blocks (sparse
    // IDCT), is only partly filled, or has no block at all (4:4:4 tiles of 96 blocks in a 128-thread CTA): it has the
    // slack.  Measured (K2 time, same box): 4:4:4 -7 %, 4:2:2 -3.8 %; 4:2:0 +4 % and gray +2 %, which keep warp 0.
    int pf_img = -1;  // image whose quantisers were staged last (uniform)
    constexpr bool FETCH_LAST = NC == 3 && !(H == 2 && V == 2);
    const int ftid = FETCH_LAST ? NT - 32 : 0;  // first thread of the fetching warp
    // quantiser table of component c: colour tiles carry the indices (pad: bit 31, three 10-bit fields holding
    // index + 1), so the table loads do not depend on the image descriptor's
    auto qindex = [&](const ZpxTileDev& tn, const ZpxImageDev* imn, int c) -> int {
        // (colour: measured -1 % for 4:2:0 and 4:4:4, +1 % for 4:2:2 -- off unless ZPX_K2_TILE_QIDX; gray: +4 %)
        if (ZPX_K2_TILE_QIDX && NC == 3 && (tn.pad >> 31)) return (int)((tn.pad >> (10 * c)) & 1023u) - 1;
        if (NC == 1 && ((tn.pad >> 8) & 0xffu)) return (int)((tn.pad >> 8) & 0xffu) - 1;  // gray: rows | index + 1 << 8 | width << 16
        return imn->qidx[c];
    };
    auto fetch = [&](const ZpxTileDev tn, int stg) {
        const ZpxImageDev* __restrict__ imn = &P.imgs[tn.img];
        if (tid == ftid) {
            const uint32_t bytes = (uint32_t)tn.n * BPM * 128;
            const bool planar = NC == 3 && imn->layout == ZPX_LAYOUT_PLANAR;
            uint8_t* const sdst = stage0 + (size_t)stg * stage_bytes;
            mbar_arrive_expect_tx(&bars[stg], bytes);
            if (!planar) {
                const uint64_t blk0 = imn->coef_base + ((uint64_t)tn.my * imn->mxx + tn.mx0) * BPM;
                bulk_g2s(sdst, coef + blk0 * 8, bytes, &bars[stg]);
            } else {
                // V runs of n*H luma blocks (one per block row of the MCU row), then n Cb and n Cr blocks: block i of
                // phase 1 lands in slot i
                const uint32_t yb = (uint32_t)tn.n * H * 128, cb = (uint32_t)tn.n * 128;
#pragma unroll
                for (int vy = 0; vy < V; vy++) {
                    const uint64_t b0 = imn->comp_base[0] + ((uint64_t)tn.my * V + vy) * (uint32_t)imn->comp_bw[0] + (uint64_t)tn.mx0 * H;
                    bulk_g2s(sdst + vy * yb, coef + b0 * 8, yb, &bars[stg]);
                }
#pragma unroll
                for (int cc = 1; cc <= 2; cc++) {
                    const uint64_t b0 = imn->comp_base[cc] + (uint64_t)tn.my * (uint32_t)imn->comp_bw[cc] + tn.mx0;
                    bulk_g2s(sdst + V * yb + (cc - 1) * cb, coef + b0 * 8, cb, &bars[stg]);
                }
            }
            TileCtx c;
            c.planar = planar ? 1u : 0u;
            c.n = tn.n;
            c.mx0 = tn.mx0;
            c.my = tn.my;
            c.img = tn.img;
            c.width = imn->width;
            c.height = imn->height;
            c.out_off = imn->out_off;
            c.nr = (NC == 1 && (tn.pad & 0xffu)) ? tn.pad & 0xffu : 1u;
            c.wt = (NC == 1 && (tn.pad >> 16)) ? tn.pad >> 16 : tn.n;
            // (progressive scans do not track the coefficient range: their frames always take the exact rows)
            // (the flags are indexed by the image's slot on the device, which is its index in the image table)
            c.wide = (P.img_flags[tn.img] & 1u) | (imn->progressive ? 1u : 0u);
            c.poff[0] = imn->plane_off[0];
            c.poff[1] = imn->plane_off[1];
            c.poff[2] = imn->plane_off[2];
            c.ystride = imn->plane_stride[0];
            c.cstride = imn->plane_stride[1];
            {
                constexpr int PXW_ = (V == 2) ? 4 : 8;
                const uint32_t ipr_ = c.wt * (uint32_t)(Cfg::MCU_W / PXW_);
                c.magic = ipr_ > 1 ? 0xffffffffu / ipr_ + 1u : 0u;  // exact it / ipr for it, ipr < 2^16
                const int rows_ = NC == 1 ? (int)c.nr * 8 : Cfg::YROWS;
                c.interior = ((c.width & 3) == 0 && (int)(c.mx0 + c.wt) * Cfg::MCU_W <= c.width && (int)c.my * Cfg::YROWS + rows_ <= c.height) ? 1u : 0u;
            }
            ctx[stg] = c;
        }
        if (FETCH_LAST) {
            if (tid >= ftid) {
                // (every stage keeps its own copy: a later tile of another image must not disturb it)
                const int k = tid & 31;
                if ((int)tn.img != pf_img) {
                    const int* qp[NC];
#pragma unroll
                    for (int c = 0; c < NC; c++) qp[c] = P.quant[qindex(tn, imn, c)].q;
                    int2 qv[NC];
#pragma unroll
                    for (int c = 0; c < NC; c++) qv[c] = __ldg(reinterpret_cast<const int2*>(qp[c]) + k);  // all loads in flight together
#pragma unroll
                    for (int c = 0; c < NC; c++) qsm[stg][c * 32 + k] = (uint32_t)qv[c].x | (uint32_t)qv[c].y << 24;
                } else {
                    const int prev = stg == 0 ? NS - 1 : stg - 1;
#pragma unroll
                    for (int c = 0; c < NC; c++) qsm[stg][c * 32 + k] = qsm[prev][c * 32 + k];
                }
            }
        } else if (tid < 32 * NC) {  // one warp per component
            const int c = tid >> 5, k = tid & 31;
            if ((int)tn.img != pf_img) {
                const int* q = P.quant[qindex(tn, imn, c)].q;
                qsm[stg][c * 32 + k] = (uint32_t)q[2 * k] | (uint32_t)q[2 * k + 1] << 24;
            } else {
                const int prev = stg == 0 ? NS - 1 : stg - 1;
                qsm[stg][c * 32 + k] = qsm[prev][c * 32 + k];
            }
        }
        pf_img = (int)tn.img;
    };
    auto load_tile = [&](int i) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.tiles) + i);
        ZpxTileDev t;
        t.img = v.x;
        t.my = (uint16_t)(v.y & 0xffffu);
        t.n = (uint16_t)(v.y >> 16);
        t.mx0 = v.z;
        t.pad = v.w;
        return t;
    };
    // prologue: the first NS-1 tiles of this CTA
    for (int i = 0; i < NS - 1; i++) {
        if (tile + i * (int)gridDim.x < P.ntiles) fetch(load_tile(tile + i * (int)gridDim.x), i);
        __syncthreads();  // the copy of qsm[prev] above must see the previous stage's values
    }
    // descriptor of the tile to prefetch next, loaded one iteration early (its latency is hidden)
    ZpxTileDev tn_pf = load_tile(min(tile + (NS - 1) * (int)gridDim.x, P.ntiles - 1));

    uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s
    int stage = 0;
    for (; tile < P.ntiles; tile += gridDim.x) {
        // prefetch NS-1 tiles ahead; that stage was consumed one iteration ago (barrier after phase 1)
        {
            const int ahead = tile + (NS - 1) * (int)gridDim.x;
            const int stg = stage == 0 ? NS - 1 : stage - 1;
            if (ahead < P.ntiles) fetch(tn_pf, stg);
            tn_pf = load_tile(min(ahead + (int)gridDim.x, P.ntiles - 1));
        }

        mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
        phase_bits ^= 1u << stage;
        const TileCtx t = ctx[stage];
        const int n = (int)t.n;
        const uint32_t* __restrict__ qs_t = qsm[stage];

        // ---------------- phase 1: one thread per 8x8 block ----------------
        const uint4* st = reinterpret_cast<const uint4*>(stage0 + (size_t)stage * stage_bytes);
        const int nY = (NC == 1) ? n : n * H * V;
        const int nblk = n * BPM;
        // block i of the tile: its slot in the stage, absolute component-x (swizzle key), destination in the
        // plane tile and quantiser
        auto locate = [&](int i, int& slot, int& bxa, uint8_t*& dst, int& pitch, const uint32_t*& q) {
            if (NC == 1) {
                // gray: blocks of nr whole MCU rows, row-major
                const int wt = (int)t.wt;
                const int r = i / wt, col = i - r * wt;
                slot = i;
                bxa = (int)t.mx0 + col;  // (absolute: a wide gray row is cut into tiles that need not start at a multiple of 8)
                pitch = wt * 8 + 16;
                dst = planeY + (r * 8) * pitch + col * 8;
                q = qs_t;
            } else if (i < nY) {
                const int nH = n * H;
                const int vy = (V == 2 && i >= nH) ? 1 : 0;
                const int bx = i - vy * nH;
                const int m = bx / H, hx = bx % H;
                slot = t.planar ? i : m * BPM + vy * H + hx;
                bxa = (int)t.mx0 * H + bx;
                dst = planeY + (vy * 8) * PY + bx * 8;
                pitch = PY;
                q = qs_t;
            } else {
                const int j = i - nY;
                const int c = j >= n ? 1 : 0;
                const int m = j - c * n;
                slot = t.planar ? i : m * BPM + H * V + c;
                bxa = (int)t.mx0 + m;
                dst = (c ? planeCr : planeCb) + m * 8;
                pitch = PC;
                q = qs_t + 32 * (1 + c);
            }
        };
        if (!t.wide) {
            // (warp-uniform trip count: the vote below needs every lane)
            for (int i = tid; (i & ~31) < nblk; i += NT) {
                const bool valid = i < nblk;
                int slot, bxa, pitch;
                uint8_t* dst;
                const uint32_t* q;
                locate(valid ? i : nblk - 1, slot, bxa, dst, pitch, q);
                const uint4* blk = st + slot * 8;
                const int key = bxa & 7;
                uint4 c[8];
#pragma unroll
                for (int r = 0; r < 8; r++) c[r] = blk[r ^ key];
                // Sparse blocks, decided per warp by a vote.  ONE variant beside the general code per instantiation (a
                // third unrolled IDCT body was measured 2.5 - 11 % slower on every sampling: instruction cache), chosen
                // by measurement (K2 time on one B200, tools/k2_variants.sh):
                //   4:2:0: coefficient rows 6 and 7 zero in all 32 blocks of the warp -- luma blocks whose highest
                //          frequencies quantise to zero, and chroma -- take the IDCT without those rows (-2.1 % against
                //          the 4x4 variant, which only chroma warps take);
                //   gray:  row 7 zero (its luma keeps more high frequencies; the 4x4 variant never applies): -1.7 %;
                //   other samplings: nothing outside the top-left 4x4 corner (chroma blocks: a tile's threads are
                //          ordered luma, Cb, Cr) -> the 4x4 IDCT (4:2:2 with the rows-6-7 variant instead: +4 %)
                // (zpx_idct.cuh: the same operations on the same operands, minus the ones on constant zeros)
                constexpr bool ROWS6 = NC == 3 && ((ZPX_K2_420_ROWS6 && H == 2 && V == 2) || (ZPX_K2_422_ROWS6 && H == 2 && V == 1));
                constexpr bool ROWS7 = ZPX_K2_GRAY_ROWS7 && NC == 1;  // gray: only row 7 (its luma keeps more high frequencies)
                uint32_t hi = c[7].x | c[7].y | c[7].z | c[7].w;
                if (!ROWS7) hi |= c[6].x | c[6].y | c[6].z | c[6].w;
                if (!ROWS6 && !ROWS7) {
                    hi |= c[4].x | c[4].y | c[4].z | c[4].w | c[5].x | c[5].y | c[5].z | c[5].w;
#pragma unroll
                    for (int r = 0; r < 4; r++) hi |= c[r].z | c[r].w;
                }
                uint32_t px[16];
                if (P.dense_only || __any_sync(0xffffffffu, valid && hi != 0))
                    dequant_idct_block_q8([&](int r) { return c[r]; }, q, px);
                else if (ROWS7)
                    dequant_idct_block_q8_sparse<7, false>([&](int r) { return c[r]; }, q, px);
                else if (ROWS6)
                    dequant_idct_block_q8_sparse<6, false>([&](int r) { return c[r]; }, q, px);
                else
                    dequant_idct_block_q8_sparse<4, true>([&](int r) { return c[r]; }, q, px);
                if (valid) {
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        *reinterpret_cast<uint2*>(dst + r * pitch) = make_uint2(px[2 * r], px[2 * r + 1]);
                }
            }
        } else {
            // the image has coefficients outside [-4096, 4095] (garbage streams only): rows whose AC are all zero the
            // reference's way (zpx_idct.cuh)
            for (int i = tid; i < nblk; i += NT) {
                int slot, bxa, pitch;
                uint8_t* dst;
                const uint32_t* q;
                locate(i, slot, bxa, dst, pitch, q);
                idct_block_q8_exact(st + slot * 8, bxa & 7, q, dst, pitch);
            }
        }
        __syncthreads();

        // ---------------- native planes (jpeg.load's Image{.YCbCr} / {.Gray}) ----------------
        // The tile's 8x8 outputs already sit in shared memory as planes: copy them out with makeImg's strides,
        // MCU padding included (reconstructBlock stores every decoded block, decoder.zig:1611-1633).
        if (P.planes != nullptr) {
            const int wt = (NC == 1) ? (int)t.wt : n;
            const int PYt = (NC == 1) ? wt * 8 + 16 : PY;
            const int rowsY = (NC == 1) ? (int)t.nr * 8 : Cfg::YROWS;
            const int v8 = wt * (Cfg::MCU_W / 8);  // 8-byte units per luma row of the tile
            uint8_t* dY = P.planes + t.poff[0] + (size_t)((int)t.my * Cfg::YROWS) * t.ystride + (size_t)t.mx0 * Cfg::MCU_W;
            for (int i = tid; i < rowsY * v8; i += NT) {
                const int r = i / v8, c8 = i - r * v8;
                *reinterpret_cast<uint2*>(dY + (size_t)r * t.ystride + c8 * 8) = *reinterpret_cast<const uint2*>(planeY + r * PYt + c8 * 8);
            }
            if (NC == 3) {
                uint8_t* dB = P.planes + t.poff[1] + (size_t)((int)t.my * 8) * t.cstride + (size_t)t.mx0 * 8;
                uint8_t* dR = P.planes + t.poff[2] + (size_t)((int)t.my * 8) * t.cstride + (size_t)t.mx0 * 8;
                for (int i = tid; i < 8 * n; i += NT) {
                    const int r = i / n, c8 = i - r * n;
                    *reinterpret_cast<uint2*>(dB + (size_t)r * t.cstride + c8 * 8) = *reinterpret_cast<const uint2*>(planeCb + r * PC + c8 * 8);
                    *reinterpret_cast<uint2*>(dR + (size_t)r * t.cstride + c8 * 8) = *reinterpret_cast<const uint2*>(planeCr + r * PC + c8 * 8);
                }
            }
        }

        // ---------------- phase 2: colour + coalesced RGBA stores ----------------
        // One item = RP luma rows x PXW pixels.  V = 2: both rows of a chroma row, 4 px wide, so the
        // per-chroma-sample terms are computed once per 2x2 (4:2:0) replication; V = 1: 8 px wide.
        if (P.out != nullptr) {
            constexpr int RP = V;                    // luma rows per item
            constexpr int PXW = (V == 2) ? 4 : 8;    // pixels per item row
            constexpr int NCS = (PXW / H) > 0 ? (PXW / H) : 1;  // chroma samples per item row
            const int W = t.width, Hh = t.height;
            const int x0 = (int)t.mx0 * Cfg::MCU_W, y0 = (int)t.my * Cfg::YROWS;
            const int wt = (NC == 1) ? (int)t.wt : n;
            const int PYt = (NC == 1) ? wt * 8 + 16 : PY;
            const int ipr = wt * (Cfg::MCU_W / PXW);  // items per row(-pair)
            const int items = ipr * ((NC == 1 ? (int)t.nr * 8 : Cfg::YROWS) / RP);
            const uint32_t magic = t.magic;  // 2^32 / ipr + 1: exact it / ipr for it, ipr < 2^16
            uint8_t* __restrict__ outp = P.out + t.out_off;
            const bool vec_ok = (W & 3) == 0;
            // interior tiles, two mappings of items to threads (measured on B200, same box, K2 time): gray -2.7 % with
            // whole item rows per warp (no index division, constant address steps: 15 % fewer instructions per item);
            // colour 0 % (4:2:0) to +2.6 % (4:2:2, 4:4:4) that way, so colour keeps the flat item index
            constexpr bool P2_ROWS = NC == 1;
            if (t.interior && P2_ROWS) {
                // tile entirely inside the image (all but the right / bottom edge tiles): no bounds tests, one
                // 64-bit base address per tile, 32-bit offsets from it.  A warp takes whole item rows (row = warp,
                // warp + warps, ...; items lane, lane + 32, ... of it), so that an item's addresses are its row's
                // plus constant steps: no division of an item index by the row length.
                uint8_t* __restrict__ ob = outp + ((size_t)y0 * (size_t)W + (size_t)x0) * 4u;
                const uint32_t rowb = (uint32_t)W * 4u;
                const int nrp = (NC == 1 ? (int)t.nr * 8 : Cfg::YROWS) / RP;
                const int lane = tid & 31;
                for (int rp = tid >> 5; rp < nrp; rp += NT >> 5) {
                    const uint8_t* const yrow = planeY + (uint32_t)rp * (uint32_t)(RP * PYt);
                    const uint32_t crow = (uint32_t)rp * (uint32_t)(V == 2 ? 1 : RP) * (uint32_t)PC;
                    uint8_t* const orow = ob + (uint32_t)rp * (uint32_t)RP * rowb;
                    // (explicit induction variables: the addresses of item xg + 32 are those of item xg plus constants)
                    const uint8_t* yp = yrow + PXW * lane;
                    uint32_t coff = crow + (uint32_t)(PXW * lane) / H;
                    uint8_t* o = orow + (uint32_t)lane * (uint32_t)(PXW * 4);
                    for (int xg = lane; xg < ipr; xg += 32, yp += 32 * PXW, coff += 32 * PXW / H, o += 32 * PXW * 4) {
                    ChromaTerms ct[NCS];
                    if (NC == 3) {
                        uint32_t cbw[2] = {0, 0}, crw[2] = {0, 0};
                        if (NCS == 8) {
                            const uint2 a2 = *reinterpret_cast<const uint2*>(planeCb + coff), b2 = *reinterpret_cast<const uint2*>(planeCr + coff);
                            cbw[0] = a2.x; cbw[1] = a2.y; crw[0] = b2.x; crw[1] = b2.y;
                        } else if (NCS == 4) {
                            cbw[0] = *reinterpret_cast<const uint32_t*>(planeCb + coff);
                            crw[0] = *reinterpret_cast<const uint32_t*>(planeCr + coff);
                        } else if (NCS == 2) {
                            cbw[0] = *reinterpret_cast<const uint16_t*>(planeCb + coff);
                            crw[0] = *reinterpret_cast<const uint16_t*>(planeCr + coff);
                        } else {
                            cbw[0] = planeCb[coff];
                            crw[0] = planeCr[coff];
                        }
#pragma unroll
                        for (int k = 0; k < NCS; k++)
                            chroma_terms_w((int)__byte_perm(cbw[k >> 2], 0, 0x4440 + (k & 3)), (int)__byte_perm(crw[k >> 2], 0, 0x4440 + (k & 3)), ct[k]);
                    }
#pragma unroll
                    for (int r = 0; r < RP; r++) {
#pragma unroll
                        for (int g = 0; g < PXW / 4; g++) {
                            const uint32_t yw = *reinterpret_cast<const uint32_t*>(yp + r * PYt + 4 * g);
                            uint32_t p[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint32_t yv = __byte_perm(yw, 0, 0x4440 + j);
                                if (NC == 1) {
                                    p[j] = yv * 0x010101u | 0xff000000u;
                                } else {
                                    p[j] = ycc_pixel_w<ZPX_YCC_WIDE>((int)yv, ct[(4 * g + j) / H]);
                                }
                            }
                            __stcs(reinterpret_cast<uint4*>(o + (uint32_t)r * rowb) + g, make_uint4(p[0], p[1], p[2], p[3]));
                        }
                    }
                    }
                }
            } else if (t.interior) {
                // tile entirely inside the image (all but the right / bottom edge tiles): no bounds tests, one
                // 64-bit base address per tile, 32-bit offsets from it
                uint8_t* __restrict__ ob = outp + ((size_t)y0 * (size_t)W + (size_t)x0) * 4u;
                const uint32_t rowb = (uint32_t)W * 4u;
                for (int it = tid; it < items; it += NT) {
                    const uint32_t rp = ipr > 1 ? __umulhi((uint32_t)it, magic) : (uint32_t)it;
                    const uint32_t xg = (uint32_t)it - rp * (uint32_t)ipr;
                    int rr[NCS], gg[NCS], bb[NCS];
                    if (NC == 3) {
                        const uint32_t coff = rp * (uint32_t)(V == 2 ? 1 : RP) * (uint32_t)PC + (PXW * xg) / H;
                        uint32_t cbw[2] = {0, 0}, crw[2] = {0, 0};
                        if (NCS == 8) {
                            const uint2 a2 = *reinterpret_cast<const uint2*>(planeCb + coff), b2 = *reinterpret_cast<const uint2*>(planeCr + coff);
                            cbw[0] = a2.x; cbw[1] = a2.y; crw[0] = b2.x; crw[1] = b2.y;
                        } else if (NCS == 4) {
                            cbw[0] = *reinterpret_cast<const uint32_t*>(planeCb + coff);
                            crw[0] = *reinterpret_cast<const uint32_t*>(planeCr + coff);
                        } else if (NCS == 2) {
                            cbw[0] = *reinterpret_cast<const uint16_t*>(planeCb + coff);
                            crw[0] = *reinterpret_cast<const uint16_t*>(planeCr + coff);
                        } else {
                            cbw[0] = planeCb[coff];
                            crw[0] = planeCr[coff];
                        }
#pragma unroll
                        for (int k = 0; k < NCS; k++)
                            chroma_terms((int)__byte_perm(cbw[k >> 2], 0, 0x4440 + (k & 3)), (int)__byte_perm(crw[k >> 2], 0, 0x4440 + (k & 3)),
                                         rr[k], gg[k], bb[k]);
                    }
                    const uint8_t* yp = planeY + rp * (uint32_t)(RP * PYt) + PXW * xg;
                    uint8_t* o = ob + (rp * (uint32_t)RP * rowb + xg * (uint32_t)(PXW * 4));
#pragma unroll
                    for (int r = 0; r < RP; r++) {
#pragma unroll
                        for (int g = 0; g < PXW / 4; g++) {
                            const uint32_t yw = *reinterpret_cast<const uint32_t*>(yp + r * PYt + 4 * g);
                            uint32_t p[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint32_t yv = __byte_perm(yw, 0, 0x4440 + j);
                                if (NC == 1) {
                                    p[j] = yv * 0x010101u | 0xff000000u;
                                } else {
                                    const int ci = (4 * g + j) / H;
                                    p[j] = ycc_pixel((int)yv, rr[ci], gg[ci], bb[ci]);
                                }
                            }
                            __stcs(reinterpret_cast<uint4*>(o + (uint32_t)r * rowb) + g, make_uint4(p[0], p[1], p[2], p[3]));
                        }
                    }
                }
            } else
            for (int it = tid; it < items; it += NT) {
                const int rp = ipr > 1 ? (int)__umulhi((uint32_t)it, magic) : it;
                const int xg = it - rp * ipr;
                const int row0 = rp * RP;
                const int y = y0 + row0, x = x0 + PXW * xg;
                if (y >= Hh || x >= W) continue;
                int rr[NCS], gg[NCS], bb[NCS];
                if (NC == 3) {
                    const int crow = (V == 2) ? rp : row0;
                    const uint8_t* cbp = planeCb + crow * PC + (PXW * xg) / H;
                    const uint8_t* crp = planeCr + crow * PC + (PXW * xg) / H;
                    uint32_t cbw[2] = {0, 0}, crw[2] = {0, 0};
                    if (NCS == 8) {
                        const uint2 a2 = *reinterpret_cast<const uint2*>(cbp), b2 = *reinterpret_cast<const uint2*>(crp);
                        cbw[0] = a2.x; cbw[1] = a2.y; crw[0] = b2.x; crw[1] = b2.y;
                    } else if (NCS == 4) {
                        cbw[0] = *reinterpret_cast<const uint32_t*>(cbp);
                        crw[0] = *reinterpret_cast<const uint32_t*>(crp);
                    } else if (NCS == 2) {
                        cbw[0] = *reinterpret_cast<const uint16_t*>(cbp);
                        crw[0] = *reinterpret_cast<const uint16_t*>(crp);
                    } else {
                        cbw[0] = *cbp;
                        crw[0] = *crp;
                    }
#pragma unroll
                    for (int k = 0; k < NCS; k++)
                        chroma_terms((int)__byte_perm(cbw[k >> 2], 0, 0x4440 + (k & 3)), (int)__byte_perm(crw[k >> 2], 0, 0x4440 + (k & 3)),
                                     rr[k], gg[k], bb[k]);
                }
#pragma unroll
                for (int r = 0; r < RP; r++) {
                    if (y + r >= Hh) break;
                    // 32-bit offset inside the image (fused images are below 2^30 pixels, zpx_api.cu)
                    uint8_t* o = outp + ((uint32_t)(y + r) * (uint32_t)W + (uint32_t)x) * 4u;
#pragma unroll
                    for (int g = 0; g < PXW / 4; g++) {
                        if (g > 0 && x + 4 * g >= W) break;
                        const uint32_t yw = *reinterpret_cast<const uint32_t*>(planeY + (row0 + r) * PYt + PXW * xg + 4 * g);
                        uint32_t p[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t yv = __byte_perm(yw, 0, 0x4440 + j);  // byte j, zero-extended: one PRMT
                            if (NC == 1) {
                                p[j] = yv * 0x010101u | 0xff000000u;  // .gray: (Y,Y,Y,255), color.zig:122-126
                            } else {
                                const int ci = (4 * g + j) / H;
                                p[j] = ycc_pixel((int)yv, rr[ci], gg[ci], bb[ci]);
                            }
                        }
                        if (vec_ok) {
                            __stcs(reinterpret_cast<uint4*>(o) + g, make_uint4(p[0], p[1], p[2], p[3]));
                        } else {
                            uint32_t* o32 = reinterpret_cast<uint32_t*>(o) + 4 * g;
                            const int xx = x + 4 * g;
                            __stcs(o32, p[0]);
                            if (xx + 1 < W) __stcs(o32 + 1, p[1]);
                            if (xx + 2 < W) __stcs(o32 + 2, p[2]);
                            if (xx + 3 < W) __stcs(o32 + 3, p[3]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        stage = stage + 1 == NS ? 0 : stage + 1;
    }
}

template <int H, int V, int NC, int NS, int NTMAX>
static cudaError_t launch_fused_ns(const K2Params& P, int sms, cudaStream_t s) {
    using Cfg = K2Cfg<H, V, NC>;
    const size_t smem = Cfg::smem_bytes(P.tmax, NS);
    cudaError_t e = cudaFuncSetAttribute(k2_fused<H, V, NC, NS, NTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nt = P.nt >= 96 && P.nt <= NTMAX ? P.nt : NTMAX;  // (>= 96: 32 threads per component stage the quantisers)
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k2_fused<H, V, NC, NS, NTMAX>, nt, smem);
    if (e != cudaSuccess) return e;
    const int grid = P.ntiles < sms * occ ? P.ntiles : sms * (occ > 0 ? occ : 1);
    k2_fused<H, V, NC, NS, NTMAX><<<grid, nt, smem, s>>>(P);
    return cudaGetLastError();
}

// three stages when all resident CTAs of that size still fit one SM (227 KB), else two
template <int H, int V, int NC>
static cudaError_t launch_fused_t(const K2Params& P, int sms, cudaStream_t s) {
    using Cfg = K2Cfg<H, V, NC>;
    constexpr int NT = (NC == 3 && H == 2 && V == 2) ? ZPX_K2_420_THREADS : 128;  // == k2_fused_threads(H, V, NC)
    if ((k2_threads_per_sm(NC) / NT) * (Cfg::smem_bytes(P.tmax, 3) + 1024) <= 227 * 1024) return launch_fused_ns<H, V, NC, 3, NT>(P, sms, s);
    return launch_fused_ns<H, V, NC, 2, NT>(P, sms, s);
}

int k2_fused_bpm(int h, int v, int nc) { return nc == 1 ? 1 : h * v + 2; }

cudaError_t k2_launch_fused(int h, int v, int nc, const K2Params& P, int sms, cudaStream_t s) {
    if (nc == 1) return launch_fused_t<1, 1, 1>(P, sms, s);
    switch (h << 4 | v) {
        case 0x11: return launch_fused_t<1, 1, 3>(P, sms, s);
        case 0x21: return launch_fused_t<2, 1, 3>(P, sms, s);
        case 0x22: return launch_fused_t<2, 2, 3>(P, sms, s);
        case 0x12: return launch_fused_t<1, 2, 3>(P, sms, s);
        case 0x41: return launch_fused_t<4, 1, 3>(P, sms, s);
        case 0x42: return launch_fused_t<4, 2, 3>(P, sms, s);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// generic (unfused) path
// ---------------------------------------------------------------------------
// One thread per block of one image; blockIdx.y = index into the image list.
// Writes 8x8 pixels into the component's native plane, exactly where reconstructBlock does.
__global__ void __launch_bounds__(128) k2g_idct_planes(const K2GParams P) {
    const ZpxImageDev* __restrict__ im = &P.imgs[P.list[blockIdx.y]];
    // quantisers of all components with columns 0/4 prescaled (idct.zig:100-101 folded in)
    __shared__ __align__(16) int qsm[4][64];
    for (int k = threadIdx.x; k < 64 * im->ncomp; k += blockDim.x) {
        const int cc = k >> 6, kk = k & 63;
        int q = P.quant[im->qidx[cc]].q[kk];
        if ((kk & 7) == 0 || (kk & 7) == 4) q <<= 11;
        qsm[cc][kk] = q;
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    for (; c < im->ncomp; c++) {
        const int cnt = im->comp_bw[c] * im->comp_bh[c];
        if (i < cnt) break;
        i -= cnt;
    }
    if (c >= im->ncomp) return;
    const int bw = im->comp_bw[c];
    const int by = i / bw, bx = i - by * bw;
    const int h = im->h[c], v = im->v[c];
    if (im->progressive) {
        // reconstructProgressiveImage (decoder.zig:1636-1661) only visits blocks that intersect the image
        const int sx = 8 * (im->hmax / h), sy = 8 * (im->vmax / v);
        if (bx * sx >= im->width || by * sy >= im->height) return;
    }
    uint64_t blk;
    if (im->layout == ZPX_LAYOUT_INTERLEAVED) {
        const int mx = bx / h, my = by / v;
        blk = im->coef_base + ((uint64_t)my * im->mxx + mx) * im->bpm + im->blk_off[c] + (by % v) * h + (bx % h);
    } else {
        blk = im->comp_base[c] + (uint64_t)by * bw + bx;
    }
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(P.coef) + blk * 8;
    const int key = bx & 7;
    uint32_t px[16];
    bool recon = (im->recon_mask >> c & 1u) != 0;
    // sequential frame, component without an interleaved scan: only blocks that touch the image were coded
    if (!im->progressive && !(im->recon_mask >> (4 + c) & 1u) && (bx * 8 >= im->width || by * 8 >= im->height)) recon = false;
    if (recon) {
        dequant_idct_block([&](int r) { return __ldg(src + (r ^ key)); }, qsm[c], px);
    } else {
        // no scan covers this component: the reference never reconstructs it and its plane keeps makeImg's zeros
#pragma unroll
        for (int k = 0; k < 16; k++) px[k] = 0;
    }
    uint8_t* dst = P.planes + im->plane_off[c] + ((size_t)by * 8) * im->plane_stride[c] + (size_t)bx * 8;
#pragma unroll
    for (int r = 0; r < 8; r++)
        *reinterpret_cast<uint2*>(dst + (size_t)r * im->plane_stride[c]) = make_uint2(px[2 * r], px[2 * r + 1]);
}

// One thread per pixel; blockIdx.y = index into the image list.  Image.rgbaPixels semantics per
// variant: image.zig:103-130 + color.zig toRGBA, decoder.zig:751-783 convertToRGB,
// :852-901 applyBlack (CMYK), :811-846 applyBlack (YCbCrK, intent = Go image/jpeg).
__global__ void __launch_bounds__(256) k2g_colour(const K2GParams P) {
    const ZpxImageDev* __restrict__ im = &P.imgs[P.list[blockIdx.y]];
    const int W = im->width, Hh = im->height;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)W * Hh) return;
    const int y = (int)(idx / W), x = (int)(idx - (size_t)y * W);
    const uint8_t* __restrict__ pl = P.planes;
    uint32_t px;
    const uint8_t Y = pl[im->plane_off[0] + (size_t)y * im->plane_stride[0] + x];
    if (im->mode == ZPX_MODE_GRAY) {
        px = (uint32_t)Y * 0x010101u | 0xff000000u;
    } else {
        const int hr = im->h[0] / im->h[1], vr = im->v[0] / im->v[1];
        if (im->mode == ZPX_MODE_CMYK) {
            // each plane t sub-sampled by >>1 iff its (h,v) differs from component 0's
            uint32_t v4[4];
            for (int t = 0; t < 4; t++) {
                const bool sub = im->h[t] != im->h[0] || im->v[t] != im->v[0];
                const int sx = sub ? x >> 1 : x, sy = sub ? y >> 1 : y;
                v4[t] = 255u - pl[im->plane_off[t] + (size_t)sy * im->plane_stride[t] + sx];
            }
            px = cmyk_pixel(v4[0], v4[1], v4[2], v4[3]);
        } else {
            const int cx = x / hr, cy = y / vr;
            const uint8_t Cb = pl[im->plane_off[1] + (size_t)cy * im->plane_stride[1] + cx];
            const uint8_t Cr = pl[im->plane_off[2] + (size_t)cy * im->plane_stride[2] + cx];
            if (im->mode == ZPX_MODE_RGB) {
                px = (uint32_t)Y | ((uint32_t)Cb << 8) | ((uint32_t)Cr << 16) | 0xff000000u;
            } else {
                int rr, gg, bb;
                chroma_terms(Cb, Cr, rr, gg, bb);
                px = ycc_pixel(Y, rr, gg, bb);
                if (im->mode == ZPX_MODE_YCCK) {
                    const uint32_t K = 255u - pl[im->plane_off[3] + (size_t)y * im->plane_stride[3] + x];
                    px = cmyk_pixel(px & 0xff, (px >> 8) & 0xff, (px >> 16) & 0xff, K);
                }
            }
        }
    }
    reinterpret_cast<uint32_t*>(P.out + im->out_off)[idx] = px;
}

// The reference's own return value for 4-component frames: Image{.CMYK} whose pixels are what applyBlack
// leaves (decoder.zig:852-901: 255 - plane, >>1 replication; :811-846 for YCbCrK: RGB of the YCbCr planes
// in C,M,Y and 255 - black in K).  rgbaPixels() of that image is what k2g_colour writes.
__global__ void __launch_bounds__(256) k2g_cmyk_native(const K2GParams P, const uint32_t img, uint8_t* __restrict__ dst) {
    const ZpxImageDev* __restrict__ im = &P.imgs[img];
    const int W = im->width, Hh = im->height;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)W * Hh) return;
    const int y = (int)(idx / W), x = (int)(idx - (size_t)y * W);
    const uint8_t* __restrict__ pl = P.planes;
    uint32_t px;
    if (im->mode == ZPX_MODE_CMYK) {
        px = 0;
        for (int t = 0; t < 4; t++) {
            const bool sub = im->h[t] != im->h[0] || im->v[t] != im->v[0];
            const int sx = sub ? x >> 1 : x, sy = sub ? y >> 1 : y;
            px |= (255u - pl[im->plane_off[t] + (size_t)sy * im->plane_stride[t] + sx]) << (8 * t);
        }
    } else {
        const int hr = im->h[0] / im->h[1], vr = im->v[0] / im->v[1];
        const int cx = x / hr, cy = y / vr;
        const uint8_t Y = pl[im->plane_off[0] + (size_t)y * im->plane_stride[0] + x];
        const uint8_t Cb = pl[im->plane_off[1] + (size_t)cy * im->plane_stride[1] + cx];
        const uint8_t Cr = pl[im->plane_off[2] + (size_t)cy * im->plane_stride[2] + cx];
        int rr, gg, bb;
        chroma_terms(Cb, Cr, rr, gg, bb);
        px = ycc_pixel(Y, rr, gg, bb) & 0x00ffffffu;
        px |= (255u - pl[im->plane_off[3] + (size_t)y * im->plane_stride[3] + x]) << 24;
    }
    reinterpret_cast<uint32_t*>(dst)[idx] = px;
}

cudaError_t k2g_launch_cmyk_native(const K2GParams& P, uint32_t img, size_t pixels, uint8_t* dst, cudaStream_t s) {
    if (pixels == 0) return cudaSuccess;
    k2g_cmyk_native<<<(unsigned)((pixels + 255) / 256), 256, 0, s>>>(P, img, dst);
    return cudaGetLastError();
}

// test hook: the device functions the colour phases above are made of (chroma_terms + ycc_pixel, cmyk_pixel) on
// free-standing samples.  YCbCrK: {Y, Cb, Cr, K plane byte} -> RGB of the YCbCr part, K = 255 - plane, then the .cmyk formula
__global__ void __launch_bounds__(256) k2_test_colour(const int mode, const uint8_t* __restrict__ in, const size_t n, uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t px;
    if (mode == ZPX_MODE_CMYK) {
        px = cmyk_pixel(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
    } else {
        const size_t st = mode == ZPX_MODE_YCBCR ? 3 : 4;
        int rr, gg, bb;
        chroma_terms(in[st * i + 1], in[st * i + 2], rr, gg, bb);
        px = ycc_pixel(in[st * i], rr, gg, bb);
        if (mode == ZPX_MODE_YCCK) px = cmyk_pixel(px & 0xff, (px >> 8) & 0xff, (px >> 16) & 0xff, 255u - in[4 * i + 3]);
    }
    out[i] = px;
}

cudaError_t k2_launch_test_colour(int mode, const uint8_t* samples, size_t n, uint8_t* rgba, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k2_test_colour<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mode, samples, n, reinterpret_cast<uint32_t*>(rgba));
    return cudaGetLastError();
}

// the IDCT half alone: native planes of the listed images (zpx_batch_fetch_native after a fused decode)
cudaError_t k2g_launch_planes(const K2GParams& P, int n_list, int max_blocks, cudaStream_t s) {
    if (n_list <= 0) return cudaSuccess;
    dim3 g1((max_blocks + 127) / 128, n_list);
    k2g_idct_planes<<<g1, 128, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t k2g_launch(const K2GParams& P, int n_list, int max_blocks, size_t max_pixels, cudaStream_t s) {
    if (n_list <= 0) return cudaSuccess;
    dim3 g1((max_blocks + 127) / 128, n_list);
    k2g_idct_planes<<<g1, 128, 0, s>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 g2((unsigned)((max_pixels + 255) / 256), n_list);
    k2g_colour<<<g2, 256, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace zpx
