// zpx_kernels.h -- kernel parameter blocks and host-callable launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_internal.h"

namespace zpx {

// ---- K1: entropy decode -----------------------------------------------------
struct K1Params {
    const uint8_t* blob;          // entropy-coded bytes of the whole device batch, still byte-stuffed (progressive scans
                                  // read it directly, zpx_k3.cu)
    int dbg;                      // experiments only (ZPX_K1_DBG): bit 0 = skip the coefficient stores
    const uint8_t* ublob;         // the same bytes with the stuffing removed (k0_unstuff), one 16-byte aligned run per
                                  // restart interval (ZpxIntervalDev::ustart / ulen): sequential scans and the
                                  // progressive ones of zpx_k3l.cu
    const ZpxIntervalDev* ivs;
    int n_iv;
    const ZpxScanDev* scans;
    const ZpxImageDev* imgs;
    const ZpxHuffDev* huff;
    uint4* coef;                  // 8 x uint4 per block
    unsigned long long* status;   // per image: smallest error key, ZPX_STATUS_NONE if none
    const uint32_t* eob_in;       // NULL, or (k3_progressive, the serial re-decode of zpx_api.cu: scans launched one by one in
    uint32_t* eob_out;            // file order) per image: the End-Of-Band run the previous scan left open / this one leaves
    uint32_t* img_flags;          // per image (status slot), zeroed before the decode: bit 0 = some coefficient of a
                                  // sequential scan lies outside [-4096, 4095] (the fused IDCT then takes the
                                  // reference's all-AC-zero row literally, zpx_idct.cuh)
};
// ---- K0: remove the byte stuffing (FF 00 -> FF) of the sequential scans' restart intervals, one warp per piece ----
cudaError_t k0_launch_unstuff(const uint8_t* blob, uint8_t* ublob, const ZpxSegDev* segs, int n_segs, cudaStream_t s);
cudaError_t k1_launch_lane_per_interval(const K1Params& P, int sm_count, cudaStream_t s);

// self-synchronising sub-sequence decoder (streams without DRI, or few large intervals)
struct K1SParams {
    K1Params k1;
    const ZpxWarpDev* warps;
    int n_warps;
    int n_iv;                 // intervals (domains), for the scan kernel
    unsigned long long* s_in;   // per sub-sequence: start state used by its latest decode
    unsigned long long* s_out;  //                   state at its end boundary
    int* s_n;                   //                   blocks started inside it (then: exclusive prefix)
    int4* s_dc;                 //                   DC difference sums per component (then: exclusive prefix)
    int* s_bad;                 //                   1: that decode skipped an invalid code (the write pass then probes
                                //                   the symbol after the lane's last block, see k1s_write)
    int* changed;               // device flag: an end state crossing a warp boundary changed in this sweep
    const int* gate;            // NULL, or: the launch has nothing to do unless this flag (the previous pass's `changed`) is set
};
cudaError_t k1s_launch_sync(const K1SParams& P, int sweep, cudaStream_t s);
cudaError_t k1s_launch_fix(const K1SParams& P, cudaStream_t s);
cudaError_t k1s_launch_scan(const K1SParams& P, cudaStream_t s);
cudaError_t k1s_launch_write(const K1SParams& P, cudaStream_t s);

// progressive scans (zpx_k3.cu): one lane per listed interval, read-modify-write of the coefficient grids
cudaError_t k3_launch_progressive(const K1Params& P, const uint32_t* list, int n_list, cudaStream_t s);
// progressive scans of frames with an ordinary successive-approximation script (zpx_k3l.cu): one LANE per listed
// interval.  list = the level's intervals grouped by pass type (DC first / AC first / AC refinement), each group
// padded to a multiple of 32 with 0xffffffff; DC refinement passes have a kernel of their own (one warp per interval)
cudaError_t k3l_launch_level(const K1Params& P, const uint32_t* list, int n_padded, cudaStream_t s);
cudaError_t k3l_launch_dc_refine(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s);
// AC refinement passes: before k3l_launch_level, the zero-position lists of every block; after it, the writes.  list =
// the level's AC refinement intervals, max_blocks = the most coded blocks any of them has
// AC first passes: after k3l_launch_level, the coefficient writes (list = the level's AC first-pass intervals)
cudaError_t k3l_launch_first_apply(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s);
cudaError_t k3l_launch_refine_prep(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s);
cudaError_t k3l_launch_refine_apply(const K1Params& P, const uint32_t* list, int n_list, uint32_t max_blocks, cudaStream_t s);

// ---- K2: fused dequant + IDCT + upsample + colour ------------------------------
// Threads per CTA of the fused kernel (= blocks a tile can hold) for a sampling; 512 / threads CTAs are resident
// per SM.  Measured on B200: 4:2:0 is 1 % faster with 256-thread CTAs (tiles of 40 MCUs fill 240 of them), every
// other sampling 4-9 % faster with 128 (4:2:2 3.69 -> 3.84 TB/s, gray + 4:4:4 512x512 3.30 -> 3.62 TB/s).
#ifndef ZPX_K2_420_THREADS
#define ZPX_K2_420_THREADS 256
#endif
static inline int k2_fused_threads(int h, int v, int nc) { return (nc == 3 && h == 2 && v == 2) ? ZPX_K2_420_THREADS : 128; }

struct K2Params {
    const int16_t* coef;
    uint8_t* out;     // RGBA (Image.rgbaPixels layout), or NULL: planes only
    uint8_t* planes;  // native planes with makeImg's strides (ZpxImageDev::plane_off / plane_stride), or NULL
    const ZpxImageDev* imgs;
    const ZpxTileDev* tiles;
    const ZpxQuantDev* quant;
    const uint32_t* img_flags;  // K1Params::img_flags
    int ntiles;
    int tmax;  // largest tile (MCUs): fixes the shared-memory layout
    int nt;    // threads per CTA: blocks of the largest tile rounded up to a warp (<= k2_fused_threads)
    int dense_only;  // 1: never take the sparse-block IDCT (ZPX_OPT_K2_DENSE, measurements and tests)
};
int k2_fused_bpm(int h, int v, int nc);
// persistent grid: min(tiles, SMs x resident CTAs)
cudaError_t k2_launch_fused(int h, int v, int nc, const K2Params& P, int sm_count, cudaStream_t s);

// ---- generic unfused path -------------------------------------------------------
struct K2GParams {
    const int16_t* coef;
    uint8_t* planes;
    uint8_t* out;
    const ZpxImageDev* imgs;
    const ZpxQuantDev* quant;
    const uint32_t* list;  // image indices taking this path
};
cudaError_t k2g_launch(const K2GParams& P, int n_list, int max_blocks, size_t max_pixels, cudaStream_t s);
cudaError_t k2g_launch_planes(const K2GParams& P, int n_list, int max_blocks, cudaStream_t s);  // planes only
// Image{.CMYK} pixels (applyBlack's result) of one 4-component image into dst (4*W*H bytes, device)
cudaError_t k2g_launch_cmyk_native(const K2GParams& P, uint32_t img, size_t pixels, uint8_t* dst, cudaStream_t s);

// test hook: the colour functions of the kernels above on n free-standing samples ({Y,Cb,Cr} / {C,M,Y,K} / {Y,Cb,Cr,K})
cudaError_t k2_launch_test_colour(int mode, const uint8_t* samples, size_t n, uint8_t* rgba, cudaStream_t s);

}  // namespace zpx
