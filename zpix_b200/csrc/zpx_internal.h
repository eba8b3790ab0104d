// zpx_internal.h -- structures shared by the host parser, the scheduler and the kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/zpix_cuda.h"

#define ZPX_MAX_COMP 4
#define ZPX_MAX_BLK_PER_MCU 16   // interleaved scans are limited to 10 (decoder.zig:1216); non-interleaved use h*v <= 16
#define ZPX_LUT_BITS 10
#define ZPX_LUT_SIZE (1 << ZPX_LUT_BITS)

// ---------------------------------------------------------------------------
// device-visible descriptors (POD, identical layout on host and device)
// ---------------------------------------------------------------------------

// One Huffman table in device form.  Built on the host from the DHT payload
// (decoder.zig:1026-1111).  lut: first ZPX_LUT_BITS bits -> (symbol << 8) | code length, 0 = longer
// code or invalid.  Longer codes: smallest l with v16 < limit[l] (v16 = next 16 bits, left aligned),
// symbol = vals[valoff[l] + (v16 >> (16 - l))]  -- the canonical decode the reference does bit by
// bit at decoder.zig:946-969; identical results for every code of the table.
// fast[]: everything the per-symbol step needs in one 32-bit load --
// ZpxHuffDev::fast entry, one per ZPX_LUT_BITS-bit prefix (byte-aligned fields: one PRMT each on the device):
//   byte 0  total bits of the symbol (code + value bits)
//   byte 1  code length
//   byte 2  32 - value bits (the shift that right-aligns the value; 32 = no value bits)
//   byte 3  zig-zag advance (DC: 1; AC: run + 1, ZRL 16, EOB 64) | special << 7
//   special: the symbol takes the kernels' rare path -- AC (r, 0) with 0 < r < 15 (End-Of-Band run),
//   an AC value of 13 or more bits (a coefficient outside [-4096, 4095]: the kernels flag the image for the exact
//   IDCT rows, zpx_idct.cuh) or a DC category > 16
//   0 = code longer than ZPX_LUT_BITS or invalid -> canonical search.
#define ZPX_FE(tot, len, size, adv, special) \
    ((uint32_t)(tot) | (uint32_t)(len) << 8 | (uint32_t)(32u - (size)) << 16 | (uint32_t)(adv) << 24 | (uint32_t)(special) << 31)
// what the rare path hands back for an End-Of-Band run: byte 2 holds the run bits r, byte 3 = 64 | special
#define ZPX_FE_RUN(tot, len, rr) ((uint32_t)(tot) | (uint32_t)(len) << 8 | (uint32_t)(rr) << 16 | 64u << 24 | 1u << 31)

struct ZpxHuffDev {
    uint32_t fast[ZPX_LUT_SIZE];
    uint16_t lut[ZPX_LUT_SIZE];
    uint32_t limit[17];  // index 1..16
    int32_t valoff[17];  // index 1..16
    uint8_t vals[256];
    uint32_t defined;    // num_codes != 0
    uint32_t pad[3];
};

// Quantisation table, natural (row-major) order, as int32 so the kernels multiply directly.
// decoder.zig:629-666 stores zig-zag order; de-zigzagging is a pure permutation.
struct ZpxQuantDev {
    int32_t q[64];
};

// colour exit of an image (decoder.zig:361-370, 699-709, 792-902)
enum { ZPX_MODE_GRAY = 0, ZPX_MODE_YCBCR = 1, ZPX_MODE_RGB = 2, ZPX_MODE_CMYK = 3, ZPX_MODE_YCCK = 4 };
// coefficient layout of an image in HBM
enum { ZPX_LAYOUT_INTERLEAVED = 0, ZPX_LAYOUT_PLANAR = 1 };

struct ZpxImageDev {
    int32_t width, height;
    int32_t mxx, myy;
    int32_t ncomp;
    int32_t mode;      // ZPX_MODE_*
    int32_t layout;    // ZPX_LAYOUT_*
    int32_t bpm;       // blocks per MCU in the interleaved layout
    int32_t progressive;
    int32_t fused;     // 1: eligible for the fused kernel
    int32_t hmax, vmax;  // = h[0], v[0]
    uint8_t h[4], v[4];
    int32_t qidx[4];          // index into the quant table array, per component
    uint32_t blk_off[4];      // interleaved: index of the component's first block inside an MCU
    uint64_t coef_base;       // first block of the image in the coefficient buffer (block units)
    uint64_t pmask_base;      // progressive frames on the lane-per-interval kernels (zpx_k3l.cu): where the per-block
                              // non-zero / sign maps start in the coefficient buffer (block units; 16 bytes per block,
                              // indexed like the coefficient blocks relative to coef_base); else 0
    uint64_t ppos_base;       // the same frames: block start positions found by the serial part of the AC refinement
                              // passes (32-bit entries, ZpxScanDev::pos_off) ...
    uint64_t pzl_base;        // ... and the zero-position lists they read (80-byte records, ZpxScanDev::zl_off)
    uint64_t comp_base[4];    // planar: first block of each component's grid (block units, absolute)
    int32_t comp_bw[4];       // planar: grid width in blocks = mxx*h
    int32_t comp_bh[4];       //         grid height in blocks = myy*v
    uint64_t out_off;         // RGBA output, byte offset in the device output buffer
    uint64_t plane_off[4];    // native planes (generic path / native output): byte offsets in plane buffer
    int32_t plane_stride[4];
    int32_t plane_rows[4];
    uint32_t status_slot;     // index into the device status array
    uint32_t recon_mask;      // bit 4+c: (sequential frames) some interleaved scan codes component c, i.e. every block
                              // of its grid is reconstructed, not only those that touch the image;
                              // bit c: component c is reconstructed (some scan covers it); the planes of the other
                              // components keep makeImg's zero fill (decoder.zig:1644, and reconstructBlock is
                              // only reached from scans)
};

// One scan (SOS) of an image, device form.
struct ZpxScanDev {
    uint32_t img;        // image index on this device
    int32_t ncomp;       // components in scan
    int32_t nblk;        // coded blocks per MCU iteration (interleaved) ; h*v iterations for non-interleaved
    int32_t interleaved; // ncomp > 1
    int32_t ss, se, ah, al;
    int32_t total_mcu;   // mxx*myy
    int32_t restart_interval;
    int32_t scan_index;  // ordinal of the scan inside the image (error ordering)
    int32_t cw, ch;      // non-interleaved: coded blocks per row / rows = blocks intersecting the image
    // 1: every block of the MCU uses the same DC and the same AC table and each component has one block:
    // parsing does not depend on the block phase, so the self-synchronising decoder leaves it out of
    // its state and attributes DC sums by phase relative to the sub-sequence start (zpx_k1s.cu)
    int32_t rotate;
    uint32_t pos_off, zl_off;  // AC refinement scans on the lane-per-interval kernels: the scan's first entry / record
                               // in the image's position and list areas (ZpxImageDev::ppos_base, pzl_base)
    int32_t pad1;
    // per block inside one MCU of this scan
    uint8_t blk_comp[ZPX_MAX_BLK_PER_MCU]; // frame component index
    uint8_t blk_hx[ZPX_MAX_BLK_PER_MCU];
    uint8_t blk_vy[ZPX_MAX_BLK_PER_MCU];
    uint8_t blk_slot[ZPX_MAX_BLK_PER_MCU]; // block's index inside the image's interleaved MCU (layout)
    uint16_t blk_dc[ZPX_MAX_BLK_PER_MCU];  // Huffman table indices (into the device table array)
    uint16_t blk_ac[ZPX_MAX_BLK_PER_MCU];
    // the same, packed for one 16-byte load per block:
    //   x = DC table index, y = AC table index,
    //   z = comp | hx << 8 | vy << 16 | slot << 24,
    //   w = h | v << 8 | (DC table undefined) << 16 | (AC table undefined) << 17 |
    //       (a later scan of the frame codes this component again: decode, do not store) << 18
    alignas(16) uint32_t blk_pack[ZPX_MAX_BLK_PER_MCU][4];
};

// One restart interval (or the whole scan when DRI == 0): the unit the
// entropy kernels parallelise over.
struct ZpxIntervalDev {
    uint64_t start;      // byte offset of the first entropy-coded byte in the device blob
    uint32_t len;        // bytes up to the limit (next marker or end of file)
    uint32_t scan;       // index into the scan array
    uint32_t first_mcu;  // MCU iteration index (my*mxx+mx) of the interval's first MCU
    uint32_t n_mcu;      // MCU iterations in this interval
    uint32_t ordinal;    // interval index inside the scan (error ordering)
    uint32_t flags;      // bit0: limit is the end of the file (UnexpectedEof instead of MissingFF00)
                         // bit1: last interval of a scan that is not the image's last scan
    uint32_t first_block;  // ordinal (inside the scan) of the interval's first coded block
    uint32_t n_blocks;     // coded blocks in this interval
    // self-synchronising mode: the interval is cut into nsub sub-sequences of sub_bytes UNSTUFFED bytes,
    // boundaries at multiples of sub_bytes from ustart; their state lives at sub_first + i
    uint32_t sub_first;
    uint32_t nsub;
    uint32_t sub_bytes;
    // sequential scans: the interval's bytes with the stuffing removed (FF 00 -> FF; k0_unstuff, zpx_k0.cu) live at
    // ustart (16-byte aligned) in the unstuffed blob, ulen bytes, zeros up to the next 16-byte boundary
    uint32_t ulen;
    uint64_t ustart;
};

// One unit of work of k0_unstuff: a run of raw bytes (about ZPX_SEG_BYTES) of one interval that does not split an FF 00 pair
struct ZpxSegDev {
    uint64_t src;    // byte offset in the raw blob
    uint64_t dst;    // byte offset in the unstuffed blob
    uint32_t len;    // raw bytes
    uint32_t flags;  // bit 0: last segment of its interval (zero fill up to the next 16-byte boundary)
};

// one warp of the self-synchronising decoder: 32 consecutive sub-sequences of one interval
struct ZpxWarpDev {
    uint32_t iv;     // interval index
    uint32_t first;  // index (inside the interval) of lane 0's sub-sequence
};

// K2 tile: a run of MCUs inside one MCU row of one image.
struct alignas(16) ZpxTileDev {  // (one 128-bit load)
    uint32_t img;
    uint16_t my;
    uint16_t n;      // MCUs in tile
    uint32_t mx0;
    uint32_t pad;
};

// device error record: smaller key = earlier in the reference's decode order
// key = scan_index << 48 | block ordinal << 8 | code
#define ZPX_STATUS_NONE 0xFFFFFFFFFFFFFFFFull

// ---------------------------------------------------------------------------
// host-side parse results
// ---------------------------------------------------------------------------
struct ZpxHuffHost {
    bool defined = false;
    uint8_t counts[16] = {0};
    uint8_t vals[256] = {0};
    int num_codes = 0;
};

struct ZpxIntervalHost {
    size_t start;   // offset in the source file
    size_t limit;   // offset of the limit (first 0xFF followed by a byte != 0x00, or file end)
    uint32_t first_mcu, n_mcu;
    bool eof_limit;
    uint32_t n_stuffed = 0;              // FF 00 pairs inside [start, limit): unstuffed length = limit - start - n_stuffed
    uint32_t seg_first = 0, n_segs = 0;  // its pieces in ZpxScanHost::segs
};

// a piece of an interval for the unstuffing kernel: [src, src + len) of the file holds whole FF 00 pairs only
#define ZPX_SEG_BYTES 16384
struct ZpxSegHost {
    size_t src;      // offset in the source file
    uint32_t len;    // raw bytes
    uint32_t uoff;   // unstuffed bytes of the interval before this piece
};

struct ZpxScanHost {
    int ncomp = 0;
    int comp[ZPX_MAX_COMP] = {0};  // frame component index per scan component
    int td[ZPX_MAX_COMP] = {0}, ta[ZPX_MAX_COMP] = {0};
    int ss = 0, se = 63, ah = 0, al = 0;
    int restart_interval = 0;
    ZpxHuffHost dc[ZPX_MAX_COMP], ac[ZPX_MAX_COMP];  // snapshot of the tables this scan uses
    int32_t quant[ZPX_MAX_COMP][64];                 // snapshot of quant[tq] (zig-zag order) per scan component
    std::vector<ZpxIntervalHost> intervals;
    std::vector<ZpxSegHost> segs;
    // error the reference raises after `err_after_interval` intervals decoded fine (findRst / EOF)
    int pending_err = 0;
    int err_after_interval = -1;
};

struct ZpxParsed {
    int status = 0;          // header-level error (image is not decoded at all)
    int width = 0, height = 0;
    int ncomp = 0;
    int h[ZPX_MAX_COMP] = {0}, v[ZPX_MAX_COMP] = {0}, tq[ZPX_MAX_COMP] = {0};
    uint8_t cid[ZPX_MAX_COMP] = {0};
    bool baseline = false, progressive = false;
    bool jfif = false, adobe_valid = false;
    int adobe_transform = 0;
    int mxx = 0, myy = 0;
    std::vector<ZpxScanHost> scans;
    int32_t final_quant[4][64];  // quant tables as they stand at EOI (progressive reconstruct)
    bool saw_sos = false;
    // error raised by the marker loop AFTER some scans were accepted (e.g. missing EOI):
    // decoding on the device still happens for earlier errors, this one wins otherwise.
    int trailing_err = 0;
    // derived
    int mode = 0;      // ZPX_MODE_*
    int variant = 0;   // ZPX_VARIANT_*
    int ratio = 0;     // ZPX_RATIO_*
};

// parse one JPEG byte buffer the way decodeInner walks it (decoder.zig:220-373), without
// decoding any entropy-coded data.  config_only mirrors decodeConfig.
void zpx_parse_jpeg(const uint8_t* data, size_t len, bool config_only, ZpxParsed* out);
void zpx_build_huff_dev(const ZpxHuffHost& h, bool is_ac, ZpxHuffDev* out, int* malformed);
uint32_t zpx_fast_entry(bool is_ac, int len, int sym);
void zpx_fill_info(const ZpxParsed& p, zpx_image_info* info);
// colour exit / Image variant / sub-sampling ratio from the frame header fields (decoder.zig:361-370, 699-709)
void zpx_derive(ZpxParsed* o);

extern const uint8_t zpx_unzig[64];
