/*
 * zpix_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of braheezy/zpix's JPEG decode path
 * (`jpeg.load` -> `Image` -> `Image.rgbaPixels()`), used as the parity
 * checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (libzpixcuda.so) never links, loads or calls it.
 *
 * Pinning status: the reference is Zig >= 0.15.1 and no Zig toolchain exists
 * in the build or GPU images, so the Zig binary itself could not be run.
 * This restatement is pinned against every test the reference holds for the
 * path (src/jpeg/decoder.zig:1843-2279: baseline/progressive plane equality
 * on 10 fixture pairs, "decode assorted", truncation -> UnexpectedEof, the
 * 504-byte fuzz input, the padded-RST image, the 10 bad-restart-marker
 * splices) and against the sha256 goldens of SURVEY.md Appendix C, which
 * come from an independent restatement.  ABSOLUTE pixel parity with the Zig
 * binary is therefore "unpinned by a reference run"; see DESIGN.md.
 */
#ifndef ZPIX_ORACLE_H
#define ZPIX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Image variants produced by jpeg.load (image.zig:24-33; only these four are
 * reachable from the JPEG decoder, decoder.zig:361-370). */
enum { ZO_GRAY = 0, ZO_YCBCR = 1, ZO_RGBA = 2, ZO_CMYK = 3 };

/* image.zig:465-472 */
enum { ZO_R444 = 0, ZO_R422 = 1, ZO_R420 = 2, ZO_R440 = 3, ZO_R411 = 4, ZO_R410 = 5 };

/* Error kinds = the Zig error names the decoder can return, in one enum.
 * 0 = success.  Names via zo_error_name(). */
enum {
    ZO_OK = 0,
    ZO_UnexpectedEof,
    ZO_InvalidSOIMarker,
    ZO_ShortSegmentLength,
    ZO_UnknownMarker,
    ZO_UnsupportedMarker,
    ZO_MissingSosMarker,
    ZO_MultipleSofMarkers,
    ZO_NumberComponents,
    ZO_Precision,
    ZO_SofWrongLength,
    ZO_RepeatedComponentIdentifier,
    ZO_BadTqValue,
    ZO_LumaChromaSubSamplingRatio,
    ZO_DriWrongLength,
    ZO_BadPqValue,
    ZO_DqtWrongLength,
    ZO_MissingFF00,
    ZO_UnsupportedColorModel,
    ZO_UninitializedHuffmanTable,
    ZO_BadHuffmanCode,
    ZO_DhtWrongLength,
    ZO_BadTcValue,
    ZO_BadThValue,
    ZO_HuffZeroLength,
    ZO_HuffTooLong,
    ZO_SosWrongLength,
    ZO_UnknownComponentSelector,
    ZO_BadTdValue,
    ZO_BadTaValue,
    ZO_SamplingFactorsTooLarge,
    ZO_BadSpectralSelection,
    ZO_ProgressiveACCoefficientsForMoreThanOneComponent,
    ZO_BadSuccessiveApproximation,
    ZO_ExcessiveDCComponent,
    ZO_UnexpectedHuffmanCode,
    ZO_TooManyCoefficients,
    ZO_BadRSTMarker,
    ZO_CreateImageFailed,
    ZO_UnsupportedComponent,
    ZO_InvalidImageType,
    ZO_ConfigOnly,
    ZO_OutOfMemory,
    ZO_ReferencePanics, /* input on which the Zig code hits a panic/unreachable/out-of-bounds */
    ZO_NUM_ERRORS
};

typedef struct zo_image {
    int32_t variant;        /* ZO_GRAY / ZO_YCBCR / ZO_RGBA / ZO_CMYK */
    int32_t width, height;  /* bounds() = (0,0)-(width,height) */
    /* ZO_GRAY: pix = gray plane, stride = 8*mxx.  ZO_RGBA / ZO_CMYK: pix =
     * interleaved 4 bytes/pixel, stride = 4*width.  ZO_YCBCR: y/cb/cr with
     * y_stride/c_stride, MCU-padded exactly like makeImg (decoder.zig:1755). */
    uint8_t *pixels;        /* the owning allocation (what Image.free releases) */
    size_t pixels_len;
    uint8_t *pix;
    size_t stride;
    uint8_t *y, *cb, *cr;
    size_t y_stride, c_stride;
    int32_t subsample_ratio; /* ZO_R4xx, only for ZO_YCBCR */
    int32_t ycck_intent;     /* 1 if produced by the YCbCrK branch whose reference code is broken
                                (SURVEY.md B2): output follows Go's image/jpeg, parity unpinned */
    int32_t eob_carry;       /* 1 if some scan started inside an End-Of-Band run the previous scan left open
                                (corrupt streams only; the GPU path answers UnsupportedStream for them) */
    int32_t pad0;
} zo_image;

/* eob_carry of the latest zo_decode / zo_decode_tap in this process, also when that call failed
 * (test aid, not thread-safe) */
int zo_last_eob_carry(void);
/* likewise: 1 if a coefficient left the int16 range during that decode (non-conforming streams; the GPU path
 * stores int16 coefficients and answers CoefficientOutOfRange) */
int zo_last_coef_overflow(void);

/* Optional coefficient tap, used by the tests of the entropy kernels.
 * Sequential frames: one record per coded block, in the order processSos
 * visits them (decoder.zig:1294-1427), coefficients BEFORE dequantisation, in
 * natural (de-zigzagged) order.  Progressive frames: after EOI, every block of
 * each component's progressive_coefficients grid in raster order. */
typedef struct zo_block_rec {
    int32_t comp, bx, by;
    int32_t coef[64];
} zo_block_rec;

typedef struct zo_tap {
    zo_block_rec *recs; /* caller-provided array, or NULL to only count */
    size_t cap;
    size_t count;       /* number of records produced (may exceed cap) */
} zo_tap;

typedef struct zo_config {
    uint32_t width, height;
    int32_t color_model; /* ZO_GRAY or ZO_YCBCR (4 components also report YCbCr, decoder.zig:210-215) */
} zo_config;

/* jpeg.loadFromBuffer / Decoder.decode (src/jpeg/root.zig:10, decoder.zig:155) */
int zo_decode(const uint8_t *data, size_t len, zo_image *out);
int zo_decode_tap(const uint8_t *data, size_t len, zo_image *out, zo_tap *tap);
/* Decoder.decodeConfig (decoder.zig:178) */
int zo_decode_config(const uint8_t *data, size_t len, zo_config *out);
/* Image.rgbaPixels (image.zig:103-130): writes width*height*4 bytes */
void zo_rgba_pixels(const zo_image *img, uint8_t *out);
/* Image.free (image.zig:68) */
void zo_free(zo_image *img);
const char *zo_error_name(int code);

/* idct.transform (idct.zig:77) on one block, in place */
void zo_idct(int32_t b[64]);
/* Color.toRGBA for .ycbcr / .cmyk (color.zig:90-121), 16-bit outputs */
void zo_ycbcr_to_rgba16(uint8_t y, uint8_t cb, uint8_t cr, uint32_t out[4]);
void zo_cmyk_to_rgba16(uint8_t c, uint8_t m, uint8_t y, uint8_t k, uint32_t out[4]);

/* Convenience for the CPU baseline: decode + rgbaPixels, discarding the image.
 * out may be NULL (a scratch buffer is used and freed). Returns error code. */
int zo_load_rgba(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, int32_t *w, int32_t *h);

/* Test helpers. reconstructBlock (decoder.zig:1553-1634) on n free-standing blocks: coef int32[n][64] natural
 * order, quant_zz 64 values in zig-zag order, out u8[n][64]. */
void zo_reconstruct_blocks(const int32_t *coef, const int32_t *quant_zz, size_t n, uint8_t *out);
/* Color.toRGBA (color.zig:90-121) + the >> 8 of rgbaPixels (image.zig:122-125) on n samples */
void zo_ycbcr_to_rgba8_batch(const uint8_t *ycc, size_t n, uint8_t *rgba);
void zo_cmyk_to_rgba8_batch(const uint8_t *cmyk, size_t n, uint8_t *rgba);
/* jpeg.load + rgbaPixels of data vs got: 0 identical, 1 different, < 0: -(error of the reference path) */
int zo_compare_rgba(const uint8_t *data, size_t len, const uint8_t *got, size_t got_len);

#ifdef __cplusplus
}
#endif
#endif
