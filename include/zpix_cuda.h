/*
 * zpix_cuda.h -- C ABI of libzpixcuda.so: batched JPEG decode on NVIDIA B200
 * behind braheezy/zpix's `jpeg.load` / `image.Image` API.
 *
 * C99-clean so that Zig can bind it with
 *     const c = @cImport(@cInclude("zpix_cuda.h"));
 * (fixed-width ints, POD structs, no bitfields, enums as int32_t constants).
 *
 * The reference has no FFI/plugin boundary of its own (it is 100 % Zig); the
 * interface this library slots under is the Zig module API.  Each entry point
 * below names the reference function(s) whose work it takes over:
 *
 *   zpx_batch_open        src/jpeg/decoder.zig:220-355  decodeInner marker loop,
 *                         :490-697 processSof/Dri/Dqt/App0/App14, :1026-1111 processDht,
 *                         :1148-1292 processSos header part, :1671-1705 findRst (host side)
 *   zpx_batch_info        src/jpeg/decoder.zig:178-218  decodeConfig
 *   zpx_batch_upload      (new) one host->device copy of the entropy-coded segments
 *   zpx_batch_decode      src/jpeg/decoder.zig:1294-1452 processSos MCU loop, :909-1134
 *                         decodeHuffman/receiveExtend, :1553-1634 reconstructBlock,
 *                         src/jpeg/idct.zig:77-201 transform, src/image/image.zig:103-130
 *                         rgbaPixels, src/color/color.zig:90-126 toRGBA,
 *                         decoder.zig:751-902 convertToRGB/applyBlack
 *   zpx_batch_fetch_rgba  src/image/image.zig:103-130 rgbaPixels (result hand-over)
 *   zpx_batch_fetch_native src/jpeg/decoder.zig:361-370 (the Image variant jpeg.load returns)
 *   zpx_decode_batch_rgba = open + upload + decode + fetch_rgba + close; what
 *                         `jpeg.decodeBatch` / `jpeg.loadBatch` (new, beside
 *                         src/jpeg/root.zig:10,36 loadFromBuffer/load) call.
 *
 * Ownership: the library never allocates or frees caller-visible host memory.
 * The caller (Zig: with its own allocator) allocates every output slice and
 * passes pointers down; the library owns device memory, streams and events
 * inside the opaque context.  zpx_host_alloc/zpx_host_free hand out pinned
 * host memory for callers that want the fast copy path (optional).
 *
 * Threading: a zpx_ctx is NOT thread-safe; use one per calling thread.
 * A context owns ONE set of device buffers: zpx_batch_upload makes that batch resident and evicts the
 * previous one (whose decode / fetch calls then return ZPX_E_BAD_STATE).  Open as many batches as you
 * like; keep one in the upload..fetch phase per context.
 * Errors: int32_t, 0 = ok.  No C++ exception and no abort crosses this ABI.
 * There is NO CPU fallback: every decode entry point fails with
 * ZPX_E_CUDA / ZPX_E_NO_DEVICE when no CUDA device is usable.
 */
#ifndef ZPIX_CUDA_H
#define ZPIX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZPX_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------
 * 1..ZPX_E_REF_LAST mirror the Zig error set of src/jpeg/decoder.zig one to
 * one (zpx_error_name() returns the Zig identifier, e.g. "UnexpectedEof").
 * Codes >= 100 are specific to this library. */
#define ZPX_OK 0
#define ZPX_E_UnexpectedEof 1
#define ZPX_E_InvalidSOIMarker 2
#define ZPX_E_ShortSegmentLength 3
#define ZPX_E_UnknownMarker 4
#define ZPX_E_UnsupportedMarker 5
#define ZPX_E_MissingSosMarker 6
#define ZPX_E_MultipleSofMarkers 7
#define ZPX_E_NumberComponents 8
#define ZPX_E_Precision 9
#define ZPX_E_SofWrongLength 10
#define ZPX_E_RepeatedComponentIdentifier 11
#define ZPX_E_BadTqValue 12
#define ZPX_E_LumaChromaSubSamplingRatio 13
#define ZPX_E_DriWrongLength 14
#define ZPX_E_BadPqValue 15
#define ZPX_E_DqtWrongLength 16
#define ZPX_E_MissingFF00 17
#define ZPX_E_UnsupportedColorModel 18
#define ZPX_E_UninitializedHuffmanTable 19
#define ZPX_E_BadHuffmanCode 20
#define ZPX_E_DhtWrongLength 21
#define ZPX_E_BadTcValue 22
#define ZPX_E_BadThValue 23
#define ZPX_E_HuffZeroLength 24
#define ZPX_E_HuffTooLong 25
#define ZPX_E_SosWrongLength 26
#define ZPX_E_UnknownComponentSelector 27
#define ZPX_E_BadTdValue 28
#define ZPX_E_BadTaValue 29
#define ZPX_E_SamplingFactorsTooLarge 30
#define ZPX_E_BadSpectralSelection 31
#define ZPX_E_ProgressiveACCoefficientsForMoreThanOneComponent 32
#define ZPX_E_BadSuccessiveApproximation 33
#define ZPX_E_ExcessiveDCComponent 34
#define ZPX_E_UnexpectedHuffmanCode 35
#define ZPX_E_TooManyCoefficients 36
#define ZPX_E_BadRSTMarker 37
#define ZPX_E_CreateImageFailed 38
#define ZPX_E_UnsupportedComponent 39
#define ZPX_E_InvalidImageType 40
#define ZPX_E_ConfigOnly 41
#define ZPX_E_OutOfMemory 42
#define ZPX_E_REF_LAST 42

#define ZPX_E_CUDA 100          /* a CUDA runtime call failed; see zpx_last_cuda_error */
#define ZPX_E_NO_DEVICE 101     /* no usable CUDA device (this library has no CPU path) */
#define ZPX_E_INVALID_ARG 102
#define ZPX_E_BAD_STATE 103     /* calls made in the wrong order for this batch */
#define ZPX_E_COEF_RANGE 104    /* a coefficient does not fit int16 (non-conforming 8-bit stream) */
#define ZPX_E_UNSUPPORTED_STREAM 105 /* reserved.  The kernels use it internally to flag a frame in which a scan ends
                                        inside an End-Of-Band run (the reference's eob_run survives scans,
                                        decoder.zig:144/:1451; corrupt files only); the status / fetch calls decode
                                        such frames again, scan by scan with the run carried, and report what the
                                        reference reports */
#define ZPX_E_MALFORMED_TABLE 106    /* DHT on which the reference itself panics (over-subscribed code) */

/* ---- image variants jpeg.load returns (src/image/image.zig:24-33) ------- */
#define ZPX_VARIANT_GRAY 0
#define ZPX_VARIANT_YCBCR 1
#define ZPX_VARIANT_RGBA 2
#define ZPX_VARIANT_CMYK 3

/* image.YCbCrSubsample (src/image/image.zig:465-472), same order */
#define ZPX_RATIO_444 0
#define ZPX_RATIO_422 1
#define ZPX_RATIO_420 2
#define ZPX_RATIO_440 3
#define ZPX_RATIO_411 4
#define ZPX_RATIO_410 5

typedef struct zpx_ctx zpx_ctx;
typedef struct zpx_batch zpx_batch;

/* What decodeConfig + makeImg would tell the caller about one image
 * (decoder.zig:178-218, 1708-1783).  All sizes in bytes. */
typedef struct zpx_image_info {
    int32_t status;          /* ZPX_OK or the header-parse error of this image */
    int32_t width;
    int32_t height;
    int32_t num_components;  /* 1, 3 or 4 */
    int32_t variant;         /* ZPX_VARIANT_* of the Image jpeg.load returns */
    int32_t subsample_ratio; /* ZPX_RATIO_* (variant YCBCR only) */
    int32_t progressive;     /* 1 for SOF2 */
    int32_t restart_interval;
    int32_t mxx, myy;        /* MCU grid */
    int32_t y_stride;        /* native YCbCr / Gray plane strides (MCU padded) */
    int32_t c_stride;
    int32_t device;          /* index into the context's device list this image was scheduled on */
    int32_t reserved;
    uint64_t rgba_len;       /* 4*width*height: length of Image.rgbaPixels() */
    uint64_t native_len;     /* length of the native variant's .pixels slice */
    uint64_t native_cb_off;  /* offsets of the Cb / Cr planes inside .pixels (YCBCR) */
    uint64_t native_cr_off;
} zpx_image_info;

/* Per-stage device times of the last zpx_batch_decode on one device, CUDA events. */
typedef struct zpx_timing {
    float h2d_ms;       /* zpx_batch_upload */
    float entropy_ms;   /* Huffman / run-length coefficient kernels */
    float idct_ms;      /* fused dequant + IDCT + upsample + colour kernels */
    float total_ms;     /* entropy_ms + idct_ms + anything between, first launch to last */
    float d2h_ms;       /* zpx_batch_fetch_* */
    int32_t entropy_launches;
    int32_t idct_launches;
    uint64_t entropy_bytes_in;   /* entropy-coded bytes consumed */
    uint64_t coef_bytes;         /* 128 * coded blocks */
    uint64_t rgba_bytes;         /* 4 * pixels */
    uint64_t pixels;
    uint64_t idct_fused_bytes;   /* (128 B + 4 P) of the images that took the fused kernel */
    float idct_fused_ms;         /* time of the fused kernel launches alone */
    int32_t images;
    int32_t images_failed;
} zpx_timing;

/* ---- context ------------------------------------------------------------ */
/* device_ids == NULL or n_devices <= 0: use device 0 only.  One worker
 * (stream set) per listed device; batches are partitioned across them. */
int32_t zpx_ctx_create(const int32_t *device_ids, int32_t n_devices, zpx_ctx **out);
void zpx_ctx_destroy(zpx_ctx *ctx);
int32_t zpx_ctx_num_devices(const zpx_ctx *ctx);
/* cudaError_t of the last failing CUDA call on this context (0 if none) and its text. */
int32_t zpx_last_cuda_error(const zpx_ctx *ctx);
const char *zpx_last_cuda_error_string(const zpx_ctx *ctx);

/* ---- batch: open -> upload -> decode -> fetch -> close -------------------- */
/* Host-only header parse of n JPEG byte buffers (kept by reference until
 * zpx_batch_upload returns).  A malformed image gets a per-image status and
 * does not fail the batch. */
int32_t zpx_batch_open(zpx_ctx *ctx, const uint8_t *const *bufs, const size_t *lens, int32_t n, zpx_batch **out);
int32_t zpx_batch_size(const zpx_batch *b);
int32_t zpx_batch_info(const zpx_batch *b, int32_t i, zpx_image_info *out);
/* Copy every image's entropy-coded segments (still byte-stuffed) and tables to its device: one H2D per device. */
int32_t zpx_batch_upload(zpx_batch *b);
/* Run the unstuffing + entropy + IDCT/colour kernels; results stay in device memory.
 * stream: a cudaStream_t to launch on (single-device contexts only), or NULL for the context's own streams.
 * With NULL the call returns after completion.  With a stream it returns once the work is enqueued, nothing is read
 * back (the self-synchronising entropy decoder enqueues a fixed number of synchronisation rounds, each gated on a
 * device flag; a stream that needs more -- adversarial data -- is found and decoded again by zpx_batch_status or
 * the fetch calls, so consumers of zpx_batch_device_rgba call zpx_batch_status first).
 * Ordering contract: work the caller enqueues on `stream` after this call sees the results.  The library's own
 * reads (zpx_batch_status, zpx_batch_fetch_*, zpx_batch_fetch_coefficients) run on the context's stream, which
 * the library orders after the decode's last kernel itself: they may be called right away, without synchronising
 * `stream`.  May be called repeatedly. */
int32_t zpx_batch_decode(zpx_batch *b, void *stream);
/* Copy RGBA (Image.rgbaPixels layout: tight rows unless out_stride says otherwise)
 * to caller memory.  out[i] may be NULL to skip image i.  out_stride may be NULL
 * (= 4*width).  status receives the final per-image status (may be NULL). */
int32_t zpx_batch_fetch_rgba(zpx_batch *b, uint8_t *const *out, const size_t *out_stride, int32_t *status);
/* Copy the native variant's .pixels buffer (exact layout of the Image jpeg.load
 * returns: MCU-padded Gray / planar YCbCr, or 4*W*H RGBA / CMYK).  Works after any decode: images that took the fused
 * kernel without ZPX_OPT_NATIVE_PLANES get their planes reconstructed now, from the coefficients still resident on the
 * device (one more IDCT pass; open the batch with ZPX_OPT_NATIVE_PLANES = 1 or 2 to have the fused kernel write them). */
int32_t zpx_batch_fetch_native(zpx_batch *b, uint8_t *const *out, int32_t *status);
/* Final per-image status after decode (header errors, device-detected entropy errors).  Waits for the decode.  Rare
 * corrupt frames (a scan that ends inside an End-Of-Band run) are decoded a second time here, scan by
 * scan, before their status is known: consumers of zpx_batch_device_rgba call this first.  (The fetch calls do.) */
int32_t zpx_batch_status(zpx_batch *b, int32_t *status);
/* Device pointer of image i's RGBA (GPU-resident hand-off, no D2H); NULL on error. */
const void *zpx_batch_device_rgba(const zpx_batch *b, int32_t i);
/* Test hook: copy image i's int16 coefficient blocks (128 B each, natural order,
 * un-swizzled, scan order for interleaved frames / per-component raster otherwise). */
int32_t zpx_batch_fetch_coefficients(zpx_batch *b, int32_t i, int16_t *out, size_t cap_blocks, size_t *n_blocks);
/* Test hooks: the reconstruction kernels (reconstructBlock decoder.zig:1553-1634, idct.zig:77-201, rgbaPixels
 * image.zig:103-130 + toRGBA color.zig:90-126, convertToRGB / applyBlack decoder.zig:751-902) without the entropy
 * stage.  zpx_batch_open_synthetic makes a batch of ONE sequential frame from its geometry alone: comp_hv[c] =
 * h << 4 | v, quant = ncomp x 64 values in zig-zag order as a DQT segment holds them (8- or 16-bit), mode 0 gray,
 * 1 YCbCr, 2 RGB-tagged, 3 CMYK, 4 YCbCrK.  After zpx_batch_upload, zpx_batch_set_coefficients injects the frame's
 * int16 blocks (natural order, the block order zpx_batch_fetch_coefficients returns); zpx_batch_decode then runs
 * the reconstruction kernels only and the fetch calls work as usual. */
int32_t zpx_batch_open_synthetic(zpx_ctx *ctx, int32_t width, int32_t height, int32_t ncomp, const uint8_t *comp_hv,
                                 const uint16_t *quant, int32_t mode, zpx_batch **out);
int32_t zpx_batch_set_coefficients(zpx_batch *b, int32_t i, const int16_t *blocks, size_t n_blocks);
/* The kernels' colour functions on n free-standing samples: mode 1 {Y,Cb,Cr} (3 bytes each), 3 {C,M,Y,K} as stored
 * in Image{.CMYK}, 4 {Y,Cb,Cr,K plane byte}; rgba receives 4 bytes per sample. */
int32_t zpx_test_colour(zpx_ctx *ctx, int32_t mode, const uint8_t *samples, size_t n, uint8_t *rgba);
int32_t zpx_batch_timing(const zpx_batch *b, int32_t device_index, zpx_timing *out);
void zpx_batch_close(zpx_batch *b);

/* One call: what jpeg.decodeBatch does.  out[i] must hold 4*w*h bytes
 * (sizes from zpx_probe / zpx_batch_info). */
int32_t zpx_decode_batch_rgba(zpx_ctx *ctx, const uint8_t *const *bufs, const size_t *lens, int32_t n,
                              uint8_t *const *out, const size_t *out_stride, int32_t *status);

/* The same for the value jpeg.load itself returns (decoder.zig:361-370): out[i] receives the native variant's
 * .pixels buffer, zpx_image_info.native_len bytes -- MCU-padded Gray / planar YCbCr (Y, Cb, Cr with makeImg's
 * strides), or 4*w*h RGBA / CMYK.  What `jpeg.loadBatch` binds when the caller wants Image{.YCbCr} like the
 * reference; 2.7x fewer bytes to bring back than RGBA for 4:2:0. */
int32_t zpx_decode_batch_native(zpx_ctx *ctx, const uint8_t *const *bufs, const size_t *lens, int32_t n,
                                uint8_t *const *out, int32_t *status);

/* Header-only probe of one buffer (decodeConfig, decoder.zig:178); no GPU needed. */
int32_t zpx_probe(const uint8_t *buf, size_t len, zpx_image_info *out);

/* Host-only view of the full header parse of one buffer (what zpx_batch_open computes), no GPU
 * needed: number of scans, total restart intervals, and the errors the reference raises only
 * after entropy-decoding part of the file (findRst / trailing marker-loop errors).  Test hook. */
typedef struct zpx_parse_report {
    int32_t status;        /* header-level error: the image is never sent to the GPU */
    int32_t n_scans;
    int32_t n_intervals;
    int32_t pending_err;   /* BadRSTMarker / UnexpectedEof found while locating restart intervals */
    int32_t pending_after_interval;
    int32_t trailing_err;  /* error of the marker loop after the last accepted scan */
    int32_t fused;         /* 1 if the image takes the fused IDCT/colour kernel */
    int32_t mode;          /* colour exit: 0 gray, 1 YCbCr, 2 RGB-tagged, 3 CMYK, 4 YCbCrK */
    uint64_t entropy_bytes;
    /* what the unstuffing pass (k0_unstuff) is handed */
    uint64_t stuffed_bytes;    /* FF 00 pairs inside the restart intervals */
    uint64_t unstuffed_bytes;  /* bytes of the intervals once the stuffing is removed */
    int32_t n_pieces;          /* work units of the unstuffing kernel */
    int32_t max_piece;         /* largest one, raw bytes */
    int32_t pieces_ok;         /* 1: the pieces tile every interval and none starts on the 0x00 of a pair */
    int32_t lane_script;       /* progressive frames: 1 = the scan script is an ordinary successive approximation (every
                                * band of a component coded once, then refined with falling Al): the frame can take the
                                * lane-per-scan kernels (ZPX_OPT_PROGRESSIVE_MODE) */
} zpx_parse_report;
int32_t zpx_parse_report_of(const uint8_t *buf, size_t len, zpx_image_info *info, zpx_parse_report *rep);

/* The scheduler's partition rule, host only (test hook): contiguous index ranges over n_devices,
 * balanced by weight (zpx_batch_open uses entropy-coded bytes + a constant per image).
 * weight 0 = image not scheduled (device_of = -1). */
int32_t zpx_partition(const uint64_t *weights, int32_t n, int32_t n_devices, int32_t *device_of);

/* ---- knobs (tests / benchmarks) ----------------------------------------- */
#define ZPX_OPT_ENTROPY_MODE 1  /* 0 auto, 1 lane-per-interval, 2 warp-per-interval/subsequence */
#define ZPX_OPT_FORCE_GENERIC 2 /* 1: always use the unfused planar IDCT + colour kernels */
#define ZPX_OPT_SUBSEQ_BYTES 3  /* sub-sequence size of the self-synchronising decoder */
#define ZPX_OPT_PIPELINE_CHUNK 4 /* images per chunk of zpx_decode_batch_rgba's copy/compute pipeline (0 = auto: about
                                    28 MB of compressed input per chunk; < 0 = off) */
#define ZPX_OPT_PIPELINE_RAMP 6  /* 1 (default): the pipeline's first two chunks are 1/4 and 1/2 of the chunk size */
#define ZPX_OPT_PIPELINE_WORKERS 7 /* host threads (each with its own device buffers) of that pipeline: 1..8, default 3 */
#define ZPX_OPT_NATIVE_PLANES 9 /* 0 (default): the decode writes native planes only for images on the unfused kernels
                                   (zpx_batch_fetch_native makes the others on demand);
                                   1: the fused kernel writes them too, beside the RGBA (zpx_batch_fetch_native works
                                   for every image); 2: planes only, no RGBA.  Read when a batch is opened. */
#define ZPX_OPT_PROGRESSIVE_MODE 10 /* 0 (default): progressive frames whose scan script is an ordinary successive
                                    * approximation decode one lane per scan (zpx_k3l.cu), the others one warp per scan
                                    * (zpx_k3.cu); 1: always one warp per scan */
#define ZPX_OPT_K2_DENSE 11 /* 1: the fused kernel never takes its sparse-block IDCT (warps whose 32 blocks have no
                             * coefficient outside the top-left 4x4 corner); results are identical either way */
#define ZPX_OPT_GATED_SWEEPS 12 /* 0..8 (-1: test hook), default 2: synchronisation rounds of the self-synchronising entropy decoder that
                                 * a decode on the caller's stream enqueues ahead of time (see zpx_batch_decode) */
#define ZPX_OPT_TEST_WIDE 8  /* test hook: 1 = frames filled by zpx_batch_set_coefficients take the kernels' exact
                                all-AC-zero-row IDCT variant whatever their coefficients */
int32_t zpx_ctx_set_option(zpx_ctx *ctx, int32_t option, int64_t value);

/* ---- misc ---------------------------------------------------------------- */
void *zpx_host_alloc(size_t bytes); /* pinned host memory (cudaHostAlloc); NULL on failure */
void zpx_host_free(void *p);
const char *zpx_error_name(int32_t code); /* Zig error identifier for 1..ZPX_E_REF_LAST */
int32_t zpx_abi_version(void);
/* Number of kernels this library launched since the context was created (all devices). */
uint64_t zpx_ctx_kernel_launches(const zpx_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ZPIX_CUDA_H */
