/*
 * zpix_oracle.c -- TEST INFRASTRUCTURE ONLY (see zpix_oracle.h).
 *
 * Sequential, literal CPU restatement of braheezy/zpix's JPEG decode path.
 * Every function cites the reference lines it follows (paths relative to the
 * reference repo).  No cleverness on purpose: byte-at-a-time stuffed reader,
 * 8-bit LUT + bit-serial Huffman, per-block reconstructBlock, per-pixel
 * rgbaPixels.  Integer arithmetic is wrapping 32-bit (the reference's `+ - *`
 * trap on overflow in safe builds, so behaviour is only defined where nothing
 * overflows; `<<` wraps in both).
 */
#include "zpix_oracle.h"

#include <stdlib.h>
#include <string.h>

#define MAX_COMPONENTS 4
#define MAX_TC 1
#define MAX_TH 3
#define MAX_TQ 3
#define BLOCK_SIZE 64
#define LUT_SIZE 8
#define MAX_CODE_LENGTH 16
#define MAX_NUM_CODES 256

#define TRY(expr)            \
    do {                     \
        int e__ = (expr);    \
        if (e__ != ZO_OK)    \
            return e__;      \
    } while (0)

/* wrapping i32 helpers */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wshl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
/* arithmetic shift right (gcc implements >> on signed as arithmetic) */
static inline int32_t asr(int32_t a, int s) { return a >> s; }

/* decoder.zig:73-82 */
static const uint8_t unzig[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
};

/* HuffTable.zig */
typedef struct {
    int32_t num_codes;
    uint16_t lut[1 << LUT_SIZE];
    uint8_t vals[MAX_NUM_CODES];
    int32_t min_codes[MAX_CODE_LENGTH];
    int32_t max_codes[MAX_CODE_LENGTH];
    int32_t vals_indices[MAX_CODE_LENGTH];
} huff_table;

typedef struct {
    int32_t h, v;
    uint8_t id, tq;
} component;

typedef struct {
    /* input stream (stands in for *std.Io.Reader) */
    const uint8_t *src;
    size_t src_len, src_pos;

    /* decoder.zig:90-97 */
    struct { uint32_t a, m; int32_t n; } bits;
    /* decoder.zig:107-116 */
    struct { uint8_t buffer[4096]; size_t i, j, num_unreadable; } bytes;

    uint32_t width, height;

    /* destination image data (decoder.zig:121-127) */
    int have_gray, have_ycbcr;
    uint8_t *gray_pixels; size_t gray_len, gray_stride;          /* gray_img */
    uint8_t *ycc_pixels; size_t ycc_len;                          /* ycbcr_img.pixels */
    uint8_t *ycc_y, *ycc_cb, *ycc_cr; size_t y_stride, c_stride; int subsample_ratio;
    uint8_t *black_pixels; size_t black_len, black_stride;

    uint16_t restart_interval;
    uint8_t num_components;
    int baseline, progressive;
    int jfif, adobe_transform_valid;
    int adobe_transform; /* 0 unknown, 1 y_cb_cr, 2 y_cb_cr_k */
    uint16_t eob_run;
    int eob_carry; /* test aid: a scan started with eob_run != 0 (left open by the previous scan) */
    int coef_overflow; /* test aid: a coefficient left the int16 range (the GPU path stores int16) */

    component comp[MAX_COMPONENTS];
    int32_t (*prog_coef[MAX_COMPONENTS])[BLOCK_SIZE];
    size_t prog_len[MAX_COMPONENTS];
    huff_table huff[MAX_TC + 1][MAX_TH + 1];
    int32_t quant[MAX_TQ + 1][BLOCK_SIZE]; /* zig-zag order */
    uint8_t tmp[2 * BLOCK_SIZE];

    zo_tap *tap;
} decoder;

static const char *const error_names[ZO_NUM_ERRORS] = {
    "ok", "UnexpectedEof", "InvalidSOIMarker", "ShortSegmentLength", "UnknownMarker", "UnsupportedMarker",
    "MissingSosMarker", "MultipleSofMarkers", "NumberComponents", "Precision", "SofWrongLength",
    "RepeatedComponentIdentifier", "BadTqValue", "LumaChromaSubSamplingRatio", "DriWrongLength", "BadPqValue",
    "DqtWrongLength", "MissingFF00", "UnsupportedColorModel", "UninitializedHuffmanTable", "BadHuffmanCode",
    "DhtWrongLength", "BadTcValue", "BadThValue", "HuffZeroLength", "HuffTooLong", "SosWrongLength",
    "UnknownComponentSelector", "BadTdValue", "BadTaValue", "SamplingFactorsTooLarge", "BadSpectralSelection",
    "ProgressiveACCoefficientsForMoreThanOneComponent", "BadSuccessiveApproximation", "ExcessiveDCComponent",
    "UnexpectedHuffmanCode", "TooManyCoefficients", "BadRSTMarker", "CreateImageFailed", "UnsupportedComponent",
    "InvalidImageType", "ConfigOnly", "OutOfMemory", "ReferencePanics",
};

const char *zo_error_name(int code) {
    if (code < 0 || code >= ZO_NUM_ERRORS) return "?";
    return error_names[code];
}

/* ------------------------------------------------------------------ */
/* byte layer                                                          */
/* ------------------------------------------------------------------ */

/* decoder.zig:447-472.  readSliceShort on a fixed/file reader: copy what is
 * available, 0 bytes at end of stream -> UnexpectedEof. */
static int fill(decoder *d) {
    if (d->bytes.i != d->bytes.j) return ZO_ReferencePanics;
    if (d->bytes.j > 2) {
        d->bytes.buffer[0] = d->bytes.buffer[d->bytes.j - 2];
        d->bytes.buffer[1] = d->bytes.buffer[d->bytes.j - 1];
        d->bytes.i = 2;
        d->bytes.j = 2;
    } else {
        d->bytes.i = 0;
        d->bytes.j = 0;
    }
    size_t room = sizeof(d->bytes.buffer) - d->bytes.j;
    size_t avail = d->src_len - d->src_pos;
    size_t n = avail < room ? avail : room;
    memcpy(d->bytes.buffer + d->bytes.j, d->src + d->src_pos, n);
    d->src_pos += n;
    d->bytes.j += n;
    if (n == 0) return ZO_UnexpectedEof;
    return ZO_OK;
}

/* decoder.zig:479-487 */
static void unread_byte_stuffed_byte(decoder *d) {
    d->bytes.i -= d->bytes.num_unreadable;
    d->bytes.num_unreadable = 0;
    if (d->bits.n >= 8) {
        d->bits.a >>= 8;
        d->bits.n -= 8;
        d->bits.m >>= 8;
    }
}

/* decoder.zig:402-410 */
static int read_byte(decoder *d, uint8_t *out) {
    while (d->bytes.i == d->bytes.j) TRY(fill(d));
    *out = d->bytes.buffer[d->bytes.i];
    d->bytes.i += 1;
    d->bytes.num_unreadable = 0;
    return ZO_OK;
}

/* decoder.zig:414-443 */
static int read_full(decoder *d, uint8_t *p, size_t len) {
    size_t offset = 0;
    if (d->bytes.num_unreadable > 0) {
        if (d->bits.n >= 8) unread_byte_stuffed_byte(d);
        d->bytes.num_unreadable = 0;
    }
    while (offset < len) {
        size_t available = d->bytes.j - d->bytes.i;
        size_t to_copy = available < len - offset ? available : len - offset;
        memcpy(p + offset, d->bytes.buffer + d->bytes.i, to_copy);
        d->bytes.i += to_copy;
        offset += to_copy;
        if (offset == len) break;
        TRY(fill(d));
    }
    return ZO_OK;
}

/* decoder.zig:376-398 */
static int ignore(decoder *d, int32_t n) {
    int32_t local_n = n;
    if (d->bytes.num_unreadable > 0) {
        if (d->bits.n >= 8) unread_byte_stuffed_byte(d);
        d->bytes.num_unreadable = 0;
    }
    for (;;) {
        size_t remaining = d->bytes.j - d->bytes.i;
        if (remaining > (size_t)local_n) remaining = (size_t)local_n;
        d->bytes.i += remaining;
        local_n -= (int32_t)remaining;
        if (local_n == 0) break;
        TRY(fill(d));
    }
    return ZO_OK;
}

/* decoder.zig:712-749 */
static int read_byte_stuffed_byte(decoder *d, uint8_t *out) {
    if (d->bytes.i + 2 <= d->bytes.j) {
        uint8_t x = d->bytes.buffer[d->bytes.i];
        d->bytes.i += 1;
        d->bytes.num_unreadable = 1;
        if (x != 0xff) { *out = x; return ZO_OK; }
        if (d->bytes.buffer[d->bytes.i] != 0x00) return ZO_MissingFF00;
        d->bytes.i += 1;
        d->bytes.num_unreadable = 2;
        *out = 0xff;
        return ZO_OK;
    }
    d->bytes.num_unreadable = 0;
    uint8_t x;
    TRY(read_byte(d, &x));
    d->bytes.num_unreadable = 1;
    if (x != 0xff) { *out = x; return ZO_OK; }
    TRY(read_byte(d, &x));
    d->bytes.num_unreadable = 2;
    if (x != 0x00) return ZO_MissingFF00;
    *out = 0xff;
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* bit layer / Huffman                                                 */
/* ------------------------------------------------------------------ */

/* decoder.zig:975-991 */
static int ensure_n_bits(decoder *d, int32_t n) {
    for (;;) {
        uint8_t c;
        TRY(read_byte_stuffed_byte(d, &c));
        d->bits.a = (d->bits.a << 8) | (uint32_t)c;
        d->bits.n += 8;
        if (d->bits.m == 0) d->bits.m = 1u << 7;
        else d->bits.m <<= 8;
        if (d->bits.n >= n) break;
    }
    return ZO_OK;
}

/* decoder.zig:1115-1134 */
static int receive_extend(decoder *d, uint8_t bit_count, int32_t *out) {
    if (d->bits.n < (int32_t)bit_count) TRY(ensure_n_bits(d, (int32_t)bit_count));
    d->bits.n -= (int32_t)bit_count;
    d->bits.m >>= bit_count;
    int32_t threshold = wshl(1, bit_count);
    int32_t value = (int32_t)((d->bits.a >> d->bits.n) & (uint32_t)(threshold - 1));
    if (value < (threshold >> 1)) value += wshl(-1, bit_count) + 1;
    *out = value;
    return ZO_OK;
}

/* decoder.zig:994-1006 */
static int decode_bit(decoder *d, int *out) {
    if (d->bits.n == 0) TRY(ensure_n_bits(d, 1));
    *out = (d->bits.a & d->bits.m) != 0;
    d->bits.n -= 1;
    d->bits.m >>= 1;
    return ZO_OK;
}

/* decoder.zig:1009-1022 */
static int decode_bits(decoder *d, int32_t n, uint32_t *out) {
    if (d->bits.n < n) TRY(ensure_n_bits(d, n));
    uint32_t ret = d->bits.a >> (d->bits.n - n);
    ret &= (1u << n) - 1;
    d->bits.n -= n;
    d->bits.m >>= n;
    *out = ret;
    return ZO_OK;
}

/* decoder.zig:909-970 */
static int decode_huffman(decoder *d, huff_table *h, uint8_t *out) {
    if (h->num_codes == 0) return ZO_UninitializedHuffmanTable;
    int slow = 0;
    if (d->bits.n < 8) {
        int e = ensure_n_bits(d, 8);
        if (e != ZO_OK) {
            if (e != ZO_MissingFF00) return e; /* ShortHuffmanData is never raised */
            if (d->bytes.num_unreadable != 0) unread_byte_stuffed_byte(d);
            slow = 1;
        }
    }
    if (!slow) {
        uint16_t v = h->lut[(d->bits.a >> (d->bits.n - LUT_SIZE)) & 0xff];
        if (v != 0) {
            int32_t bit_count = (int32_t)(v & 0xff) - 1;
            d->bits.n -= bit_count;
            d->bits.m >>= bit_count;
            *out = (uint8_t)(v >> 8);
            return ZO_OK;
        }
    }
    int32_t code = 0;
    for (int i = 0; i < MAX_CODE_LENGTH; i++) {
        if (d->bits.n == 0) TRY(ensure_n_bits(d, 1));
        if ((d->bits.a & d->bits.m) != 0) code |= 1;
        d->bits.n -= 1;
        d->bits.m >>= 1;
        if (code <= h->max_codes[i]) {
            int32_t idx = h->vals_indices[i] + code - h->min_codes[i];
            if (idx < 0 || idx >= MAX_NUM_CODES) return ZO_ReferencePanics;
            *out = h->vals[idx];
            return ZO_OK;
        }
        code <<= 1;
    }
    return ZO_BadHuffmanCode;
}

/* ------------------------------------------------------------------ */
/* header segments                                                     */
/* ------------------------------------------------------------------ */

/* decoder.zig:490-618 */
static int process_sof(decoder *d, int32_t n) {
    if (d->num_components != 0) return ZO_MultipleSofMarkers;
    switch (n) {
    case 6 + 3 * 1: d->num_components = 1; break;
    case 6 + 3 * 3: d->num_components = 3; break;
    case 6 + 3 * 4: d->num_components = 4; break;
    default: return ZO_NumberComponents;
    }
    TRY(read_full(d, d->tmp, (size_t)n));
    if (d->tmp[0] != 8) return ZO_Precision;
    d->height = ((uint32_t)d->tmp[1] << 8) + d->tmp[2];
    d->width = ((uint32_t)d->tmp[3] << 8) + d->tmp[4];
    if (d->tmp[5] != d->num_components) return ZO_SofWrongLength;

    for (int i = 0; i < d->num_components; i++) {
        d->comp[i].id = d->tmp[6 + 3 * i];
        for (int j = 0; j < i; j++)
            if (d->comp[i].id == d->comp[j].id) return ZO_RepeatedComponentIdentifier;
        d->comp[i].tq = d->tmp[8 + 3 * i];
        if (d->comp[i].tq > MAX_TQ) return ZO_BadTqValue;
        uint8_t hv = d->tmp[7 + 3 * i];
        int32_t h = hv >> 4, v = hv & 0x0f;
        if (h < 1 || 4 < h || v < 1 || 4 < v) return ZO_LumaChromaSubSamplingRatio;
        if (h == 3 || v == 3) return ZO_LumaChromaSubSamplingRatio;
        switch (d->num_components) {
        case 1:
            h = 1;
            v = 1;
            break;
        case 3:
            switch (i) {
            case 0:
                if (v == 4) return ZO_LumaChromaSubSamplingRatio;
                break;
            case 1:
                if (d->comp[0].h % h != 0 || d->comp[0].v % v != 0) return ZO_LumaChromaSubSamplingRatio;
                break;
            case 2:
                if (d->comp[1].h != h || d->comp[1].v != v) return ZO_LumaChromaSubSamplingRatio;
                break;
            }
            break;
        case 4:
            switch (i) {
            case 0:
                if (hv != 0x11 && hv != 0x22) return ZO_LumaChromaSubSamplingRatio;
                break;
            case 1:
            case 2:
                if (hv != 0x11) return ZO_LumaChromaSubSamplingRatio;
                break;
            case 3:
                if (d->comp[0].h != h || d->comp[0].v != v) return ZO_LumaChromaSubSamplingRatio;
                break;
            }
            break;
        }
        d->comp[i].h = h;
        d->comp[i].v = v;
    }
    return ZO_OK;
}

/* decoder.zig:621-627 */
static int process_dri(decoder *d, int32_t n) {
    if (n != 2) return ZO_DriWrongLength;
    TRY(read_full(d, d->tmp, 2));
    d->restart_interval = (uint16_t)(((uint16_t)d->tmp[0] << 8) + d->tmp[1]);
    return ZO_OK;
}

/* decoder.zig:629-666 */
static int process_dqt(decoder *d, int32_t n) {
    int32_t local_n = n;
    while (local_n > 0) {
        local_n -= 1;
        uint8_t qi;
        TRY(read_byte(d, &qi));
        uint8_t tq = qi & 0x0f;
        if (tq > MAX_TQ) return ZO_BadTqValue;
        int stop = 0;
        switch (qi >> 4) {
        case 0:
            if (local_n < BLOCK_SIZE) { stop = 1; break; }
            local_n -= BLOCK_SIZE;
            TRY(read_full(d, d->tmp, BLOCK_SIZE));
            for (int i = 0; i < BLOCK_SIZE; i++) d->quant[tq][i] = d->tmp[i];
            break;
        case 1:
            if (local_n < 2 * BLOCK_SIZE) { stop = 1; break; }
            local_n -= 2 * BLOCK_SIZE;
            TRY(read_full(d, d->tmp, 2 * BLOCK_SIZE));
            for (int i = 0; i < BLOCK_SIZE; i++)
                d->quant[tq][i] = ((int32_t)d->tmp[2 * i] << 8) | d->tmp[2 * i + 1];
            break;
        default: return ZO_BadPqValue;
        }
        if (stop) break;
    }
    if (local_n != 0) return ZO_DqtWrongLength;
    return ZO_OK;
}

/* decoder.zig:668-680 */
static int process_app0(decoder *d, int32_t n) {
    if (n < 5) return ignore(d, n);
    TRY(read_full(d, d->tmp, 5));
    int32_t local_n = n - 5;
    d->jfif = d->tmp[0] == 'J' && d->tmp[1] == 'F' && d->tmp[2] == 'I' && d->tmp[3] == 'F' && d->tmp[4] == 0;
    if (n > 0) return ignore(d, local_n);
    return ZO_OK;
}

/* decoder.zig:682-697 */
static int process_app14(decoder *d, int32_t n) {
    if (n < 12) return ignore(d, n);
    TRY(read_full(d, d->tmp, 12));
    int32_t local_n = n - 12;
    if (d->tmp[0] == 'A' && d->tmp[1] == 'd' && d->tmp[2] == 'o' && d->tmp[3] == 'b' && d->tmp[4] == 'e') {
        d->adobe_transform_valid = 1;
        /* @enumFromInt on a 3-value enum: values > 2 are illegal behaviour in Zig;
         * kept as "not unknown" here. */
        d->adobe_transform = d->tmp[11];
    }
    if (n > 0) return ignore(d, local_n);
    return ZO_OK;
}

/* decoder.zig:1026-1111 */
static int process_dht(decoder *d, int32_t n) {
    int32_t local_n = n;
    while (local_n > 0) {
        if (local_n < MAX_CODE_LENGTH + 1) return ZO_DhtWrongLength;
        TRY(read_full(d, d->tmp, MAX_CODE_LENGTH + 1));
        uint8_t tc = d->tmp[0] >> 4;
        if (tc > MAX_TC) return ZO_BadTcValue;
        uint8_t th = d->tmp[0] & 0x0f;
        if (th > MAX_TH || (d->baseline && th > 1)) return ZO_BadThValue;
        huff_table *h = &d->huff[tc][th];

        h->num_codes = 0;
        int32_t num_codes[MAX_CODE_LENGTH];
        for (int i = 0; i < MAX_CODE_LENGTH; i++) {
            num_codes[i] = d->tmp[i + 1];
            h->num_codes += num_codes[i];
        }
        if (h->num_codes == 0) return ZO_HuffZeroLength;
        if (h->num_codes > MAX_NUM_CODES) return ZO_HuffTooLong;
        local_n -= h->num_codes + MAX_CODE_LENGTH + 1;
        if (local_n < 0) return ZO_DhtWrongLength;
        TRY(read_full(d, h->vals, (size_t)h->num_codes));

        memset(h->lut, 0, sizeof(h->lut));
        uint32_t code = 0;
        size_t val_index = 0;
        for (int i = 0; i < LUT_SIZE; i++) {
            code <<= 1;
            for (int32_t j = 0; j < num_codes[i]; j++) {
                uint32_t base = code << (7 - i);
                uint16_t lut_value = (uint16_t)(((uint16_t)h->vals[val_index] << 8) | (uint16_t)(2 + i));
                for (uint32_t k = 0; k < (1u << (7 - i)); k++) {
                    if ((base | k) >= (1u << LUT_SIZE)) return ZO_ReferencePanics; /* index out of bounds in Zig */
                    h->lut[base | k] = lut_value;
                }
                code += 1;
                val_index += 1;
            }
        }

        int32_t code_base = 0, index = 0;
        for (int i = 0; i < MAX_CODE_LENGTH; i++) {
            if (num_codes[i] == 0) {
                h->min_codes[i] = -1;
                h->max_codes[i] = -1;
                h->vals_indices[i] = -1;
            } else {
                h->min_codes[i] = code_base;
                h->max_codes[i] = code_base + num_codes[i] - 1;
                h->vals_indices[i] = index;
                code_base += num_codes[i];
                index += num_codes[i];
            }
            code_base <<= 1;
        }
    }
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* IDCT (idct.zig:77-201)                                              */
/* ------------------------------------------------------------------ */
#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565
#define W1PW7 (W1 + W7)
#define W1MW7 (W1 - W7)
#define W2PW6 (W2 + W6)
#define W2MW6 (W2 - W6)
#define W3PW5 (W3 + W5)
#define W3MW5 (W3 - W5)
#define R2 181

void zo_idct(int32_t src[64]) {
    for (int y = 0; y < 8; y++) {
        int32_t *s = src + y * 8;
        if (s[1] == 0 && s[2] == 0 && s[3] == 0 && s[4] == 0 && s[5] == 0 && s[6] == 0 && s[7] == 0) {
            int32_t dc = wshl(s[0], 3);
            for (int k = 0; k < 8; k++) s[k] = dc;
            continue;
        }
        int32_t x0 = wadd(wshl(s[0], 11), 128);
        int32_t x1 = wshl(s[4], 11);
        int32_t x2 = s[6], x3 = s[2], x4 = s[1], x5 = s[7], x6 = s[5], x7 = s[3];

        int32_t x8 = wmul(W7, wadd(x4, x5));
        x4 = wadd(x8, wmul(W1MW7, x4));
        x5 = wsub(x8, wmul(W1PW7, x5));
        x8 = wmul(W3, wadd(x6, x7));
        x6 = wsub(x8, wmul(W3MW5, x6));
        x7 = wsub(x8, wmul(W3PW5, x7));

        x8 = wadd(x0, x1);
        x0 = wsub(x0, x1);
        x1 = wmul(W6, wadd(x3, x2));
        x2 = wsub(x1, wmul(W2PW6, x2));
        x3 = wadd(x1, wmul(W2MW6, x3));
        x1 = wadd(x4, x6);
        x4 = wsub(x4, x6);
        x6 = wadd(x5, x7);
        x5 = wsub(x5, x7);

        x7 = wadd(x8, x3);
        x8 = wsub(x8, x3);
        x3 = wadd(x0, x2);
        x0 = wsub(x0, x2);
        x2 = asr(wadd(wmul(R2, wadd(x4, x5)), 128), 8);
        x4 = asr(wadd(wmul(R2, wsub(x4, x5)), 128), 8);

        s[0] = asr(wadd(x7, x1), 8);
        s[1] = asr(wadd(x3, x2), 8);
        s[2] = asr(wadd(x0, x4), 8);
        s[3] = asr(wadd(x8, x6), 8);
        s[4] = asr(wsub(x8, x6), 8);
        s[5] = asr(wsub(x0, x4), 8);
        s[6] = asr(wsub(x3, x2), 8);
        s[7] = asr(wsub(x7, x1), 8);
    }
    for (int x = 0; x < 8; x++) {
        int32_t *s = src + x;
        int32_t y0 = wadd(wshl(s[8 * 0], 8), 8192);
        int32_t y1 = wshl(s[8 * 4], 8);
        int32_t y2 = s[8 * 6], y3 = s[8 * 2], y4 = s[8 * 1], y5 = s[8 * 7], y6 = s[8 * 5], y7 = s[8 * 3];

        int32_t y8 = wadd(wmul(W7, wadd(y4, y5)), 4);
        y4 = asr(wadd(y8, wmul(W1MW7, y4)), 3);
        y5 = asr(wsub(y8, wmul(W1PW7, y5)), 3);
        y8 = wadd(wmul(W3, wadd(y6, y7)), 4);
        y6 = asr(wsub(y8, wmul(W3MW5, y6)), 3);
        y7 = asr(wsub(y8, wmul(W3PW5, y7)), 3);

        y8 = wadd(y0, y1);
        y0 = wsub(y0, y1);
        y1 = wadd(wmul(W6, wadd(y3, y2)), 4);
        y2 = asr(wsub(y1, wmul(W2PW6, y2)), 3);
        y3 = asr(wadd(y1, wmul(W2MW6, y3)), 3);
        y1 = wadd(y4, y6);
        y4 = wsub(y4, y6);
        y6 = wadd(y5, y7);
        y5 = wsub(y5, y7);

        y7 = wadd(y8, y3);
        y8 = wsub(y8, y3);
        y3 = wadd(y0, y2);
        y0 = wsub(y0, y2);
        y2 = asr(wadd(wmul(R2, wadd(y4, y5)), 128), 8);
        y4 = asr(wadd(wmul(R2, wsub(y4, y5)), 128), 8);

        s[8 * 0] = asr(wadd(y7, y1), 14);
        s[8 * 1] = asr(wadd(y3, y2), 14);
        s[8 * 2] = asr(wadd(y0, y4), 14);
        s[8 * 3] = asr(wadd(y8, y6), 14);
        s[8 * 4] = asr(wsub(y8, y6), 14);
        s[8 * 5] = asr(wsub(y0, y4), 14);
        s[8 * 6] = asr(wsub(y3, y2), 14);
        s[8 * 7] = asr(wsub(y7, y1), 14);
    }
}

/* ------------------------------------------------------------------ */
/* image allocation (decoder.zig:1708-1783, image.zig:484-583,638-668) */
/* ------------------------------------------------------------------ */
static int make_img(decoder *d, int32_t mxx, int32_t myy) {
    if (d->num_components == 1) {
        /* GrayImage.init (no zero fill) + subImage: same buffer, rect clipped.
         * A fresh buffer on every SOS; the previous one is only
         * released for progressive frames in the reference -- here always, the
         * leak is not part of the observable result.  calloc: contents of the
         * never-written MCU padding are undefined in the reference. */
        free(d->gray_pixels);
        size_t len = (size_t)(8 * mxx) * (size_t)(8 * myy);
        d->gray_pixels = (uint8_t *)calloc(len ? len : 1, 1);
        if (!d->gray_pixels) return ZO_OutOfMemory;
        d->gray_len = len;
        d->gray_stride = (size_t)(8 * mxx);
        if (d->width == 0 || d->height == 0) return ZO_CreateImageFailed; /* Intersect -> null */
        d->have_gray = 1;
        return ZO_OK;
    }
    int32_t h0 = d->comp[0].h, v0 = d->comp[0].v;
    /* @divExact: validated by processSof for 3 components; 4 components use 0x11/0x22 only */
    int32_t h_ratio = h0 / d->comp[1].h, v_ratio = v0 / d->comp[1].v;
    int ratio;
    switch (h_ratio << 4 | v_ratio) {
    case 0x11: ratio = ZO_R444; break;
    case 0x12: ratio = ZO_R440; break;
    case 0x21: ratio = ZO_R422; break;
    case 0x22: ratio = ZO_R420; break;
    case 0x41: ratio = ZO_R411; break;
    case 0x42: ratio = ZO_R410; break;
    default: return ZO_ReferencePanics; /* unreachable in the reference */
    }
    int32_t w = 8 * h0 * mxx, h = 8 * v0 * myy, cw, ch;
    switch (ratio) { /* yCbCrSize, image.zig:521-555, rect.min = 0 */
    case ZO_R422: cw = (w + 1) / 2; ch = h; break;
    case ZO_R420: cw = (w + 1) / 2; ch = (h + 1) / 2; break;
    case ZO_R440: cw = w; ch = (h + 1) / 2; break;
    case ZO_R411: cw = (w + 3) / 4; ch = h; break;
    case ZO_R410: cw = (w + 3) / 4; ch = (h + 1) / 2; break;
    default: cw = w; ch = h; break;
    }
    size_t i0 = (size_t)w * h, i1 = i0 + (size_t)cw * ch, i2 = i0 + 2 * (size_t)cw * ch;
    /* YCbCrImage.init zero-fills; subImage copies the whole buffer and keeps strides */
    d->ycc_pixels = (uint8_t *)calloc(i2 ? i2 : 1, 1);
    if (!d->ycc_pixels) return ZO_OutOfMemory;
    d->ycc_len = i2;
    d->ycc_y = d->ycc_pixels;
    d->ycc_cb = d->ycc_pixels + i0;
    d->ycc_cr = d->ycc_pixels + i1;
    d->y_stride = (size_t)w;
    d->c_stride = (size_t)cw;
    d->subsample_ratio = ratio;
    if (d->width == 0 || d->height == 0) return ZO_CreateImageFailed;
    d->have_ycbcr = 1;
    if (d->num_components == 4) {
        int32_t h3 = d->comp[3].h, v3 = d->comp[3].v;
        d->black_len = (size_t)(8 * h3 * mxx) * (size_t)(8 * v3 * myy);
        d->black_pixels = (uint8_t *)calloc(d->black_len ? d->black_len : 1, 1);
        if (!d->black_pixels) return ZO_OutOfMemory;
        d->black_stride = (size_t)(8 * h3 * mxx);
    }
    return ZO_OK;
}

/* decoder.zig:1553-1634 */
/* the arithmetic of reconstructBlock (decoder.zig:1564-1570, 1611-1633): dequantise (tables in zig-zag order),
 * idct.transform, level shift, clamp, store 8x8 */
static void dequant_idct_store(int32_t *b, const int32_t *qt, uint8_t *dst, size_t stride) {
    for (int zig = 0; zig < BLOCK_SIZE; zig++) b[unzig[zig]] = wmul(b[unzig[zig]], qt[zig]);
    zo_idct(b);
    for (int y = 0; y < 8; y++) {
        for (int x = 0; x < 8; x++) {
            int32_t c = b[y * 8 + x];
            if (c < -128) c = 0;
            else if (c > 127) c = 255;
            else c += 128;
            dst[(size_t)y * stride + x] = (uint8_t)c;
        }
    }
}

static int reconstruct_block(decoder *d, int32_t *b, int32_t bx, int32_t by, int ci) {
    const int32_t *qt = d->quant[d->comp[ci].tq];
    uint8_t *dst;
    size_t stride;
    if (d->num_components == 1) {
        if (!d->have_gray) return ZO_ReferencePanics;
        dst = d->gray_pixels + 8 * ((size_t)by * d->gray_stride + (size_t)bx);
        stride = d->gray_stride;
    } else {
        if (!d->have_ycbcr) return ZO_ReferencePanics;
        switch (ci) {
        case 0: dst = d->ycc_y + 8 * ((size_t)by * d->y_stride + (size_t)bx); stride = d->y_stride; break;
        case 1: dst = d->ycc_cb + 8 * ((size_t)by * d->c_stride + (size_t)bx); stride = d->c_stride; break;
        case 2: dst = d->ycc_cr + 8 * ((size_t)by * d->c_stride + (size_t)bx); stride = d->c_stride; break;
        case 3: dst = d->black_pixels + 8 * ((size_t)by * d->black_stride + (size_t)bx); stride = d->black_stride; break;
        default: return ZO_UnsupportedComponent;
        }
    }
    dequant_idct_store(b, qt, dst, stride);
    return ZO_OK;
}

/* test helper: reconstructBlock on n free-standing blocks (coef: int32[n][64], natural order as decoded; quant_zz:
 * 64 values in zig-zag order as a DQT segment holds them; out: u8[n][64], row-major 8x8 each) */
void zo_reconstruct_blocks(const int32_t *coef, const int32_t *quant_zz, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        int32_t b[BLOCK_SIZE];
        memcpy(b, coef + i * BLOCK_SIZE, sizeof(b));
        dequant_idct_store(b, quant_zz, out + i * BLOCK_SIZE, 8);
    }
}

static void tap_block(decoder *d, int ci, int32_t bx, int32_t by, const int32_t *b) {
    if (!d->tap) return;
    if (d->tap->recs && d->tap->count < d->tap->cap) {
        zo_block_rec *r = &d->tap->recs[d->tap->count];
        r->comp = ci;
        r->bx = bx;
        r->by = by;
        memcpy(r->coef, b, sizeof(r->coef));
    }
    d->tap->count++;
}

/* decoder.zig:1636-1661 */
static int reconstruct_progressive_image(decoder *d) {
    int32_t h0 = d->comp[0].h;
    int32_t mxx = ((int32_t)d->width + 8 * h0 - 1) / (8 * h0);
    if (d->tap) {
        for (int i = 0; i < d->num_components; i++) {
            if (!d->prog_coef[i]) continue;
            int32_t stride = mxx * d->comp[i].h;
            for (size_t k = 0; k < d->prog_len[i]; k++)
                tap_block(d, i, (int32_t)(k % (size_t)stride), (int32_t)(k / (size_t)stride), d->prog_coef[i][k]);
        }
    }
    for (int i = 0; i < d->num_components; i++) {
        if (!d->prog_coef[i]) continue;
        size_t v = (size_t)(8 * (d->comp[0].v / d->comp[i].v));
        size_t h = (size_t)(8 * (d->comp[0].h / d->comp[i].h));
        size_t stride = (size_t)(mxx * d->comp[i].h);
        for (size_t by = 0; by * v < d->height; by++)
            for (size_t bx = 0; bx * h < d->width; bx++)
                TRY(reconstruct_block(d, d->prog_coef[i][by * stride + bx], (int32_t)bx, (int32_t)by, i));
    }
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* progressive refinement (decoder.zig:1459-1549)                      */
/* ------------------------------------------------------------------ */
static int refine_non_zeroes(decoder *d, int32_t *b, int32_t zig, int32_t zig_end, int32_t nz, int32_t delta,
                             int32_t *out_zig) {
    for (; zig <= zig_end; zig++) {
        int index = unzig[zig];
        if (b[index] == 0) {
            if (nz == 0) break;
            nz -= 1;
            continue;
        }
        int bit;
        TRY(decode_bit(d, &bit));
        if (!bit) continue;
        if (b[index] >= 0) b[index] = wadd(b[index], delta);
        else b[index] = wsub(b[index], delta);
        if (b[index] < -32768 || b[index] > 32767) d->coef_overflow = 1;
    }
    *out_zig = zig;
    return ZO_OK;
}

static int refine(decoder *d, int32_t *b, huff_table *h, int32_t zig_start, int32_t zig_end, int32_t delta) {
    if (zig_start == 0) {
        if (zig_end != 0) return ZO_ReferencePanics;
        int bit;
        TRY(decode_bit(d, &bit));
        if (bit) b[0] |= delta;
        return ZO_OK;
    }
    int32_t zig = zig_start;
    if (d->eob_run == 0) {
        for (; zig <= zig_end; zig++) {
            int32_t z = 0;
            uint8_t value;
            TRY(decode_huffman(d, h, &value));
            uint8_t val0 = value >> 4, val1 = value & 0x0f;
            int brk = 0;
            switch (val1) {
            case 0:
                if (val0 != 0x0f) {
                    d->eob_run = (uint16_t)(1u << val0);
                    if (val0 != 0) {
                        uint32_t bits;
                        TRY(decode_bits(d, val0, &bits));
                        d->eob_run |= (uint16_t)bits;
                    }
                    brk = 1;
                }
                break;
            case 1: {
                z = delta;
                int bit;
                TRY(decode_bit(d, &bit));
                if (!bit) z = -z;
                break;
            }
            default: return ZO_UnexpectedHuffmanCode;
            }
            if (brk) break;
            TRY(refine_non_zeroes(d, b, zig, zig_end, (int32_t)val0, delta, &zig));
            if (zig > zig_end) return ZO_TooManyCoefficients;
            if (z != 0) b[unzig[zig]] = z;
        }
    }
    if (d->eob_run > 0) {
        d->eob_run -= 1;
        int32_t dummy;
        TRY(refine_non_zeroes(d, b, zig, zig_end, -1, delta, &dummy));
    }
    return ZO_OK;
}

/* decoder.zig:1671-1705 */
static int find_rst(decoder *d, uint8_t expected_rst) {
    for (;;) {
        size_t i = 0;
        if (d->tmp[0] == 0xff) {
            if (d->tmp[1] == expected_rst) return ZO_OK;
            else if (d->tmp[1] == 0xff) i = 1;
            else if (d->tmp[1] != 0x00) return ZO_BadRSTMarker;
        } else if (d->tmp[1] == 0xff) {
            d->tmp[0] = 0xff;
            i = 1;
        }
        TRY(read_full(d, d->tmp + i, 2 - i));
    }
}

/* ------------------------------------------------------------------ */
/* scan (decoder.zig:1148-1455)                                        */
/* ------------------------------------------------------------------ */
static int process_sos(decoder *d, int32_t n) {
    if (d->num_components == 0) return ZO_MissingSosMarker;
    if (n < 6 || 4 + 2 * (int32_t)d->num_components < n || n % 2 != 0) return ZO_SosWrongLength;
    TRY(read_full(d, d->tmp, (size_t)n));
    int32_t n_comp = d->tmp[0];
    if (n != 4 + 2 * n_comp) return ZO_SosWrongLength;

    struct { uint8_t id, td, ta; } scan[MAX_COMPONENTS];
    memset(scan, 0, sizeof(scan));
    int32_t total_hv = 0;
    for (int i = 0; i < n_comp; i++) {
        uint8_t cs = d->tmp[1 + 2 * i];
        int ci = -1;
        for (int j = 0; j < d->num_components; j++)
            if (cs == d->comp[j].id) { ci = j; break; }
        if (ci < 0) return ZO_UnknownComponentSelector;
        scan[i].id = (uint8_t)ci;
        for (int j = 0; j < i; j++)
            if (scan[i].id == scan[j].id) return ZO_RepeatedComponentIdentifier;
        total_hv += d->comp[ci].h * d->comp[ci].v;
        scan[i].td = d->tmp[2 + 2 * i] >> 4;
        if (scan[i].td > MAX_TH || (d->baseline && scan[i].td > 1)) return ZO_BadTdValue;
        scan[i].ta = d->tmp[2 + 2 * i] & 0x0f;
        if (scan[i].ta > MAX_TH || (d->baseline && scan[i].ta > 1)) return ZO_BadTaValue;
    }
    if (d->num_components > 1 && total_hv > 10) return ZO_SamplingFactorsTooLarge;

    int32_t zig_start = 0, zig_end = BLOCK_SIZE - 1;
    uint32_t ah = 0, al = 0;
    if (d->progressive) {
        zig_start = d->tmp[1 + 2 * n_comp];
        zig_end = d->tmp[2 + 2 * n_comp];
        ah = d->tmp[3 + 2 * n_comp] >> 4;
        al = d->tmp[3 + 2 * n_comp] & 0x0f;
        if ((zig_start == 0 && zig_end != 0) || zig_start > zig_end || BLOCK_SIZE <= zig_end)
            return ZO_BadSpectralSelection;
        if (zig_start != 0 && n_comp != 1) return ZO_ProgressiveACCoefficientsForMoreThanOneComponent;
        if (ah != 0 && ah != al + 1) return ZO_BadSuccessiveApproximation;
    }

    int32_t h0 = d->comp[0].h, v0 = d->comp[0].v;
    int32_t w = (int32_t)d->width, hgt = (int32_t)d->height;
    int32_t mxx = (w + 8 * h0 - 1) / (8 * h0);
    int32_t myy = (hgt + 8 * v0 - 1) / (8 * v0);
    if (!d->have_ycbcr) TRY(make_img(d, mxx, myy));

    if (d->progressive) {
        /* loops num_components over scan[i] (unset entries are component 0) */
        for (int i = 0; i < d->num_components; i++) {
            int ci = scan[i].id;
            if (!d->prog_coef[ci]) {
                size_t cnt = (size_t)(mxx * myy * d->comp[ci].h * d->comp[ci].v);
                d->prog_coef[ci] = (int32_t(*)[BLOCK_SIZE])calloc(cnt ? cnt : 1, sizeof(int32_t[BLOCK_SIZE]));
                if (!d->prog_coef[ci]) return ZO_OutOfMemory;
                d->prog_len[ci] = cnt;
            }
        }
    }

    d->bits.a = 0;
    d->bits.m = 0;
    d->bits.n = 0;
    int32_t mcu = 0;
    uint8_t expected_rst = 0xd0;
    int32_t bx = 0, by = 0, block_count = 0;
    int32_t dc[MAX_COMPONENTS] = {0, 0, 0, 0};
    int32_t b[BLOCK_SIZE];

    /* eob_run is a decoder field that only RSTn resets (decoder.zig:144, :1451): a run left open by the
     * previous scan is still counted down here.  Recorded for the tests (the GPU path refuses such files). */
    if (d->eob_run != 0) d->eob_carry = 1;
    for (int32_t my = 0; my < myy; my++) {
        for (int32_t mx = 0; mx < mxx; mx++) {
            for (int k = 0; k < n_comp; k++) {
                int ci = scan[k].id;
                int32_t hi = d->comp[ci].h, vi = d->comp[ci].v;
                for (int32_t j = 0; j < hi * vi; j++) {
                    if (n_comp != 1) {
                        bx = hi * mx + j % hi;
                        by = vi * my + j / hi;
                    } else {
                        bx = block_count % (mxx * hi);
                        by = block_count / (mxx * hi);
                        block_count += 1;
                        if ((uint32_t)(bx * 8) >= d->width || (uint32_t)(by * 8) >= d->height) continue;
                    }

                    if (d->progressive) {
                        size_t bi = (size_t)(by * mxx * hi + bx);
                        memcpy(b, d->prog_coef[ci][bi], sizeof(b));
                    } else {
                        memset(b, 0, sizeof(b));
                    }

                    if (ah != 0) {
                        TRY(refine(d, b, &d->huff[1][scan[k].ta], zig_start, zig_end, wshl(1, (int)al)));
                    } else {
                        int32_t zig = zig_start;
                        if (zig == 0) {
                            zig += 1;
                            uint8_t value;
                            TRY(decode_huffman(d, &d->huff[0][scan[k].td], &value));
                            if (value > 16) return ZO_ExcessiveDCComponent;
                            int32_t dc_delta;
                            TRY(receive_extend(d, value, &dc_delta));
                            dc[ci] = wadd(dc[ci], dc_delta);
                            b[0] = wshl(dc[ci], (int)al);
                            if (b[0] < -32768 || b[0] > 32767) d->coef_overflow = 1;
                        }
                        if (zig <= zig_end && d->eob_run > 0) {
                            d->eob_run -= 1;
                        } else {
                            huff_table *huff = &d->huff[1][scan[k].ta];
                            for (; zig <= zig_end; zig++) {
                                uint8_t value;
                                TRY(decode_huffman(d, huff, &value));
                                uint8_t val0 = value >> 4, val1 = value & 0x0f;
                                if (val1 != 0) {
                                    zig += val0;
                                    if (zig > zig_end) break;
                                    int32_t ac;
                                    TRY(receive_extend(d, val1, &ac));
                                    b[unzig[zig]] = wshl(ac, (int)al);
                                    if (b[unzig[zig]] < -32768 || b[unzig[zig]] > 32767) d->coef_overflow = 1;
                                } else {
                                    if (val0 != 0x0f) {
                                        d->eob_run = (uint16_t)(1u << val0);
                                        if (val0 != 0) {
                                            uint32_t bits;
                                            TRY(decode_bits(d, val0, &bits));
                                            d->eob_run |= (uint16_t)bits;
                                        }
                                        d->eob_run -= 1;
                                        break;
                                    }
                                    zig += 0x0f;
                                }
                            }
                        }
                    }

                    if (d->progressive) {
                        size_t bi = (size_t)(by * mxx * hi + bx);
                        memcpy(d->prog_coef[ci][bi], b, sizeof(b));
                        continue;
                    }
                    tap_block(d, ci, bx, by, b);
                    TRY(reconstruct_block(d, b, bx, by, ci));
                }
            }

            mcu += 1;
            if (d->restart_interval > 0 && mcu % d->restart_interval == 0 && mcu < mxx * myy) {
                TRY(read_full(d, d->tmp, 2));
                if (d->tmp[0] != 0xff || d->tmp[1] != expected_rst) TRY(find_rst(d, expected_rst));
                expected_rst += 1;
                if (expected_rst == 0xd7 + 1) expected_rst = 0xd0;
                d->bits.a = 0;
                d->bits.m = 0;
                d->bits.n = 0;
                memset(dc, 0, sizeof(dc));
                d->eob_run = 0;
            }
        }
    }
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* colour exits                                                        */
/* ------------------------------------------------------------------ */

/* image.zig:594-605 with rect.min = 0 */
static inline size_t c_offset(const zo_image *m, int32_t x, int32_t y) {
    size_t s = m->c_stride;
    switch (m->subsample_ratio) {
    case ZO_R422: return (size_t)y * s + (size_t)(x / 2);
    case ZO_R420: return (size_t)(y / 2) * s + (size_t)(x / 2);
    case ZO_R440: return (size_t)(y / 2) * s + (size_t)x;
    case ZO_R411: return (size_t)y * s + (size_t)(x / 4);
    case ZO_R410: return (size_t)(y / 2) * s + (size_t)(x / 4);
    default: return (size_t)y * s + (size_t)x;
    }
}

/* decoder.zig:699-709 */
static int is_rgb(const decoder *d) {
    if (d->jfif) return 0;
    if (d->adobe_transform_valid && d->adobe_transform == 0) return 1;
    return d->comp[0].id == 'R' && d->comp[1].id == 'G' && d->comp[2].id == 'B';
}

static void ycc_view(const decoder *d, zo_image *m) {
    memset(m, 0, sizeof(*m));
    m->variant = ZO_YCBCR;
    m->width = (int32_t)d->width;
    m->height = (int32_t)d->height;
    m->pixels = d->ycc_pixels;
    m->pixels_len = d->ycc_len;
    m->y = d->ycc_y;
    m->cb = d->ycc_cb;
    m->cr = d->ycc_cr;
    m->y_stride = d->y_stride;
    m->c_stride = d->c_stride;
    m->subsample_ratio = d->subsample_ratio;
}

/* decoder.zig:751-783 */
static int convert_to_rgb(decoder *d, zo_image *out) {
    zo_image src;
    ycc_view(d, &src);
    size_t c_scale = (size_t)(d->comp[0].h / d->comp[1].h);
    size_t len = (size_t)4 * d->width * d->height;
    uint8_t *pix = (uint8_t *)malloc(len ? len : 1);
    if (!pix) return ZO_OutOfMemory;
    size_t stride = (size_t)4 * d->width;
    for (int32_t y = 0; y < (int32_t)d->height; y++) {
        size_t po = (size_t)y * stride;
        size_t yo = (size_t)y * src.y_stride;
        size_t co = c_offset(&src, 0, y);
        for (size_t i = 0; i < d->width; i++) {
            pix[po + 4 * i + 0] = src.y[yo + i];
            pix[po + 4 * i + 1] = src.cb[co + i / c_scale];
            pix[po + 4 * i + 2] = src.cr[co + i / c_scale];
            pix[po + 4 * i + 3] = 255;
        }
    }
    memset(out, 0, sizeof(*out));
    out->variant = ZO_RGBA;
    out->width = (int32_t)d->width;
    out->height = (int32_t)d->height;
    out->pixels = out->pix = pix;
    out->pixels_len = len;
    out->stride = stride;
    return ZO_OK;
}

static inline void ycbcr_to_rgb8(uint8_t yv, uint8_t cbv, uint8_t crv, uint8_t *r8, uint8_t *g8, uint8_t *b8) {
    /* Go image/draw drawYCbCr == the intent of util.zig:40-83 */
    int32_t yy1 = (int32_t)yv * 0x10101, cb1 = (int32_t)cbv - 128, cr1 = (int32_t)crv - 128;
    int32_t r = yy1 + 91881 * cr1;
    if (((uint32_t)r & 0xff000000u) == 0) r >>= 16; else r = ~(r >> 31);
    int32_t g = yy1 - 22554 * cb1 - 46802 * cr1;
    if (((uint32_t)g & 0xff000000u) == 0) g >>= 16; else g = ~(g >> 31);
    int32_t b = yy1 + 116130 * cb1;
    if (((uint32_t)b & 0xff000000u) == 0) b >>= 16; else b = ~(b >> 31);
    *r8 = (uint8_t)r;
    *g8 = (uint8_t)g;
    *b8 = (uint8_t)b;
}

/* decoder.zig:792-902 */
static int apply_black(decoder *d, zo_image *out) {
    if (!d->adobe_transform_valid) return ZO_UnsupportedColorModel;
    zo_image src;
    ycc_view(d, &src);
    size_t len = (size_t)4 * d->width * d->height;
    size_t stride = (size_t)4 * d->width;
    uint8_t *pix = (uint8_t *)malloc(len ? len : 1);
    if (!pix) return ZO_OutOfMemory;
    memset(out, 0, sizeof(*out));
    out->variant = ZO_CMYK;
    out->width = (int32_t)d->width;
    out->height = (int32_t)d->height;
    out->pixels = out->pix = pix;
    out->pixels_len = len;
    out->stride = stride;

    if (d->adobe_transform != 0) {
        /* YCbCrK.  The reference's drawYCbCr (util.zig:10-291) is broken
         * (off-by-one loops, out-of-bounds on the last row); what
         * follows is its evident intent, i.e. Go image/jpeg's applyBlack:
         * RGB from the YCbCr planes, K = 255 - black.  Parity unpinned. */
        out->ycck_intent = 1;
        for (int32_t y = 0; y < (int32_t)d->height; y++) {
            for (int32_t x = 0; x < (int32_t)d->width; x++) {
                size_t yi = (size_t)y * src.y_stride + (size_t)x;
                size_t ci = c_offset(&src, x, y);
                uint8_t *p = pix + (size_t)y * stride + (size_t)4 * x;
                ycbcr_to_rgb8(src.y[yi], src.cb[ci], src.cr[ci], &p[0], &p[1], &p[2]);
                p[3] = (uint8_t)(255 - d->black_pixels[(size_t)y * d->black_stride + (size_t)x]);
            }
        }
        return ZO_OK;
    }

    const uint8_t *tsrc[4] = {src.y, src.cb, src.cr, d->black_pixels};
    size_t tstride[4] = {src.y_stride, src.c_stride, src.c_stride, d->black_stride};
    for (int t = 0; t < 4; t++) {
        int subsample = d->comp[t].h != d->comp[0].h || d->comp[t].v != d->comp[0].v;
        for (int32_t y = 0; y < (int32_t)d->height; y++) {
            size_t sy = (size_t)y;
            if (subsample) sy >>= 1;
            for (int32_t x = 0; x < (int32_t)d->width; x++) {
                size_t sx = (size_t)x;
                if (subsample) sx >>= 1;
                pix[(size_t)y * stride + (size_t)4 * x + t] = (uint8_t)(255 - tsrc[t][sy * tstride[t] + sx]);
            }
        }
    }
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* marker loop (decoder.zig:220-373)                                   */
/* ------------------------------------------------------------------ */
static int decode_inner(decoder *d, int config_only, zo_image *out) {
    TRY(read_full(d, d->tmp, 2));
    if (d->tmp[0] != 0xff || d->tmp[1] != 0xd8) return ZO_InvalidSOIMarker;

    for (;;) {
        TRY(read_full(d, d->tmp, 2));
        while (d->tmp[0] != 0xff) {
            d->tmp[0] = d->tmp[1];
            TRY(read_byte(d, &d->tmp[1]));
        }
        uint8_t marker = d->tmp[1];
        if (marker == 0) continue;
        while (marker == 0xff) TRY(read_byte(d, &marker));
        if (marker == 0xd9) break;
        if (0xd0 <= marker && marker <= 0xd7) continue;

        TRY(read_full(d, d->tmp, 2));
        int32_t n = ((int32_t)d->tmp[0] << 8) + (int32_t)d->tmp[1] - 2;
        if (n < 0) return ZO_ShortSegmentLength;

        switch (marker) {
        case 0xc0:
        case 0xc1:
        case 0xc2:
            d->baseline = marker == 0xc0;
            d->progressive = marker == 0xc2;
            TRY(process_sof(d, n));
            if (config_only && d->jfif) return ZO_ConfigOnly;
            break;
        case 0xdb:
            if (config_only) TRY(ignore(d, n)); else TRY(process_dqt(d, n));
            break;
        case 0xdd:
            if (config_only) TRY(ignore(d, n)); else TRY(process_dri(d, n));
            break;
        case 0xc4:
            if (config_only) TRY(ignore(d, n)); else TRY(process_dht(d, n));
            break;
        case 0xda:
            if (config_only) return ZO_ConfigOnly;
            TRY(process_sos(d, n));
            break;
        case 0xe0: TRY(process_app0(d, n)); break;
        case 0xee: TRY(process_app14(d, n)); break;
        default:
            if ((0xe0 <= marker && marker <= 0xef) || marker == 0xfe) TRY(ignore(d, n));
            else if (marker < 0xc0) return ZO_UnknownMarker;
            else return ZO_UnsupportedMarker;
        }
    }

    if (d->progressive) TRY(reconstruct_progressive_image(d));

    if (d->have_gray) {
        memset(out, 0, sizeof(*out));
        out->variant = ZO_GRAY;
        out->width = (int32_t)d->width;
        out->height = (int32_t)d->height;
        out->pixels = out->pix = d->gray_pixels;
        out->pixels_len = d->gray_len;
        out->stride = d->gray_stride;
        d->gray_pixels = NULL; /* ownership moves to the image */
        return ZO_OK;
    } else if (d->have_ycbcr) {
        if (d->black_pixels) return apply_black(d, out);
        if (is_rgb(d)) return convert_to_rgb(d, out);
        ycc_view(d, out);
        d->ycc_pixels = NULL;
        return ZO_OK;
    }
    return ZO_MissingSosMarker;
}

static decoder *new_decoder(const uint8_t *data, size_t len) {
    decoder *d = (decoder *)calloc(1, sizeof(decoder));
    if (!d) return NULL;
    d->src = data;
    d->src_len = len;
    return d;
}

static void free_decoder(decoder *d) {
    free(d->gray_pixels);
    free(d->ycc_pixels);
    free(d->black_pixels);
    for (int i = 0; i < MAX_COMPONENTS; i++) free(d->prog_coef[i]);
    free(d);
}

/* test aid (not thread-safe): eob_carry of the latest zo_decode / zo_decode_tap, also when it failed */
static int g_last_eob_carry = 0, g_last_coef_overflow = 0;
int zo_last_eob_carry(void) { return g_last_eob_carry; }
int zo_last_coef_overflow(void) { return g_last_coef_overflow; }

int zo_decode_tap(const uint8_t *data, size_t len, zo_image *out, zo_tap *tap) {
    decoder *d = new_decoder(data, len);
    if (!d) return ZO_OutOfMemory;
    d->tap = tap;
    if (tap) tap->count = 0;
    memset(out, 0, sizeof(*out));
    int e = decode_inner(d, 0, out);
    if (e != ZO_OK) memset(out, 0, sizeof(*out));
    else out->eob_carry = d->eob_carry;
    g_last_eob_carry = d->eob_carry;
    g_last_coef_overflow = d->coef_overflow;
    free_decoder(d);
    return e;
}

int zo_decode(const uint8_t *data, size_t len, zo_image *out) { return zo_decode_tap(data, len, out, NULL); }

/* decoder.zig:178-218 */
int zo_decode_config(const uint8_t *data, size_t len, zo_config *out) {
    decoder *d = new_decoder(data, len);
    if (!d) return ZO_OutOfMemory;
    zo_image dummy;
    memset(&dummy, 0, sizeof(dummy));
    int e = decode_inner(d, 1, &dummy);
    if (e == ZO_OK) zo_free(&dummy);
    int ret = ZO_OK;
    if (e != ZO_OK && e != ZO_ConfigOnly) ret = e;
    else {
        out->width = d->width;
        out->height = d->height;
        switch (d->num_components) {
        case 1: out->color_model = ZO_GRAY; break;
        case 3:
        case 4: out->color_model = ZO_YCBCR; break;
        default: ret = ZO_InvalidSOIMarker; break;
        }
    }
    free_decoder(d);
    return ret;
}

void zo_free(zo_image *img) {
    if (!img) return;
    free(img->pixels);
    memset(img, 0, sizeof(*img));
}

/* ------------------------------------------------------------------ */
/* colour + rgbaPixels (color.zig:31-131, image.zig:103-130)           */
/* ------------------------------------------------------------------ */
void zo_ycbcr_to_rgba16(uint8_t yv, uint8_t cbv, uint8_t crv, uint32_t out[4]) {
    int32_t yy1 = (int32_t)yv * 0x10101, cb1 = (int32_t)cbv - 128, cr1 = (int32_t)crv - 128;
    int32_t r = yy1 + 91881 * cr1;
    r = (((uint32_t)r & 0xff000000u) == 0) ? (r >> 8) : (~(r >> 31) & 0xffff);
    int32_t g = yy1 - 22554 * cb1 - 46802 * cr1;
    g = (((uint32_t)g & 0xff000000u) == 0) ? (g >> 8) : (~(g >> 31) & 0xffff);
    int32_t b = yy1 + 116130 * cb1;
    b = (((uint32_t)b & 0xff000000u) == 0) ? (b >> 8) : (~(b >> 31) & 0xffff);
    out[0] = (uint32_t)r;
    out[1] = (uint32_t)g;
    out[2] = (uint32_t)b;
    out[3] = 0xffff;
}

void zo_cmyk_to_rgba16(uint8_t c, uint8_t m, uint8_t y, uint8_t k, uint32_t out[4]) {
    uint32_t w = 0xffffu - (uint32_t)k * 0x101u;
    out[0] = (0xffffu - (uint32_t)c * 0x101u) * w / 0xffffu;
    out[1] = (0xffffu - (uint32_t)m * 0x101u) * w / 0xffffu;
    out[2] = (0xffffu - (uint32_t)y * 0x101u) * w / 0xffffu;
    out[3] = 0xffff;
}

void zo_rgba_pixels(const zo_image *img, uint8_t *out) {
    int32_t W = img->width, H = img->height;
    for (int32_t y = 0; y < H; y++) {
        for (int32_t x = 0; x < W; x++) {
            uint32_t c[4];
            switch (img->variant) {
            case ZO_GRAY: {
                uint32_t v = img->pix[(size_t)y * img->stride + (size_t)x];
                v |= v << 8;
                c[0] = c[1] = c[2] = v;
                c[3] = 0xffff;
                break;
            }
            case ZO_YCBCR: {
                size_t yi = (size_t)y * img->y_stride + (size_t)x;
                size_t ci = c_offset(img, x, y);
                zo_ycbcr_to_rgba16(img->y[yi], img->cb[ci], img->cr[ci], c);
                break;
            }
            case ZO_RGBA: {
                const uint8_t *s = img->pix + (size_t)y * img->stride + (size_t)4 * x;
                for (int k = 0; k < 4; k++) c[k] = (uint32_t)s[k] | ((uint32_t)s[k] << 8);
                break;
            }
            default: { /* ZO_CMYK */
                const uint8_t *s = img->pix + (size_t)y * img->stride + (size_t)4 * x;
                zo_cmyk_to_rgba16(s[0], s[1], s[2], s[3], c);
                break;
            }
            }
            uint8_t *o = out + ((size_t)y * (size_t)W + (size_t)x) * 4;
            o[0] = (uint8_t)(c[0] >> 8);
            o[1] = (uint8_t)(c[1] >> 8);
            o[2] = (uint8_t)(c[2] >> 8);
            o[3] = (uint8_t)(c[3] >> 8);
        }
    }
}

int zo_load_rgba(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, int32_t *w, int32_t *h) {
    zo_image img;
    int e = zo_decode(data, len, &img);
    if (e != ZO_OK) return e;
    size_t need = (size_t)4 * (size_t)img.width * (size_t)img.height;
    if (w) *w = img.width;
    if (h) *h = img.height;
    if (out && out_cap >= need) {
        zo_rgba_pixels(&img, out);
    } else {
        uint8_t *tmp = (uint8_t *)malloc(need ? need : 1);
        if (!tmp) { zo_free(&img); return ZO_OutOfMemory; }
        zo_rgba_pixels(&img, tmp);
        free(tmp);
    }
    zo_free(&img);
    return ZO_OK;
}

/* test helpers: Color.toRGBA (color.zig:90-121) followed by the >> 8 of Image.rgbaPixels (image.zig:122-125) on n
 * samples.  ycc: n x {Y, Cb, Cr}; cmyk: n x {C, M, Y, K} as stored in Image{.CMYK}; rgba: n x 4 bytes */
void zo_ycbcr_to_rgba8_batch(const uint8_t *ycc, size_t n, uint8_t *rgba) {
    for (size_t i = 0; i < n; i++) {
        uint32_t c[4];
        zo_ycbcr_to_rgba16(ycc[3 * i], ycc[3 * i + 1], ycc[3 * i + 2], c);
        for (int k = 0; k < 4; k++) rgba[4 * i + k] = (uint8_t)(c[k] >> 8);
    }
}
void zo_cmyk_to_rgba8_batch(const uint8_t *cmyk, size_t n, uint8_t *rgba) {
    for (size_t i = 0; i < n; i++) {
        uint32_t c[4];
        zo_cmyk_to_rgba16(cmyk[4 * i], cmyk[4 * i + 1], cmyk[4 * i + 2], cmyk[4 * i + 3], c);
        for (int k = 0; k < 4; k++) rgba[4 * i + k] = (uint8_t)(c[k] >> 8);
    }
}

/* parity helper: jpeg.load + rgbaPixels of `data` compared with `got` (what the GPU path produced).
 * Returns 0 when identical, 1 when the bytes (or the length) differ, a negative error code (-ZO_*) when the
 * reference path fails on the file. */
int zo_compare_rgba(const uint8_t *data, size_t len, const uint8_t *got, size_t got_len) {
    zo_image img;
    int e = zo_decode(data, len, &img);
    if (e != ZO_OK) return -e;
    size_t need = (size_t)4 * (size_t)img.width * (size_t)img.height;
    int r = 1;
    if (need == got_len) {
        uint8_t *tmp = (uint8_t *)malloc(need ? need : 1);
        if (!tmp) { zo_free(&img); return -ZO_OutOfMemory; }
        zo_rgba_pixels(&img, tmp);
        r = memcmp(tmp, got, need) != 0;
        free(tmp);
    }
    zo_free(&img);
    return r;
}
