// zpx_parse.cpp -- host-side marker / header parse of one JPEG.
//
// Walks the byte stream exactly the way the reference's marker loop does
// (src/jpeg/decoder.zig:220-355) and validates SOF/DQT/DHT/DRI/APPn/SOS headers
// with the same checks and the same error kinds (decoder.zig:490-697, 1026-1111,
// 1148-1255), but never decodes entropy-coded data: for every scan it only
// locates the restart intervals and the end of the entropy-coded segment
// (the job decoder.zig:1430-1452 + findRst :1671-1705 do while decoding).
//
// Why locating intervals without decoding is exact: the reference's bit reader
// never consumes a 0xFF that is followed by anything but 0x00
// (readByteStuffedByte, decoder.zig:712-749), so inside one interval it stops at
// or before the first such position (the "limit").  After the interval's MCUs,
// readFull/findRst (or the marker loop) walk the remaining bytes token by token
// (plain byte | FF 00 pair | FF fill) and therefore reach the limit in the same
// state no matter where the bit reader stopped.  Resuming the walk AT the limit
// is thus equivalent.
#include <string.h>

#include "zpx_internal.h"

const uint8_t zpx_unzig[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
};

namespace {

struct Parser {
    const uint8_t* d;
    size_t len;
    size_t pos = 0;
    ZpxParsed* out;

    // decoder state the header segments mutate
    ZpxHuffHost huff[2][4];
    int32_t quant[4][64];
    int restart_interval = 0;
    bool have_img = false;  // ycbcr_img != null (decoder.zig:1264)

    int err = 0;

    bool read_full(uint8_t* p, size_t n) {
        if (len - pos < n) {
            pos = len;
            err = ZPX_E_UnexpectedEof;
            return false;
        }
        memcpy(p, d + pos, n);
        pos += n;
        return true;
    }
    bool read_byte(uint8_t* b) { return read_full(b, 1); }
    bool ignore(int32_t n) {
        if (len - pos < (size_t)n) {
            pos = len;
            err = ZPX_E_UnexpectedEof;
            return false;
        }
        pos += (size_t)n;
        return true;
    }
    bool fail(int code) {
        err = code;
        return false;
    }

    // decoder.zig:490-618
    bool process_sof(int32_t n) {
        ZpxParsed& o = *out;
        if (o.ncomp != 0) return fail(ZPX_E_MultipleSofMarkers);
        int nc;
        switch (n) {
            case 6 + 3 * 1: nc = 1; break;
            case 6 + 3 * 3: nc = 3; break;
            case 6 + 3 * 4: nc = 4; break;
            default: return fail(ZPX_E_NumberComponents);
        }
        o.ncomp = nc;
        uint8_t tmp[32];
        if (!read_full(tmp, (size_t)n)) return false;
        if (tmp[0] != 8) return fail(ZPX_E_Precision);
        o.height = (tmp[1] << 8) + tmp[2];
        o.width = (tmp[3] << 8) + tmp[4];
        if (tmp[5] != nc) return fail(ZPX_E_SofWrongLength);
        for (int i = 0; i < nc; i++) {
            o.cid[i] = tmp[6 + 3 * i];
            for (int j = 0; j < i; j++)
                if (o.cid[i] == o.cid[j]) return fail(ZPX_E_RepeatedComponentIdentifier);
            o.tq[i] = tmp[8 + 3 * i];
            if (o.tq[i] > 3) return fail(ZPX_E_BadTqValue);
            uint8_t hv = tmp[7 + 3 * i];
            int h = hv >> 4, v = hv & 0x0f;
            if (h < 1 || 4 < h || v < 1 || 4 < v) return fail(ZPX_E_LumaChromaSubSamplingRatio);
            if (h == 3 || v == 3) return fail(ZPX_E_LumaChromaSubSamplingRatio);
            switch (nc) {
                case 1:
                    h = 1;
                    v = 1;
                    break;
                case 3:
                    if (i == 0) {
                        if (v == 4) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    } else if (i == 1) {
                        if (o.h[0] % h != 0 || o.v[0] % v != 0) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    } else {
                        if (o.h[1] != h || o.v[1] != v) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    }
                    break;
                case 4:
                    if (i == 0) {
                        if (hv != 0x11 && hv != 0x22) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    } else if (i == 1 || i == 2) {
                        if (hv != 0x11) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    } else {
                        if (o.h[0] != h || o.v[0] != v) return fail(ZPX_E_LumaChromaSubSamplingRatio);
                    }
                    break;
            }
            o.h[i] = h;
            o.v[i] = v;
        }
        return true;
    }

    // decoder.zig:621-627
    bool process_dri(int32_t n) {
        if (n != 2) return fail(ZPX_E_DriWrongLength);
        uint8_t t[2];
        if (!read_full(t, 2)) return false;
        restart_interval = (t[0] << 8) + t[1];
        return true;
    }

    // decoder.zig:629-666
    bool process_dqt(int32_t n) {
        int32_t ln = n;
        while (ln > 0) {
            ln -= 1;
            uint8_t qi;
            if (!read_byte(&qi)) return false;
            int tq = qi & 0x0f;
            if (tq > 3) return fail(ZPX_E_BadTqValue);
            uint8_t tmp[128];
            bool stop = false;
            switch (qi >> 4) {
                case 0:
                    if (ln < 64) { stop = true; break; }
                    ln -= 64;
                    if (!read_full(tmp, 64)) return false;
                    for (int i = 0; i < 64; i++) quant[tq][i] = tmp[i];
                    break;
                case 1:
                    if (ln < 128) { stop = true; break; }
                    ln -= 128;
                    if (!read_full(tmp, 128)) return false;
                    for (int i = 0; i < 64; i++) quant[tq][i] = (tmp[2 * i] << 8) | tmp[2 * i + 1];
                    break;
                default: return fail(ZPX_E_BadPqValue);
            }
            if (stop) break;
        }
        if (ln != 0) return fail(ZPX_E_DqtWrongLength);
        return true;
    }

    // decoder.zig:668-680
    bool process_app0(int32_t n) {
        if (n < 5) return ignore(n);
        uint8_t t[5];
        if (!read_full(t, 5)) return false;
        out->jfif = t[0] == 'J' && t[1] == 'F' && t[2] == 'I' && t[3] == 'F' && t[4] == 0;
        return ignore(n - 5);
    }

    // decoder.zig:682-697
    bool process_app14(int32_t n) {
        if (n < 12) return ignore(n);
        uint8_t t[12];
        if (!read_full(t, 12)) return false;
        if (t[0] == 'A' && t[1] == 'd' && t[2] == 'o' && t[3] == 'b' && t[4] == 'e') {
            out->adobe_valid = true;
            out->adobe_transform = t[11];
        }
        return ignore(n - 12);
    }

    // decoder.zig:1026-1111 (table construction itself happens in zpx_build_huff_dev)
    bool process_dht(int32_t n) {
        int32_t ln = n;
        while (ln > 0) {
            if (ln < 17) return fail(ZPX_E_DhtWrongLength);
            uint8_t t[17];
            if (!read_full(t, 17)) return false;
            int tc = t[0] >> 4;
            if (tc > 1) return fail(ZPX_E_BadTcValue);
            int th = t[0] & 0x0f;
            if (th > 3 || (out->baseline && th > 1)) return fail(ZPX_E_BadThValue);
            ZpxHuffHost& h = huff[tc][th];
            h.num_codes = 0;
            for (int i = 0; i < 16; i++) {
                h.counts[i] = t[i + 1];
                h.num_codes += t[i + 1];
            }
            h.defined = false;
            if (h.num_codes == 0) return fail(ZPX_E_HuffZeroLength);
            if (h.num_codes > 256) return fail(ZPX_E_HuffTooLong);
            ln -= h.num_codes + 17;
            if (ln < 0) return fail(ZPX_E_DhtWrongLength);
            if (!read_full(h.vals, (size_t)h.num_codes)) return false;
            h.defined = true;
            // the reference's 8-bit LUT fill indexes out of bounds (panic) for over-subscribed codes
            uint32_t code = 0;
            for (int i = 0; i < 8; i++) {
                code <<= 1;
                for (int j = 0; j < h.counts[i]; j++) {
                    uint32_t base = code << (7 - i);
                    if ((base | ((1u << (7 - i)) - 1)) >= 256) return fail(ZPX_E_MALFORMED_TABLE);
                    code += 1;
                }
            }
        }
        return true;
    }

    // first 0xFF followed by a byte != 0x00 at or after `from`; len if none.
    // Also counts the FF 00 pairs on the way (*n_stuffed) and, when `segs` is given, cuts [from, limit) into pieces
    // of about ZPX_SEG_BYTES raw bytes that never split a pair: the work units of the unstuffing kernel (zpx_k0.cu).
    size_t find_limit(size_t from, bool* eof_limit, uint32_t* n_stuffed, std::vector<ZpxSegHost>* segs) const {
        size_t i = from, seg_begin = from;
        uint32_t stuffed = 0, seg_uoff = 0;
        size_t limit;
        for (;;) {
            const uint8_t* p = i < len ? (const uint8_t*)memchr(d + i, 0xff, len - i) : nullptr;
            const size_t j = p ? (size_t)(p - d) : len;  // [i, j) holds no 0xFF
            if (segs) {
                while (j - seg_begin > ZPX_SEG_BYTES) {
                    size_t c = seg_begin + ZPX_SEG_BYTES;
                    if (c < i) c = i;  // i = just past the last pair: never cut between an 0xFF and its 0x00
                    segs->push_back({seg_begin, (uint32_t)(c - seg_begin), seg_uoff});
                    seg_begin = c;
                    seg_uoff = (uint32_t)(c - from) - stuffed;
                }
            }
            if (!p) {
                *eof_limit = true;
                limit = len;
                break;
            }
            if (j + 1 >= len) {  // 0xFF is the last byte: the reader hits end of stream looking for the 0x00
                *eof_limit = true;
                limit = j;
                break;
            }
            if (d[j + 1] == 0x00) {
                stuffed++;
                i = j + 2;
                continue;
            }
            *eof_limit = false;
            limit = j;
            break;
        }
        if (segs && limit > seg_begin) segs->push_back({seg_begin, (uint32_t)(limit - seg_begin), seg_uoff});
        *n_stuffed = stuffed;
        return limit;
    }

    // decoder.zig:1436-1439 + findRst :1671-1705, started at `q` (see file header).
    // Returns 0 and sets *next to the byte after the marker, or the Zig error.
    int find_rst_from(size_t q, uint8_t expected, size_t* next) {
        uint8_t tmp[2];
        size_t p = q;
        if (len - p < 2) return ZPX_E_UnexpectedEof;
        tmp[0] = d[p];
        tmp[1] = d[p + 1];
        p += 2;
        if (tmp[0] == 0xff && tmp[1] == expected) {
            *next = p;
            return 0;
        }
        for (;;) {
            size_t i = 0;
            if (tmp[0] == 0xff) {
                if (tmp[1] == expected) {
                    *next = p;
                    return 0;
                } else if (tmp[1] == 0xff) {
                    i = 1;
                } else if (tmp[1] != 0x00) {
                    return ZPX_E_BadRSTMarker;
                }
            } else if (tmp[1] == 0xff) {
                tmp[0] = 0xff;
                i = 1;
            }
            size_t need = 2 - i;
            if (len - p < need) return ZPX_E_UnexpectedEof;
            for (size_t k = i; k < 2; k++) tmp[k] = d[p++];
        }
    }

    // decoder.zig:1148-1292 (header part) + interval location
    bool process_sos(int32_t n) {
        ZpxParsed& o = *out;
        if (o.ncomp == 0) return fail(ZPX_E_MissingSosMarker);
        if (n < 6 || 4 + 2 * o.ncomp < n || n % 2 != 0) return fail(ZPX_E_SosWrongLength);
        uint8_t tmp[16];
        if (!read_full(tmp, (size_t)n)) return false;
        int n_comp = tmp[0];
        if (n != 4 + 2 * n_comp) return fail(ZPX_E_SosWrongLength);
        ZpxScanHost sc;
        sc.ncomp = n_comp;
        int total_hv = 0;
        for (int i = 0; i < n_comp; i++) {
            uint8_t cs = tmp[1 + 2 * i];
            int ci = -1;
            for (int j = 0; j < o.ncomp; j++)
                if (cs == o.cid[j]) { ci = j; break; }
            if (ci < 0) return fail(ZPX_E_UnknownComponentSelector);
            sc.comp[i] = ci;
            for (int j = 0; j < i; j++)
                if (sc.comp[i] == sc.comp[j]) return fail(ZPX_E_RepeatedComponentIdentifier);
            total_hv += o.h[ci] * o.v[ci];
            sc.td[i] = tmp[2 + 2 * i] >> 4;
            if (sc.td[i] > 3 || (o.baseline && sc.td[i] > 1)) return fail(ZPX_E_BadTdValue);
            sc.ta[i] = tmp[2 + 2 * i] & 0x0f;
            if (sc.ta[i] > 3 || (o.baseline && sc.ta[i] > 1)) return fail(ZPX_E_BadTaValue);
        }
        if (o.ncomp > 1 && total_hv > 10) return fail(ZPX_E_SamplingFactorsTooLarge);
        sc.ss = 0;
        sc.se = 63;
        sc.ah = sc.al = 0;
        if (o.progressive) {
            sc.ss = tmp[1 + 2 * n_comp];
            sc.se = tmp[2 + 2 * n_comp];
            sc.ah = tmp[3 + 2 * n_comp] >> 4;
            sc.al = tmp[3 + 2 * n_comp] & 0x0f;
            if ((sc.ss == 0 && sc.se != 0) || sc.ss > sc.se || 64 <= sc.se) return fail(ZPX_E_BadSpectralSelection);
            if (sc.ss != 0 && n_comp != 1) return fail(ZPX_E_ProgressiveACCoefficientsForMoreThanOneComponent);
            if (sc.ah != 0 && sc.ah != sc.al + 1) return fail(ZPX_E_BadSuccessiveApproximation);
        }
        int h0 = o.h[0], v0 = o.v[0];
        int mxx = (o.width + 8 * h0 - 1) / (8 * h0);
        int myy = (o.height + 8 * v0 - 1) / (8 * v0);
        o.mxx = mxx;
        o.myy = myy;
        if (!have_img) {
            // makeImg (decoder.zig:1708-1783): subImage of an empty rectangle -> CreateImageFailed
            if (o.width == 0 || o.height == 0) return fail(ZPX_E_CreateImageFailed);
            if (o.ncomp != 1) have_img = true;
        }
        sc.restart_interval = restart_interval;
        for (int i = 0; i < n_comp; i++) {
            sc.dc[i] = huff[0][sc.td[i]];
            sc.ac[i] = huff[1][sc.ta[i]];
            memcpy(sc.quant[i], quant[o.tq[sc.comp[i]]], sizeof(sc.quant[i]));
        }

        // ---- locate the restart intervals (decoder.zig:1430-1452) ----
        uint32_t total_mcu = (uint32_t)mxx * (uint32_t)myy;
        uint32_t ri = (uint32_t)restart_interval;
        uint32_t n_int = ri > 0 ? (total_mcu + ri - 1) / ri : 1;
        size_t p = pos;
        uint8_t expected = 0xd0;
        for (uint32_t k = 0; k < n_int; k++) {
            ZpxIntervalHost iv;
            iv.start = p;
            iv.seg_first = (uint32_t)sc.segs.size();
            iv.limit = find_limit(p, &iv.eof_limit, &iv.n_stuffed, &sc.segs);
            iv.n_segs = (uint32_t)sc.segs.size() - iv.seg_first;
            iv.first_mcu = ri > 0 ? k * ri : 0;
            iv.n_mcu = ri > 0 ? (total_mcu - iv.first_mcu < ri ? total_mcu - iv.first_mcu : ri) : total_mcu;
            sc.intervals.push_back(iv);
            if (k + 1 < n_int) {
                size_t next = 0;
                int e = iv.eof_limit ? ZPX_E_UnexpectedEof : find_rst_from(iv.limit, expected, &next);
                if (e != 0) {
                    sc.pending_err = e;
                    sc.err_after_interval = (int)k;
                    p = len;
                    break;
                }
                p = next;
                expected = expected == 0xd7 ? 0xd0 : (uint8_t)(expected + 1);
            } else {
                p = iv.limit;
            }
        }
        pos = p;
        o.saw_sos = true;
        o.scans.push_back(std::move(sc));
        if (o.scans.back().pending_err) return fail(o.scans.back().pending_err);
        return true;
    }

    // decoder.zig:220-355.  Returns false with err set on the first error.
    bool run(bool config_only) {
        ZpxParsed& o = *out;
        memset(quant, 0, sizeof(quant));
        uint8_t tmp[2];
        if (!read_full(tmp, 2)) return false;
        if (tmp[0] != 0xff || tmp[1] != 0xd8) return fail(ZPX_E_InvalidSOIMarker);
        for (;;) {
            if (!read_full(tmp, 2)) return false;
            while (tmp[0] != 0xff) {
                tmp[0] = tmp[1];
                if (!read_byte(&tmp[1])) return false;
            }
            uint8_t marker = tmp[1];
            if (marker == 0) continue;
            while (marker == 0xff)
                if (!read_byte(&marker)) return false;
            if (marker == 0xd9) break;
            if (0xd0 <= marker && marker <= 0xd7) continue;
            if (!read_full(tmp, 2)) return false;
            int32_t n = (tmp[0] << 8) + tmp[1] - 2;
            if (n < 0) return fail(ZPX_E_ShortSegmentLength);
            switch (marker) {
                case 0xc0:
                case 0xc1:
                case 0xc2:
                    // (the reference sets these before processSof, decoder.zig:303-305; a second SOF then fails with
                    // MultipleSofMarkers and nothing reads them again -- here the scans accepted so far are still
                    // to be decoded, so a rejected SOF must not change how)
                    if (o.ncomp == 0) {
                        o.baseline = marker == 0xc0;
                        o.progressive = marker == 0xc2;
                    }
                    if (!process_sof(n)) return false;
                    if (config_only && o.jfif) return fail(ZPX_E_ConfigOnly);
                    break;
                case 0xdb:
                    if (config_only ? !ignore(n) : !process_dqt(n)) return false;
                    break;
                case 0xdd:
                    if (config_only ? !ignore(n) : !process_dri(n)) return false;
                    break;
                case 0xc4:
                    if (config_only ? !ignore(n) : !process_dht(n)) return false;
                    break;
                case 0xda:
                    if (config_only) return fail(ZPX_E_ConfigOnly);
                    if (!process_sos(n)) return false;
                    break;
                case 0xe0:
                    if (!process_app0(n)) return false;
                    break;
                case 0xee:
                    if (!process_app14(n)) return false;
                    break;
                default:
                    if ((0xe0 <= marker && marker <= 0xef) || marker == 0xfe) {
                        if (!ignore(n)) return false;
                    } else if (marker < 0xc0) {
                        return fail(ZPX_E_UnknownMarker);
                    } else {
                        return fail(ZPX_E_UnsupportedMarker);
                    }
            }
        }
        return true;
    }
};

}  // namespace

void zpx_derive(ZpxParsed* o) {
    // decoder.zig:361-370 / 699-709 / 792-811 / 1743-1753
    if (o->ncomp == 1) {
        o->mode = ZPX_MODE_GRAY;
        o->variant = ZPX_VARIANT_GRAY;
        return;
    }
    if (o->ncomp < 3) return;
    if (o->h[1] <= 0 || o->v[1] <= 0 || o->h[0] <= 0 || o->v[0] <= 0) return;  // SOF rejected half way
    int hr = o->h[0] / o->h[1], vr = o->v[0] / o->v[1];
    switch (hr << 4 | vr) {
        case 0x11: o->ratio = ZPX_RATIO_444; break;
        case 0x12: o->ratio = ZPX_RATIO_440; break;
        case 0x21: o->ratio = ZPX_RATIO_422; break;
        case 0x22: o->ratio = ZPX_RATIO_420; break;
        case 0x41: o->ratio = ZPX_RATIO_411; break;
        case 0x42: o->ratio = ZPX_RATIO_410; break;
        default: o->ratio = ZPX_RATIO_444; break;
    }
    if (o->ncomp == 4) {
        o->variant = ZPX_VARIANT_CMYK;
        o->mode = o->adobe_transform != 0 ? ZPX_MODE_YCCK : ZPX_MODE_CMYK;
        return;
    }
    bool is_rgb = false;
    if (!o->jfif) {
        if (o->adobe_valid && o->adobe_transform == 0) is_rgb = true;
        else is_rgb = o->cid[0] == 'R' && o->cid[1] == 'G' && o->cid[2] == 'B';
    }
    o->mode = is_rgb ? ZPX_MODE_RGB : ZPX_MODE_YCBCR;
    o->variant = is_rgb ? ZPX_VARIANT_RGBA : ZPX_VARIANT_YCBCR;
}

void zpx_parse_jpeg(const uint8_t* data, size_t len, bool config_only, ZpxParsed* out) {
    Parser p;
    p.d = data;
    p.len = len;
    p.out = out;
    bool ok = p.run(config_only);
    memcpy(out->final_quant, p.quant, sizeof(p.quant));
    if (out->ncomp >= 1) zpx_derive(out);
    if (config_only) {
        // decodeConfig (decoder.zig:178-218)
        if (!ok && p.err != ZPX_E_ConfigOnly) out->status = p.err;
        else if (out->ncomp != 1 && out->ncomp != 3 && out->ncomp != 4) out->status = ZPX_E_InvalidSOIMarker;
        return;
    }
    if (!ok) {
        // Errors raised after at least one scan was accepted come, in the reference, after that
        // scan's entropy decode: the device may still find an earlier error inside it.
        if (out->saw_sos && !out->scans.empty() && !out->scans.back().intervals.empty()) {
            if (out->scans.back().pending_err == 0) out->trailing_err = p.err;
        } else {
            out->status = p.err;
        }
        return;
    }
    if (!out->saw_sos) {
        out->status = ZPX_E_MissingSosMarker;
        return;
    }
    // decoder.zig:792-795: 4 components without an Adobe marker
    if (out->ncomp == 4 && !out->adobe_valid) out->trailing_err = ZPX_E_UnsupportedColorModel;
}

// decoder.zig:1070-1109 restated for a ZPX_LUT_BITS-bit first level
// fields of ZpxHuffDev::fast for a code of length len that decodes to sym
uint32_t zpx_fast_entry(bool is_ac, int len, int sym) {
    uint32_t size, adv, special = 0;
    if (!is_ac) {
        size = (uint32_t)sym;
        adv = 1;
        if (sym > 16) { special = 1; size = 0; }
    } else {
        const uint32_t r = (uint32_t)sym >> 4, s2 = (uint32_t)sym & 15;
        if (s2 != 0) { size = s2; adv = r + 1; special = s2 >= 13 ? 1 : 0; }
        else if (r == 15) { size = 0; adv = 16; }
        else if (r == 0) { size = 0; adv = 64; }
        else { size = 0; adv = 64; special = 1; }
    }
    return ZPX_FE((uint32_t)len + size, len, size, adv, special);
}

void zpx_build_huff_dev(const ZpxHuffHost& h, bool is_ac, ZpxHuffDev* o, int* malformed) {
    memset(o, 0, sizeof(*o));
    *malformed = 0;
    if (!h.defined) return;
    o->defined = 1;
    memcpy(o->vals, h.vals, 256);
    uint32_t code = 0;
    int index = 0;
    for (int l = 1; l <= 16; l++) {
        int cnt = h.counts[l - 1];
        if (cnt == 0) {
            o->limit[l] = 0;
            o->valoff[l] = 0;
        } else {
            // codes of this length: code .. code+cnt-1 (min_codes / max_codes, decoder.zig:1102-1104)
            if (code + (uint32_t)cnt > (1u << l)) *malformed = 1;
            uint64_t lim = ((uint64_t)(code + (uint32_t)cnt)) << (16 - l);
            o->limit[l] = (uint32_t)(lim > 0x10000 ? 0x10000 : lim);
            o->valoff[l] = index - (int32_t)code;
            if (l <= ZPX_LUT_BITS) {
                for (int j = 0; j < cnt; j++) {
                    uint32_t c = code + (uint32_t)j;
                    uint32_t base = c << (ZPX_LUT_BITS - l);
                    uint16_t v = (uint16_t)((uint16_t)h.vals[index + j] << 8 | (uint16_t)l);
                    const uint32_t f = zpx_fast_entry(is_ac, l, h.vals[index + j]);
                    for (uint32_t k = 0; k < (1u << (ZPX_LUT_BITS - l)); k++)
                        if ((base | k) < ZPX_LUT_SIZE) {
                            o->lut[base | k] = v;
                            o->fast[base | k] = f;
                        }
                }
            }
            code += (uint32_t)cnt;
            index += cnt;
        }
        code <<= 1;
    }
}

void zpx_fill_info(const ZpxParsed& p, zpx_image_info* info) {
    memset(info, 0, sizeof(*info));
    info->status = p.status;
    info->width = p.width;
    info->height = p.height;
    info->num_components = p.ncomp;
    info->variant = p.variant;
    info->subsample_ratio = p.ratio;
    info->progressive = p.progressive ? 1 : 0;
    info->restart_interval = p.scans.empty() ? 0 : p.scans[0].restart_interval;
    // the MCU grid is fixed by the frame header (decoder.zig:1258-1263 computes it at each SOS): a header-only
    // probe (zpx_probe stops before the first SOS) reports the same sizes as the full parse
    int mxx = p.mxx, myy = p.myy;
    if (mxx == 0 && p.ncomp >= 1 && p.h[0] > 0 && p.v[0] > 0 && p.width > 0 && p.height > 0) {
        mxx = (p.width + 8 * p.h[0] - 1) / (8 * p.h[0]);
        myy = (p.height + 8 * p.v[0] - 1) / (8 * p.v[0]);
    }
    info->mxx = mxx;
    info->myy = myy;
    info->rgba_len = (uint64_t)4 * (uint64_t)p.width * (uint64_t)p.height;
    if (p.ncomp == 1) {
        info->y_stride = 8 * mxx;
        info->native_len = (uint64_t)(8 * mxx) * (uint64_t)(8 * myy);
    } else if (p.ncomp >= 3 && mxx > 0) {
        // makeImg + yCbCrSize (decoder.zig:1755-1760, image.zig:521-555)
        uint64_t w = (uint64_t)8 * p.h[0] * mxx, h = (uint64_t)8 * p.v[0] * myy, cw, ch;
        switch (p.ratio) {
            case ZPX_RATIO_422: cw = (w + 1) / 2; ch = h; break;
            case ZPX_RATIO_420: cw = (w + 1) / 2; ch = (h + 1) / 2; break;
            case ZPX_RATIO_440: cw = w; ch = (h + 1) / 2; break;
            case ZPX_RATIO_411: cw = (w + 3) / 4; ch = h; break;
            case ZPX_RATIO_410: cw = (w + 3) / 4; ch = (h + 1) / 2; break;
            default: cw = w; ch = h; break;
        }
        info->y_stride = (int32_t)w;
        info->c_stride = (int32_t)cw;
        if (p.variant == ZPX_VARIANT_YCBCR) {
            info->native_len = w * h + 2 * cw * ch;
            info->native_cb_off = w * h;
            info->native_cr_off = w * h + cw * ch;
        } else {
            info->native_len = info->rgba_len;
        }
    }
}
