// zpix.hpp -- C++ host-side mirror of zpix's module API for the JPEG path, above the C ABI of
// include/zpix_cuda.h.  (The north star asks for this layer in Zig; zig/src/jpeg/batch.zig is that
// translation but cannot be compiled here, so the compiled, tested host layer is this header.)
//
//   zpix::jpeg::loadFromBuffer / load / decodeBatch / loadBatch / decodeConfig / probeBuffer
//       reference src/jpeg/root.zig:10-53, src/jpeg/decoder.zig:155,178
//   zpix::image::Image (tagged union Gray | YCbCr | RGBA | CMYK), Rectangle, Point, YCbCrSubsample, Config
//       reference src/image/image.zig:16-131,465-472, src/image/geometry.zig
//   zpix::color::Color::toRGBA
//       reference src/color/color.zig:31-132
//
// Same names, argument meaning and error behaviour: errors are thrown as zpix::Error carrying the Zig
// error name (error.UnexpectedEof -> "UnexpectedEof").  Pixel storage is std::vector<uint8_t>
// (the Zig `pixels: []u8` slice owned by the caller).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../include/zpix_cuda.h"

namespace zpix {

struct Error : std::runtime_error {
    int code;
    explicit Error(int c) : std::runtime_error(std::string("error.") + zpx_error_name(c)), code(c) {}
    const char* name() const { return zpx_error_name(code); }
};

namespace color {
// color.zig:13-23, the variants the JPEG path produces
struct Color {
    enum Kind { gray, ycbcr, cmyk, rgba } kind;
    uint8_t v[4];
    // color.zig:31-132: alpha-premultiplied 16-bit RGBA
    std::array<uint32_t, 4> toRGBA() const {
        switch (kind) {
            case gray: { uint32_t y = v[0] | (uint32_t)v[0] << 8; return {y, y, y, 0xffff}; }
            case rgba: return {v[0] | (uint32_t)v[0] << 8, v[1] | (uint32_t)v[1] << 8, v[2] | (uint32_t)v[2] << 8, v[3] | (uint32_t)v[3] << 8};
            case ycbcr: {
                const int32_t yy1 = (int32_t)v[0] * 0x10101, cb1 = (int32_t)v[1] - 128, cr1 = (int32_t)v[2] - 128;
                auto cl = [](int32_t x) -> uint32_t { return ((uint32_t)x & 0xff000000u) == 0 ? (uint32_t)(x >> 8) : (uint32_t)(~(x >> 31) & 0xffff); };
                return {cl(yy1 + 91881 * cr1), cl(yy1 - 22554 * cb1 - 46802 * cr1), cl(yy1 + 116130 * cb1), 0xffff};
            }
            default: {
                const uint32_t w = 0xffffu - v[3] * 0x101u;
                return {(0xffffu - v[0] * 0x101u) * w / 0xffffu, (0xffffu - v[1] * 0x101u) * w / 0xffffu, (0xffffu - v[2] * 0x101u) * w / 0xffffu, 0xffff};
            }
        }
    }
};
}  // namespace color

namespace image {
struct Point { int32_t x, y; };
struct Rectangle {
    Point min, max;
    static Rectangle init(int32_t x0, int32_t y0, int32_t x1, int32_t y1) {
        return {{x0 < x1 ? x0 : x1, y0 < y1 ? y0 : y1}, {x0 < x1 ? x1 : x0, y0 < y1 ? y1 : y0}};
    }
    int32_t dX() const { return max.x - min.x; }
    int32_t dY() const { return max.y - min.y; }
};
enum class YCbCrSubsample { Ratio444, Ratio422, Ratio420, Ratio440, Ratio411, Ratio410 };
struct Config { uint32_t width, height; enum { Gray, YCbCr } color_model; };

struct GrayImage { std::vector<uint8_t> pixels; size_t stride; Rectangle rect; };
struct RGBAImage { std::vector<uint8_t> pixels; size_t stride; Rectangle rect; };
struct CMYKImage { std::vector<uint8_t> pixels; size_t stride; Rectangle rect; };
struct YCbCrImage {
    size_t y_off, cb_off, cr_off;  // y / cb / cr slices are pixels[y_off..], [cb_off..], [cr_off..]
    size_t y_stride, c_stride;
    YCbCrSubsample subsample_ratio;
    Rectangle rect;
    std::vector<uint8_t> pixels;
};

// image.zig:24-131 (the variants jpeg.load can return)
struct Image {
    std::variant<GrayImage, YCbCrImage, RGBAImage, CMYKImage> v;
    std::vector<uint8_t> rgba;  // bytes of rgbaPixels(), computed on the GPU

    Rectangle bounds() const {
        return std::visit([](auto const& m) { return m.rect; }, v);
    }
    // image.zig:103-130.  Tight W*H*4, A = 255.
    const std::vector<uint8_t>& rgbaPixels() const { return rgba; }
};
}  // namespace image

namespace jpeg {

class Context {
   public:
    explicit Context(const std::vector<int32_t>& devices = {}) {
        int rc = zpx_ctx_create(devices.empty() ? nullptr : devices.data(), (int32_t)devices.size(), &h_);
        if (rc) throw Error(rc);
    }
    ~Context() { zpx_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    zpx_ctx* handle() const { return h_; }

   private:
    zpx_ctx* h_ = nullptr;
};

struct Result {
    int status = 0;  // 0 or the ZPX_E_* code (zpx_error_name gives the Zig error name)
    image::Image img;
};

enum class Output { rgba, native };

// NEW beside loadFromBuffer, ONE library call (zpx_decode_batch_rgba / zpx_decode_batch_native: header parse, upload,
// kernels and download pipelined in chunks).  Output::rgba: Image{.RGBA} per input holding jpeg.load(..).rgbaPixels()
// bytes.  Output::native: the variant jpeg.load itself returns (planes with makeImg's strides, decoder.zig:1708-1783).
// One corrupt input does not fail the batch: it gets its own status.
inline std::vector<Result> decodeBatch(Context& ctx, const std::vector<std::pair<const uint8_t*, size_t>>& buffers,
                                       Output output = Output::rgba) {
    const int n = (int)buffers.size();
    std::vector<const uint8_t*> ptrs(n);
    std::vector<size_t> lens(n);
    std::vector<zpx_image_info> infos(n);
    std::vector<Result> res(n);
    std::vector<uint8_t*> outs(n, nullptr);
    std::vector<int32_t> st(n, 0);
    for (int i = 0; i < n; i++) {
        ptrs[i] = buffers[i].first;
        lens[i] = buffers[i].second;
        zpx_image_info& info = infos[i];
        zpx_probe(ptrs[i], lens[i], &info);  // sizes from the header-only probe (decodeConfig)
        res[i].status = info.status;
        if (info.status) continue;
        const image::Rectangle rect = image::Rectangle::init(0, 0, info.width, info.height);
        const size_t w4 = (size_t)4 * info.width;
        if (output == Output::rgba || info.variant == ZPX_VARIANT_RGBA) {
            res[i].img.v = image::RGBAImage{std::vector<uint8_t>(output == Output::rgba ? info.rgba_len : info.native_len), w4, rect};
            outs[i] = std::get<image::RGBAImage>(res[i].img.v).pixels.data();
        } else if (info.variant == ZPX_VARIANT_GRAY) {
            res[i].img.v = image::GrayImage{std::vector<uint8_t>(info.native_len), (size_t)info.y_stride, rect};
            outs[i] = std::get<image::GrayImage>(res[i].img.v).pixels.data();
        } else if (info.variant == ZPX_VARIANT_CMYK) {
            res[i].img.v = image::CMYKImage{std::vector<uint8_t>(info.native_len), w4, rect};
            outs[i] = std::get<image::CMYKImage>(res[i].img.v).pixels.data();
        } else {
            res[i].img.v = image::YCbCrImage{0, (size_t)info.native_cb_off, (size_t)info.native_cr_off, (size_t)info.y_stride,
                                             (size_t)info.c_stride, (image::YCbCrSubsample)info.subsample_ratio, rect,
                                             std::vector<uint8_t>(info.native_len)};
            outs[i] = std::get<image::YCbCrImage>(res[i].img.v).pixels.data();
        }
    }
    const int rc = output == Output::rgba
                       ? zpx_decode_batch_rgba(ctx.handle(), ptrs.data(), lens.data(), n, outs.data(), nullptr, st.data())
                       : zpx_decode_batch_native(ctx.handle(), ptrs.data(), lens.data(), n, outs.data(), st.data());
    if (rc) throw Error(rc);
    for (int i = 0; i < n; i++) {
        res[i].status = st[i];
        if (!st[i] && output == Output::rgba) res[i].img.rgba = std::get<image::RGBAImage>(res[i].img.v).pixels;
    }
    return res;
}

// reference src/jpeg/root.zig:10
inline image::Image loadFromBuffer(Context& ctx, const uint8_t* data, size_t len) {
    auto r = decodeBatch(ctx, {{data, len}});
    if (r[0].status) throw Error(r[0].status);
    return std::move(r[0].img);
}

inline std::vector<uint8_t> readFile(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("error.FileNotFound: " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// reference src/jpeg/root.zig:36
inline image::Image load(Context& ctx, const std::string& path) {
    auto d = readFile(path);
    return loadFromBuffer(ctx, d.data(), d.size());
}

inline std::vector<Result> loadBatch(Context& ctx, const std::vector<std::string>& paths, Output output = Output::rgba) {
    std::vector<std::vector<uint8_t>> files;
    std::vector<std::pair<const uint8_t*, size_t>> bufs;
    for (auto& p : paths) files.push_back(readFile(p));
    for (auto& f : files) bufs.push_back({f.data(), f.size()});
    return decodeBatch(ctx, bufs, output);
}

// reference src/jpeg/decoder.zig:178-218; host only
inline image::Config decodeConfig(const uint8_t* data, size_t len) {
    zpx_image_info info;
    int rc = zpx_probe(data, len, &info);
    if (rc) throw Error(rc);
    return {(uint32_t)info.width, (uint32_t)info.height, info.num_components == 1 ? image::Config::Gray : image::Config::YCbCr};
}

// reference src/jpeg/root.zig:17-20
inline bool probeBuffer(const uint8_t* data, size_t len) { return len >= 2 && data[0] == 0xFF && data[1] == 0xD8; }

}  // namespace jpeg

// Batch dispatcher beside zpix.fromBuffer / fromFilePath (reference src/root.zig:24-40): same probing order (PNG,
// JPEG, QOI, BMP).  JPEGs of the batch are decoded together on the GPU and come back as the native variant; this
// C++ mirror has no CPU decoders for the other formats (zpix's own stay in charge of them): they, and anything
// unrecognised, get kUnknownImageFormat.
enum class Format { png, jpeg, qoi, bmp, unknown };
constexpr int kUnknownImageFormat = -1;  // error.UnknownImageFormat (src/root.zig:30,39)

inline Format probeBuffer(const uint8_t* d, size_t n) {
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n >= 8 && std::equal(png_sig, png_sig + 8, d)) return Format::png;
    if (jpeg::probeBuffer(d, n)) return Format::jpeg;
    if (n >= 4 && d[0] == 'q' && d[1] == 'o' && d[2] == 'i' && d[3] == 'f') return Format::qoi;
    if (n >= 2 && d[0] == 'B' && d[1] == 'M') return Format::bmp;
    return Format::unknown;
}

inline std::vector<jpeg::Result> fromBuffers(jpeg::Context& ctx, const std::vector<std::pair<const uint8_t*, size_t>>& buffers) {
    std::vector<jpeg::Result> res(buffers.size());
    std::vector<std::pair<const uint8_t*, size_t>> jpegs;
    std::vector<size_t> where;
    for (size_t i = 0; i < buffers.size(); i++) {
        if (probeBuffer(buffers[i].first, buffers[i].second) == Format::jpeg) {
            jpegs.push_back(buffers[i]);
            where.push_back(i);
        } else {
            res[i].status = kUnknownImageFormat;
        }
    }
    if (!jpegs.empty()) {
        auto r = jpeg::decodeBatch(ctx, jpegs, jpeg::Output::native);
        for (size_t k = 0; k < r.size(); k++) res[where[k]] = std::move(r[k]);
    }
    return res;
}

inline image::Image fromBuffer(jpeg::Context& ctx, const uint8_t* d, size_t n) {
    auto r = fromBuffers(ctx, {{d, n}});
    if (r[0].status == kUnknownImageFormat) throw std::runtime_error("error.UnknownImageFormat");
    if (r[0].status) throw Error(r[0].status);
    return std::move(r[0].img);
}
}  // namespace zpix
