//! Batch JPEG decode on the GPU: the new entry points that sit beside `jpeg.load` /
//! `jpeg.loadFromBuffer` (reference src/jpeg/root.zig:10,36) and return `image.Image{ .RGBA }`
//! holding exactly the bytes `jpeg.load(..).rgbaPixels()` yields.
//!
//! Host code stays in Zig; everything below the header parse runs in libzpixcuda.so through the
//! C ABI of include/zpix_cuda.h (no Triton, no CPU fallback).  NOT COMPILED in the build
//! environment of this repository (no Zig toolchain there, see DESIGN.md): this file is the thin,
//! reviewable translation of the header a zpix maintainer drops into src/jpeg/.
//!
//! build.zig additions (next to the jpeg module, reference build.zig:44-68):
//!     jpeg_mod.addIncludePath(b.path("include"));          // zpix_cuda.h
//!     jpeg_mod.addLibraryPath(b.path("zpix_b200"));        // libzpixcuda.so
//!     jpeg_mod.linkSystemLibrary("zpixcuda", .{});
//!     jpeg_mod.linkSystemLibrary("cudart", .{});
//!     jpeg_mod.link_libcpp = true;
const std = @import("std");
const image = @import("image");

const c = @cImport({
    @cInclude("zpix_cuda.h");
});

pub const BatchOptions = struct {
    /// CUDA device ordinals to spread the batch over (images are independent: the host scheduler
    /// partitions the batch, there is no collective).  Empty = device 0.
    devices: []const i32 = &.{},
    /// Allocate the pixel slices from pinned host memory (zpx_host_alloc) instead of `al`.
    /// Faster D2H; such images must be released with `freePinned`, not `Image.free`.
    pinned: bool = false,
};

/// Every Zig error of src/jpeg/decoder.zig, in the order of the ZPX_E_* codes (1..42).
pub const DecodeError = error{
    UnexpectedEof, InvalidSOIMarker, ShortSegmentLength, UnknownMarker, UnsupportedMarker, MissingSosMarker,
    MultipleSofMarkers, NumberComponents, Precision, SofWrongLength, RepeatedComponentIdentifier, BadTqValue,
    LumaChromaSubSamplingRatio, DriWrongLength, BadPqValue, DqtWrongLength, MissingFF00, UnsupportedColorModel,
    UninitializedHuffmanTable, BadHuffmanCode, DhtWrongLength, BadTcValue, BadThValue, HuffZeroLength, HuffTooLong,
    SosWrongLength, UnknownComponentSelector, BadTdValue, BadTaValue, SamplingFactorsTooLarge, BadSpectralSelection,
    ProgressiveACCoefficientsForMoreThanOneComponent, BadSuccessiveApproximation, ExcessiveDCComponent,
    UnexpectedHuffmanCode, TooManyCoefficients, BadRSTMarker, CreateImageFailed, UnsupportedComponent,
    InvalidImageType, ConfigOnly, OutOfMemory,
    // library-specific (codes >= 100)
    CudaFailure, NoCudaDevice, InvalidArgument, BadState, CoefficientOutOfRange, UnsupportedStream, MalformedHuffmanTable,
};

fn toError(code: i32) DecodeError {
    return switch (code) {
        c.ZPX_E_UnexpectedEof => error.UnexpectedEof,
        c.ZPX_E_InvalidSOIMarker => error.InvalidSOIMarker,
        c.ZPX_E_ShortSegmentLength => error.ShortSegmentLength,
        c.ZPX_E_UnknownMarker => error.UnknownMarker,
        c.ZPX_E_UnsupportedMarker => error.UnsupportedMarker,
        c.ZPX_E_MissingSosMarker => error.MissingSosMarker,
        c.ZPX_E_MultipleSofMarkers => error.MultipleSofMarkers,
        c.ZPX_E_NumberComponents => error.NumberComponents,
        c.ZPX_E_Precision => error.Precision,
        c.ZPX_E_SofWrongLength => error.SofWrongLength,
        c.ZPX_E_RepeatedComponentIdentifier => error.RepeatedComponentIdentifier,
        c.ZPX_E_BadTqValue => error.BadTqValue,
        c.ZPX_E_LumaChromaSubSamplingRatio => error.LumaChromaSubSamplingRatio,
        c.ZPX_E_DriWrongLength => error.DriWrongLength,
        c.ZPX_E_BadPqValue => error.BadPqValue,
        c.ZPX_E_DqtWrongLength => error.DqtWrongLength,
        c.ZPX_E_MissingFF00 => error.MissingFF00,
        c.ZPX_E_UnsupportedColorModel => error.UnsupportedColorModel,
        c.ZPX_E_UninitializedHuffmanTable => error.UninitializedHuffmanTable,
        c.ZPX_E_BadHuffmanCode => error.BadHuffmanCode,
        c.ZPX_E_DhtWrongLength => error.DhtWrongLength,
        c.ZPX_E_BadTcValue => error.BadTcValue,
        c.ZPX_E_BadThValue => error.BadThValue,
        c.ZPX_E_HuffZeroLength => error.HuffZeroLength,
        c.ZPX_E_HuffTooLong => error.HuffTooLong,
        c.ZPX_E_SosWrongLength => error.SosWrongLength,
        c.ZPX_E_UnknownComponentSelector => error.UnknownComponentSelector,
        c.ZPX_E_BadTdValue => error.BadTdValue,
        c.ZPX_E_BadTaValue => error.BadTaValue,
        c.ZPX_E_SamplingFactorsTooLarge => error.SamplingFactorsTooLarge,
        c.ZPX_E_BadSpectralSelection => error.BadSpectralSelection,
        c.ZPX_E_ProgressiveACCoefficientsForMoreThanOneComponent => error.ProgressiveACCoefficientsForMoreThanOneComponent,
        c.ZPX_E_BadSuccessiveApproximation => error.BadSuccessiveApproximation,
        c.ZPX_E_ExcessiveDCComponent => error.ExcessiveDCComponent,
        c.ZPX_E_UnexpectedHuffmanCode => error.UnexpectedHuffmanCode,
        c.ZPX_E_TooManyCoefficients => error.TooManyCoefficients,
        c.ZPX_E_BadRSTMarker => error.BadRSTMarker,
        c.ZPX_E_CreateImageFailed => error.CreateImageFailed,
        c.ZPX_E_UnsupportedComponent => error.UnsupportedComponent,
        c.ZPX_E_InvalidImageType => error.InvalidImageType,
        c.ZPX_E_ConfigOnly => error.ConfigOnly,
        c.ZPX_E_OutOfMemory => error.OutOfMemory,
        c.ZPX_E_CUDA => error.CudaFailure,
        c.ZPX_E_NO_DEVICE => error.NoCudaDevice,
        c.ZPX_E_BAD_STATE => error.BadState,
        c.ZPX_E_COEF_RANGE => error.CoefficientOutOfRange,
        c.ZPX_E_UNSUPPORTED_STREAM => error.UnsupportedStream,
        c.ZPX_E_MALFORMED_TABLE => error.MalformedHuffmanTable,
        else => error.InvalidArgument,
    };
}

/// One decoded image or the error the reference decoder would have returned for that input.
pub const Result = union(enum) {
    ok: image.Image,
    err: DecodeError,
};

/// Decode a batch of in-memory JPEGs.  The returned slice and every `.ok` image's pixels are
/// owned by the caller (`Image.free(al)` each, then `al.free(results)`), exactly like the images
/// `jpeg.loadFromBuffer` returns.  One corrupt input does not fail the batch.
pub fn decodeBatch(al: std.mem.Allocator, buffers: []const []const u8, opts: BatchOptions) ![]Result {
    var ctx: ?*c.zpx_ctx = null;
    var rc = c.zpx_ctx_create(if (opts.devices.len == 0) null else opts.devices.ptr, @intCast(opts.devices.len), &ctx);
    if (rc != c.ZPX_OK) return toError(rc);
    defer c.zpx_ctx_destroy(ctx);

    const n = buffers.len;
    const ptrs = try al.alloc([*c]const u8, n);
    defer al.free(ptrs);
    const lens = try al.alloc(usize, n);
    defer al.free(lens);
    for (buffers, 0..) |buf, i| {
        ptrs[i] = buf.ptr;
        lens[i] = buf.len;
    }

    var batch: ?*c.zpx_batch = null;
    rc = c.zpx_batch_open(ctx, ptrs.ptr, lens.ptr, @intCast(n), &batch);
    if (rc != c.ZPX_OK) return toError(rc);
    defer c.zpx_batch_close(batch);

    // the caller's allocator owns every output slice: size them from the header parse
    const results = try al.alloc(Result, n);
    errdefer al.free(results);
    const outs = try al.alloc([*c]u8, n);
    defer al.free(outs);
    const status = try al.alloc(i32, n);
    defer al.free(status);
    for (0..n) |i| {
        var info: c.zpx_image_info = undefined;
        _ = c.zpx_batch_info(batch, @intCast(i), &info);
        outs[i] = null;
        if (info.status != c.ZPX_OK) {
            results[i] = .{ .err = toError(info.status) };
            continue;
        }
        const rect = image.Rectangle.init(0, 0, info.width, info.height);
        var img = try image.RGBAImage.init(al, rect); // pixels.len == 4*W*H, stride == 4*W
        outs[i] = img.pixels.ptr;
        results[i] = .{ .ok = .{ .RGBA = img } };
    }

    rc = c.zpx_batch_upload(batch);
    if (rc == c.ZPX_OK) rc = c.zpx_batch_decode(batch, null);
    if (rc == c.ZPX_OK) rc = c.zpx_batch_fetch_rgba(batch, outs.ptr, null, status.ptr);
    if (rc != c.ZPX_OK) {
        for (results) |r| switch (r) {
            .ok => |img| img.free(al),
            .err => {},
        };
        return toError(rc);
    }
    for (0..n) |i| {
        if (status[i] != c.ZPX_OK) {
            switch (results[i]) {
                .ok => |img| img.free(al),
                .err => {},
            }
            results[i] = .{ .err = toError(status[i]) };
        }
    }
    return results;
}

/// Same, reading the files first (beside `jpeg.load`, reference src/jpeg/root.zig:36).
pub fn loadBatch(al: std.mem.Allocator, paths: []const []const u8, opts: BatchOptions) ![]Result {
    const bufs = try al.alloc([]const u8, paths.len);
    var loaded: usize = 0;
    defer {
        for (bufs[0..loaded]) |b| al.free(b);
        al.free(bufs);
    }
    for (paths, 0..) |p, i| {
        bufs[i] = try std.fs.cwd().readFileAlloc(al, p, std.math.maxInt(usize));
        loaded = i + 1;
    }
    return decodeBatch(al, bufs, opts);
}

/// Header-only probe (decodeConfig, reference decoder.zig:178) through the same parser the batch
/// path uses; no GPU involved.
pub fn decodeConfigBuffer(buffer: []const u8) !image.Config {
    var info: c.zpx_image_info = undefined;
    const rc = c.zpx_probe(buffer.ptr, buffer.len, &info);
    if (rc != c.ZPX_OK) return toError(rc);
    return image.Config{
        .width = @intCast(info.width),
        .height = @intCast(info.height),
        .color_model = if (info.num_components == 1) .Gray else .YCbCr,
    };
}
