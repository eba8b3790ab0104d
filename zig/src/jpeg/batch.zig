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
    /// partitions the batch, there is no collective).  Empty = device 0.  Read by `BatchContext.init`
    /// and by the context-less convenience functions.
    devices: []const i32 = &.{},
    /// Put all pixel data of the batch into ONE pinned host allocation (zpx_host_alloc) instead of one
    /// `al` allocation per image: the device->host copy then runs at PCIe speed and in few large copies.
    /// The images of such a batch are views into that slab: release the whole batch with `Decoded.deinit`,
    /// never with `Image.free`.
    pinned: bool = false,
    /// What each image is: `.rgba` = `Image{ .RGBA }` holding the bytes `jpeg.load(..).rgbaPixels()` yields;
    /// `.native` = the variant `jpeg.load` itself returns (`.YCbCr` / `.Gray` planes with makeImg's strides,
    /// `.RGBA` for RGB-tagged files, `.CMYK`): 1.5 bytes per pixel instead of 4 for 4:2:0 files.
    output: enum { rgba, native } = .rgba,
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

/// The results of one batch.  `deinit` releases everything (the images too).
pub const Decoded = struct {
    results: []Result,
    slab: ?[]u8 = null, // pinned host memory all pixel slices point into (BatchOptions.pinned)

    pub fn deinit(self: *Decoded, al: std.mem.Allocator) void {
        if (self.slab) |s| {
            c.zpx_host_free(s.ptr); // the images are views: nothing else to free
        } else {
            for (self.results) |r| switch (r) {
                .ok => |img| img.free(al),
                .err => {},
            };
        }
        al.free(self.results);
        self.* = undefined;
    }
};

/// Owns the zpx_ctx: device buffers, pinned staging, streams, the chunk pipeline's workers.  Create ONE per
/// calling thread and keep it for the life of the program: the first batch sizes the device buffers (about
/// 15 GB for 1024 x 1080p), later batches reuse them.  Not thread-safe.
pub const BatchContext = struct {
    handle: *c.zpx_ctx,

    pub fn init(devices: []const i32) DecodeError!BatchContext {
        var h: ?*c.zpx_ctx = null;
        const rc = c.zpx_ctx_create(if (devices.len == 0) null else devices.ptr, @intCast(devices.len), &h);
        if (rc != c.ZPX_OK) return toError(rc);
        return .{ .handle = h.? };
    }

    pub fn deinit(self: *BatchContext) void {
        c.zpx_ctx_destroy(self.handle);
        self.* = undefined;
    }

    /// Decode a batch of in-memory JPEGs with ONE library call (zpx_decode_batch_rgba / zpx_decode_batch_native:
    /// header parse, upload, kernels and download, pipelined in chunks so that the device->host link stays busy).
    /// One corrupt input does not fail the batch: it gets its own `.err`.
    pub fn decodeBatch(self: *BatchContext, al: std.mem.Allocator, buffers: []const []const u8, opts: BatchOptions) !Decoded {
        const n = buffers.len;
        const ptrs = try al.alloc([*c]const u8, n);
        defer al.free(ptrs);
        const lens = try al.alloc(usize, n);
        defer al.free(lens);
        const infos = try al.alloc(c.zpx_image_info, n);
        defer al.free(infos);
        const outs = try al.alloc([*c]u8, n);
        defer al.free(outs);
        const status = try al.alloc(i32, n);
        defer al.free(status);

        // sizes from the header-only probe (decodeConfig); the caller's side owns every output byte
        var total: usize = 0;
        for (buffers, 0..) |buf, i| {
            ptrs[i] = buf.ptr;
            lens[i] = buf.len;
            _ = c.zpx_probe(buf.ptr, buf.len, &infos[i]);
            if (infos[i].status == c.ZPX_OK) total += outLen(infos[i], opts) + 255 & ~@as(usize, 255);
        }
        var out = Decoded{ .results = try al.alloc(Result, n) };
        errdefer al.free(out.results);
        if (opts.pinned and total > 0) {
            const p: ?[*]u8 = @ptrCast(c.zpx_host_alloc(total));
            if (p == null) return error.OutOfMemory;
            out.slab = p.?[0..total];
        }
        errdefer if (out.slab) |s| c.zpx_host_free(s.ptr);

        var off: usize = 0;
        for (0..n) |i| {
            outs[i] = null;
            if (infos[i].status != c.ZPX_OK) {
                out.results[i] = .{ .err = toError(infos[i].status) };
                continue;
            }
            const len = outLen(infos[i], opts);
            const pixels: []u8 = if (out.slab) |s| s[off .. off + len] else try al.alloc(u8, len);
            off += len + 255 & ~@as(usize, 255);
            outs[i] = pixels.ptr;
            out.results[i] = .{ .ok = wrap(infos[i], pixels, opts) };
        }

        const rc = switch (opts.output) {
            .rgba => c.zpx_decode_batch_rgba(self.handle, ptrs.ptr, lens.ptr, @intCast(n), outs.ptr, null, status.ptr),
            .native => c.zpx_decode_batch_native(self.handle, ptrs.ptr, lens.ptr, @intCast(n), outs.ptr, status.ptr),
        };
        if (rc != c.ZPX_OK) {
            out.deinit(al);
            return toError(rc);
        }
        for (0..n) |i| {
            if (status[i] == c.ZPX_OK) continue;
            switch (out.results[i]) {
                .ok => |img| if (out.slab == null) img.free(al),
                .err => {},
            }
            out.results[i] = .{ .err = toError(status[i]) };
        }
        return out;
    }

    /// Same, reading the files first (beside `jpeg.load`, reference src/jpeg/root.zig:36).
    pub fn loadBatch(self: *BatchContext, al: std.mem.Allocator, paths: []const []const u8, opts: BatchOptions) !Decoded {
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
        return self.decodeBatch(al, bufs, opts);
    }

    /// GPU-resident hand-off: the staged calls, for consumers that take the RGBA where the kernels left it
    /// (zpx_batch_device_rgba) on their own stream.  See include/zpix_cuda.h for the ordering contract.
    pub fn raw(self: *BatchContext) *c.zpx_ctx {
        return self.handle;
    }
};

fn outLen(info: c.zpx_image_info, opts: BatchOptions) usize {
    return @intCast(if (opts.output == .rgba) info.rgba_len else info.native_len);
}

/// The Image a result slice is: Image{.RGBA} for rgbaPixels bytes, else the variant jpeg.load returns
/// (decoder.zig:361-370) with makeImg's strides (decoder.zig:1708-1783).
fn wrap(info: c.zpx_image_info, pixels: []u8, opts: BatchOptions) image.Image {
    const rect = image.Rectangle.init(0, 0, info.width, info.height);
    const w4: usize = @intCast(4 * info.width);
    if (opts.output == .rgba or info.variant == c.ZPX_VARIANT_RGBA)
        return .{ .RGBA = .{ .pixels = pixels, .stride = w4, .rect = rect } };
    return switch (info.variant) {
        c.ZPX_VARIANT_GRAY => .{ .Gray = .{ .pixels = pixels, .stride = @intCast(info.y_stride), .rect = rect } },
        c.ZPX_VARIANT_CMYK => .{ .CMYK = .{ .pixels = pixels, .stride = w4, .rect = rect } },
        else => .{ .YCbCr = .{
            .y = pixels[0..@intCast(info.native_cb_off)],
            .cb = pixels[@intCast(info.native_cb_off)..@intCast(info.native_cr_off)],
            .cr = pixels[@intCast(info.native_cr_off)..],
            .y_stride = @intCast(info.y_stride),
            .c_stride = @intCast(info.c_stride),
            .subsample_ratio = @enumFromInt(info.subsample_ratio), // same order as image.YCbCrSubsample
            .rect = rect,
            .pixels = pixels,
        } },
    };
}

/// Convenience for a single batch: creates a context, decodes, destroys it.  Costs a device-buffer allocation
/// per call -- programs that decode more than once keep a `BatchContext`.
pub fn decodeBatch(al: std.mem.Allocator, buffers: []const []const u8, opts: BatchOptions) !Decoded {
    var ctx = try BatchContext.init(opts.devices);
    defer ctx.deinit();
    return ctx.decodeBatch(al, buffers, opts);
}

pub fn loadBatch(al: std.mem.Allocator, paths: []const []const u8, opts: BatchOptions) !Decoded {
    var ctx = try BatchContext.init(opts.devices);
    defer ctx.deinit();
    return ctx.loadBatch(al, paths, opts);
}

/// `jpeg.loadFromBuffer` itself on the GPU path (reference src/jpeg/root.zig:10): one image, the native variant,
/// pixels owned by `al` exactly as the reference's.
pub fn loadFromBufferGpu(ctx: *BatchContext, al: std.mem.Allocator, buffer: []const u8) !image.Image {
    var d = try ctx.decodeBatch(al, &.{buffer}, .{ .output = .native });
    defer al.free(d.results);
    return switch (d.results[0]) {
        .ok => |img| img,
        .err => |e| e,
    };
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
