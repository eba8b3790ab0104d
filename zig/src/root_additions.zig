//! Batch dispatcher: lines a maintainer adds to the reference's src/root.zig (beside fromFilePath / fromBuffer,
//! lines 24-40).  Same probing order as the reference (PNG, JPEG, QOI, BMP): JPEGs of a batch go to the GPU
//! path in ONE call, every other format to the CPU decoder the reference already has.
//! NOT COMPILED in this repository (no Zig toolchain, see DESIGN.md).
const std = @import("std");
const image = @import("image");
const jpeg = @import("jpeg");
const png = @import("png");
const qoi = @import("qoi");
const bmp = @import("bmp");

pub const BatchResult = union(enum) {
    ok: image.Image,
    err: anyerror,
};

/// `fromBuffer` for a whole batch.  results[i] is what `fromBuffer(allocator, buffers[i])` returns -- the native
/// Image variant, pixels owned by `allocator` -- with the JPEG ones decoded together on the GPU.
pub fn fromBuffers(allocator: std.mem.Allocator, ctx: *jpeg.BatchContext, buffers: []const []const u8) ![]BatchResult {
    const results = try allocator.alloc(BatchResult, buffers.len);
    errdefer allocator.free(results);
    var jpegs = std.ArrayList([]const u8).init(allocator);
    defer jpegs.deinit();
    var where = std.ArrayList(usize).init(allocator);
    defer where.deinit();
    for (buffers, 0..) |buf, i| {
        if (png.probeBuffer(buf)) {
            results[i] = if (png.loadFromBuffer(allocator, buf)) |img| .{ .ok = img } else |e| .{ .err = e };
        } else if (jpeg.probeBuffer(buf)) {
            try jpegs.append(buf);
            try where.append(i);
        } else if (qoi.probeBuffer(buf)) {
            results[i] = if (qoi.loadFromBuffer(allocator, buf)) |img| .{ .ok = img } else |e| .{ .err = e };
        } else if (bmp.probeBuffer(buf)) {
            results[i] = if (bmp.loadFromBuffer(allocator, buf)) |img| .{ .ok = img } else |e| .{ .err = e };
        } else {
            results[i] = .{ .err = error.UnknownImageFormat };
        }
    }
    if (jpegs.items.len > 0) {
        var d = try ctx.decodeBatch(allocator, jpegs.items, .{ .output = .native });
        defer allocator.free(d.results); // the images move into `results`
        for (d.results, where.items) |r, i| results[i] = switch (r) {
            .ok => |img| .{ .ok = img },
            .err => |e| .{ .err = e },
        };
    }
    return results;
}

/// `fromFilePath` for a batch of paths.
pub fn fromFilePaths(allocator: std.mem.Allocator, ctx: *jpeg.BatchContext, paths: []const []const u8) ![]BatchResult {
    const bufs = try allocator.alloc([]const u8, paths.len);
    var loaded: usize = 0;
    defer {
        for (bufs[0..loaded]) |b| allocator.free(b);
        allocator.free(bufs);
    }
    for (paths, 0..) |p, i| {
        bufs[i] = try std.fs.cwd().readFileAlloc(allocator, p, std.math.maxInt(usize));
        loaded = i + 1;
    }
    return fromBuffers(allocator, ctx, bufs);
}
