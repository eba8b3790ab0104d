//! Lines a maintainer adds to the reference's src/jpeg/root.zig (after line 7) to expose the batch
//! path under the existing module; nothing else in the module changes.
pub const batch = @import("batch.zig");
pub const BatchContext = batch.BatchContext;
pub const BatchOptions = batch.BatchOptions;
pub const Decoded = batch.Decoded;
pub const decodeBatch = batch.decodeBatch;
pub const loadBatch = batch.loadBatch;
pub const loadFromBufferGpu = batch.loadFromBufferGpu;
