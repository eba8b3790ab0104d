//! Lines a maintainer adds to the reference's src/jpeg/root.zig (after line 7) to expose the batch
//! path under the existing module; nothing else in the module changes.
pub const batch = @import("batch.zig");
pub const decodeBatch = batch.decodeBatch;
pub const loadBatch = batch.loadBatch;
pub const BatchOptions = batch.BatchOptions;
