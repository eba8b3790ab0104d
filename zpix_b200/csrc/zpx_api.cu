// zpx_api.cu -- the extern "C" boundary (include/zpix_cuda.h): context, host scheduler,
// descriptor building, uploads, kernel launches, result hand-over.  No decode arithmetic lives
// here and there is no CPU decode path: without a CUDA device every decode entry point fails.
#include <cuda_runtime.h>
#include <string.h>

#include <sched.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "zpx_internal.h"
#include "zpx_kernels.h"

using namespace zpx;

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
namespace {

// CPUs this thread may run on (a rank bound to its GPU's NUMA node sees only those), and a per-thread cap for
// code that already runs several such loops side by side (the chunk pipeline's workers)
static size_t usable_cpus() {
    // ZPX_HOST_THREADS: a budget set by whoever knows how many processes share the box (bench.py under torchrun)
    if (const char* e = getenv("ZPX_HOST_THREADS")) {
        const long v = atol(e);
        if (v > 0) return (size_t)v;
    }
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return (size_t)n;
    }
    const size_t hw = std::thread::hardware_concurrency();
    return hw ? hw : 4;
}
static thread_local size_t t_par_limit = 0;  // 0 = no extra cap

template <typename F>
void parallel_for(size_t n, size_t min_chunk, F fn) {
    size_t hw = usable_cpus();
    if (hw > 64) hw = 64;
    if (t_par_limit && hw > t_par_limit) hw = t_par_limit;
    size_t nt = std::min(hw, (n + min_chunk - 1) / std::max<size_t>(min_chunk, 1));
    if (nt <= 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t i0 = next.fetch_add(min_chunk);
                if (i0 >= n) break;
                size_t i1 = std::min(n, i0 + min_chunk);
                for (size_t i = i0; i < i1; i++) fn(i);
            }
        });
    for (auto& t : th) t.join();
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();  // the failed attempt must not be what the next launch check reports
            e = cudaMalloc(&p, bytes);
            want = bytes;
            if (e != cudaSuccess) (void)cudaGetLastError();
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        else (void)cudaGetLastError();
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceCtx {
    int dev = 0;
    int sm_count = 148;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};
    DevBuf blob, ublob, coef, out, planes, desc, status, subs, scratch;
    DevBuf planes_late, late_list;  // native planes of fused images, made on demand (zpx_batch_fetch_native)
    DevBuf carry;                   // End-Of-Band runs handed from scan to scan (rescue_eob_carry)
    HostBuf stage, hdesc, hstatus, hflag;
};

// images of one format that take the fused kernel in one launch
struct FusedGroup {
    int h, v, nc;
    int tmax = 0;
    std::vector<ZpxTileDev> tiles;
    size_t tiles_off = 0;  // byte offset of the tile table inside the descriptor buffer
    uint64_t bytes = 0;    // algorithmic bytes: 128 B + 4 P over the group's images
};

struct DevicePlan {
    std::vector<int> images;  // batch indices scheduled here, in order
    std::vector<ZpxImageDev> imgs;
    std::vector<ZpxScanDev> scans;
    std::vector<ZpxIntervalDev> ivs;
    std::vector<ZpxHuffDev> huff;
    std::vector<ZpxQuantDev> quant;
    std::vector<FusedGroup> groups;
    std::vector<ZpxIntervalDev> ivs_prog;            // intervals of progressive scans (appended after the sequential ones)
    // progressive scans: work lists of interval indices (into the final interval array), launched in order.
    // List 3 * level + kind; kind 0: one warp per interval (zpx_k3.cu), 1: one lane per interval, grouped by pass type
    // and padded with ~0 (k3l_level), 2: DC refinement passes (k3l_dc_refine)
    std::vector<std::vector<uint32_t>> prog_lists;
    struct LaneGroup { uint32_t first = 0, count = 0, max_blocks = 0; };  // a pass type's group inside list 3 * level + 1
    std::vector<LaneGroup> prog_grp;                                      // [3 * level + pass type]
    std::vector<size_t> prog_off;                    // byte offsets of those lists in the descriptor buffer
    std::vector<std::pair<uint64_t, uint64_t>> prog_zero;  // (first block, blocks) of progressive images: zeroed before the scans
    size_t n_seq = 0;                                // sequential intervals = ivs[0, n_seq)
    size_t n_sub_iv = 0;                             // self-synchronising mode: ivs[0, n_sub_iv) take it, the rest
                                                     // of the sequential ones one lane per interval
    std::vector<uint8_t> iv_lane_only;               // per sequential interval (build time only)
    std::vector<ZpxWarpDev> warps;  // self-synchronising mode: one entry per warp
    std::vector<ZpxSegDev> segs;    // pieces of the sequential intervals for the unstuffing kernel
    size_t ublob_bytes = 0, off_segs = 0;
    size_t n_subs = 0;
    bool sub_mode = false;
    size_t off_warps = 0;
    std::vector<uint32_t> generic;  // device image indices on the unfused path
    int generic_max_blocks = 0;
    size_t generic_max_pixels = 0;
    // blob assembly: (batch image, src offset, length, dst offset)
    struct Copy { int img; size_t src, len, dst; };
    std::vector<Copy> copies;
    size_t blob_bytes = 0, coef_blocks = 0, out_bytes = 0, plane_bytes = 0;
    // descriptor buffer layout (byte offsets)
    size_t off_imgs = 0, off_scans = 0, off_ivs = 0, off_huff = 0, off_quant = 0, off_generic = 0, desc_bytes = 0;
    std::vector<uint64_t> out_off, plane_off0;  // per device image
    std::vector<uint32_t> inject_flags;         // test hook (zpx_batch_set_coefficients): per image, copied over img_flags
    zpx_timing timing;
    bool uploaded = false, decoded = false;
    // fused images whose planes nobody asked for at open time (ZPX_OPT_NATIVE_PLANES = 0): laid out in a buffer of their
    // own, reconstructed by the unfused IDCT kernel the first time zpx_batch_fetch_native wants them
    std::vector<uint32_t> late;
    size_t late_plane_bytes = 0;
    int late_max_blocks = 0;
    bool late_done = false;
    const int* conv_flag = nullptr;  // device flag: set if the sweeps enqueued on a caller's stream did not converge
    int native = 0;  // ZPX_OPT_NATIVE_PLANES at the time the plan was built
};

}  // namespace

struct zpx_ctx {
    std::vector<DeviceCtx> devs;
    int last_cuda = 0;
    std::string last_cuda_str;
    std::atomic<uint64_t> launches{0};
    int64_t opt_entropy_mode = 0, opt_force_generic = 0, opt_subseq = 0, opt_pipeline_chunk = 0;
    int64_t opt_pipeline_ramp = 1, opt_pipeline_workers = 3, opt_test_wide = 0, opt_native = 0, opt_progressive_mode = 0, opt_k2_dense = 0, opt_gated_sweeps = 2;
    bool busy = false;
    const zpx_batch* resident = nullptr;  // the batch whose data currently occupies the device buffers
    std::vector<zpx_ctx*> shadows;  // further sets of device buffers/streams for the chunk pipeline of zpx_decode_batch_rgba
};

struct zpx_batch {
    zpx_ctx* ctx = nullptr;
    int n = 0;
    std::vector<const uint8_t*> bufs;
    std::vector<size_t> lens;
    std::vector<ZpxParsed> parsed;
    std::vector<int> dev_of;      // device index per image (-1: not decoded)
    std::vector<int> slot_of;     // index inside the device's image list
    std::vector<int32_t> status;  // final status per image
    std::vector<DevicePlan> plans;
    bool status_ready = false;
};

namespace {

int cuda_fail(zpx_ctx* c, cudaError_t e) {
    c->last_cuda = (int)e;
    c->last_cuda_str = cudaGetErrorString(e);
    return ZPX_E_CUDA;
}
#define CU(ctx, call)                                \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__); \
    } while (0)

const char* const kErrNames[] = {
    "ok", "UnexpectedEof", "InvalidSOIMarker", "ShortSegmentLength", "UnknownMarker", "UnsupportedMarker",
    "MissingSosMarker", "MultipleSofMarkers", "NumberComponents", "Precision", "SofWrongLength",
    "RepeatedComponentIdentifier", "BadTqValue", "LumaChromaSubSamplingRatio", "DriWrongLength", "BadPqValue",
    "DqtWrongLength", "MissingFF00", "UnsupportedColorModel", "UninitializedHuffmanTable", "BadHuffmanCode",
    "DhtWrongLength", "BadTcValue", "BadThValue", "HuffZeroLength", "HuffTooLong", "SosWrongLength",
    "UnknownComponentSelector", "BadTdValue", "BadTaValue", "SamplingFactorsTooLarge", "BadSpectralSelection",
    "ProgressiveACCoefficientsForMoreThanOneComponent", "BadSuccessiveApproximation", "ExcessiveDCComponent",
    "UnexpectedHuffmanCode", "TooManyCoefficients", "BadRSTMarker", "CreateImageFailed", "UnsupportedComponent",
    "InvalidImageType", "ConfigOnly", "OutOfMemory",
};

// Components some scan covers.  Progressive frames allocate their coefficient arrays by looping the FRAME's
// component count over the scan's component list (entries past the scan's own count read as
// component 0), and only allocated components are reconstructed (decoder.zig:1272-1279, :1644).
// bits 4-7: the component is coded by some INTERLEAVED scan of a sequential frame, which visits every block
// of the MCU grid; a component that only has scans of its own is reconstructed where 8*bx < width and
// 8*by < height (decoder.zig:1331-1336) and keeps zeros in the rest of its padded plane
uint32_t recon_mask_of(const ZpxParsed& p) {
    uint32_t m = 0;
    for (const ZpxScanHost& sc : p.scans) {
        for (int i = 0; i < sc.ncomp; i++) {
            m |= 1u << sc.comp[i];
            if (sc.ncomp > 1) m |= 16u << sc.comp[i];
        }
        if (p.progressive && sc.ncomp < p.ncomp) m |= 1u;
    }
    return m;
}

// Frames the fused kernel takes: gray, or YCbCr with full-resolution luma sampling factors it has an instantiation
// for.  A frame made of ONE interleaved scan keeps its coefficients MCU by MCU (one bulk copy per tile); every other
// scan script (components in scans of their own, several scans, progressive) keeps one block grid per component and
// the kernel gathers a tile with one copy per block row.  Those frames are fused only for RGBA output: in their native
// planes the reference leaves blocks it never reconstructs at zero (decoder.zig:1331-1336, :1644-1651), which only
// the unfused kernels reproduce -- and only when every component is covered by some scan (same reason).
bool fused_eligible(const ZpxParsed& p, uint32_t recon_mask, int native) {
    if ((uint64_t)p.width * (uint64_t)p.height >= (1ull << 30)) return false;  // the fused kernel keeps 32-bit pixel offsets
    if (p.scans.empty()) return false;
    bool single = !p.progressive && p.scans.size() == 1 && p.scans[0].ncomp == p.ncomp;
    if (single)
        for (int i = 0; i < p.ncomp; i++) single = single && p.scans[0].comp[i] == i;
    if (!single && native != 0) return false;
    if (p.ncomp == 1) return p.mode == ZPX_MODE_GRAY && (single || (recon_mask & 1u));
    if (p.ncomp != 3 || p.mode != ZPX_MODE_YCBCR) return false;
    if (!single && (recon_mask & 7u) != 7u) return false;
    if (p.h[1] != 1 || p.v[1] != 1 || p.h[2] != 1 || p.v[2] != 1) return false;
    const int hv = p.h[0] << 4 | p.v[0];
    return hv == 0x11 || hv == 0x21 || hv == 0x22 || hv == 0x12 || hv == 0x41 || hv == 0x42;
}

struct TableDedup {
    std::map<std::string, int> huff_ix, quant_ix;
    int add_huff(const ZpxHuffHost& h, bool is_ac, std::vector<ZpxHuffDev>& out) {
        std::string key((const char*)h.counts, 16);
        key.append((const char*)h.vals, 256);
        key.push_back(h.defined ? 1 : 0);
        key.push_back(is_ac ? 1 : 0);
        auto it = huff_ix.find(key);
        if (it != huff_ix.end()) return it->second;
        ZpxHuffDev d;
        int malformed = 0;
        zpx_build_huff_dev(h, is_ac, &d, &malformed);
        out.push_back(d);
        int ix = (int)out.size() - 1;
        huff_ix.emplace(std::move(key), ix);
        return ix;
    }
    int add_quant(const int32_t* zz, std::vector<ZpxQuantDev>& out) {
        std::string key((const char*)zz, 64 * sizeof(int32_t));
        auto it = quant_ix.find(key);
        if (it != quant_ix.end()) return it->second;
        ZpxQuantDev d;
        for (int z = 0; z < 64; z++) d.q[zpx_unzig[z]] = zz[z];
        out.push_back(d);
        int ix = (int)out.size() - 1;
        quant_ix.emplace(std::move(key), ix);
        return ix;
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Can the lane-per-interval progressive kernels (zpx_k3l.cu) decode this frame?  They never read a coefficient back:
// an AC refinement pass learns which coefficients are non-zero (and their signs) from per-block bit maps and applies a
// correction bit as a blind add of +-2^Al.  That is exact when the scan script is an ordinary successive
// approximation: every band position of a component is coded by at most one first pass, before any refinement of it,
// and refined with strictly falling Al (the bit being added is clear, the magnitude stays below 2^15).  Anything else
// (only hand-made or damaged files) takes the warp-per-interval kernel, which works on the coefficients themselves.
bool lane_script_ok(const ZpxParsed& p) {
    int8_t gran[ZPX_MAX_COMP][64];  // lowest bit position coded so far; 127: untouched
    memset(gran, 127, sizeof(gran));
    for (const ZpxScanHost& s : p.scans) {
        if (s.al > 13) return false;
        if (s.ss == 0) continue;  // DC passes: a store or an OR
        const int c = s.comp[0];
        for (int z = s.ss; z <= s.se; z++) {
            if (s.ah == 0 ? gran[c][z] != 127 : (gran[c][z] != 127 && s.al >= gran[c][z])) return false;
            gran[c][z] = (int8_t)s.al;
        }
    }
    return true;
}

// Build everything one device needs for its share of the batch.
void build_plan(zpx_batch* b, int di) {
    DevicePlan& pl = b->plans[di];
    TableDedup dd;
    const bool force_generic = b->ctx->opt_force_generic != 0;
    pl.native = (int)b->ctx->opt_native;
    std::map<int, int> group_ix;  // (h<<8|v<<4|nc) -> index
    std::vector<std::vector<uint32_t>> lane_groups;  // progressive, lane per interval: [3 * level + pass type]
    for (size_t k = 0; k < pl.images.size(); k++) {
        const int bi = pl.images[k];
        const ZpxParsed& p = b->parsed[bi];
        ZpxImageDev im;
        memset(&im, 0, sizeof(im));
        im.width = p.width;
        im.height = p.height;
        im.mxx = p.mxx;
        im.myy = p.myy;
        im.ncomp = p.ncomp;
        im.mode = p.mode;
        im.progressive = p.progressive ? 1 : 0;
        im.hmax = p.h[0];
        im.vmax = p.v[0];
        im.status_slot = (uint32_t)k;
        // components some scan covers.  Progressive frames allocate their coefficient arrays by looping the FRAME's
        // component count over the scan's component list (entries past the scan's own count read as
        // component 0), and only allocated components are reconstructed (decoder.zig:1272-1279, :1644).
        im.recon_mask = recon_mask_of(p);
        int bpm = 0;
        for (int c = 0; c < p.ncomp; c++) {
            im.h[c] = (uint8_t)p.h[c];
            im.v[c] = (uint8_t)p.v[c];
            im.blk_off[c] = (uint32_t)bpm;
            bpm += p.h[c] * p.v[c];
            im.comp_bw[c] = p.mxx * p.h[c];
            im.comp_bh[c] = p.myy * p.v[c];
        }
        // one interleaved scan holding every component in frame order -> interleaved layout
        bool single = !p.progressive && p.scans.size() == 1 && p.scans[0].ncomp == p.ncomp;
        if (single)
            for (int i = 0; i < p.ncomp; i++) single = single && p.scans[0].comp[i] == i;
        im.layout = single ? ZPX_LAYOUT_INTERLEAVED : ZPX_LAYOUT_PLANAR;
        im.bpm = bpm;
        im.fused = (!force_generic && fused_eligible(p, im.recon_mask, pl.native)) ? 1 : 0;
        // coefficient storage
        const uint64_t nblocks = (uint64_t)p.mxx * p.myy * bpm;
        im.coef_base = pl.coef_blocks;
        uint64_t cb = pl.coef_blocks;
        for (int c = 0; c < p.ncomp; c++) {
            im.comp_base[c] = cb;
            cb += (uint64_t)im.comp_bw[c] * im.comp_bh[c];
        }
        pl.coef_blocks += nblocks;
        bool lanep = false;       // progressive frame on the lane-per-interval kernels (zpx_k3l.cu), decided below
        uint64_t map_blocks = 0, zero_pos_blocks = 0;  // its non-zero / sign maps and block positions, zeroed together
                                                       // with the coefficients
        // quantisers: sequential frames use the tables in force at the component's SOS,
        // progressive frames those at EOI
        for (int c = 0; c < p.ncomp; c++) {
            const int32_t* zz = p.final_quant[p.tq[c]];
            if (!p.progressive)
                for (const ZpxScanHost& s : p.scans)
                    for (int i = 0; i < s.ncomp; i++)
                        if (s.comp[i] == c) zz = s.quant[i];
            im.qidx[c] = dd.add_quant(zz, pl.quant);
            // the fused kernel multiplies with 8-bit quantisers (dp2a); 16-bit tables take the unfused path
            for (int z = 0; z < 64; z++)
                if (zz[z] > 255 || zz[z] < 0) im.fused = 0;
        }
        // output
        im.out_off = pl.out_bytes;
        pl.out_off.push_back(pl.out_bytes);
        pl.out_bytes += align_up((size_t)4 * p.width * p.height, 256);
        // native planes (generic path): exact makeImg layout: Y, Cb, Cr contiguous, then black
        {
            const bool late = im.fused && pl.native == 0;  // planes on demand, in dc.planes_late
            size_t& total = late ? pl.late_plane_bytes : pl.plane_bytes;
            pl.plane_off0.push_back(total);
            zpx_image_info info;
            zpx_fill_info(p, &info);
            size_t base = total;
            if (p.ncomp == 1) {
                im.plane_off[0] = base;
                im.plane_stride[0] = 8 * p.mxx;
                im.plane_rows[0] = 8 * p.myy;
                base += (size_t)im.plane_stride[0] * im.plane_rows[0];
            } else {
                const size_t w = (size_t)8 * p.h[0] * p.mxx, hh = (size_t)8 * p.v[0] * p.myy;
                const size_t cw = (size_t)info.c_stride, ch = (size_t)8 * p.v[1] * p.myy;
                im.plane_off[0] = base;
                im.plane_stride[0] = (int)w;
                im.plane_rows[0] = (int)hh;
                im.plane_off[1] = base + w * hh;
                im.plane_off[2] = base + w * hh + cw * ch;
                im.plane_stride[1] = im.plane_stride[2] = (int)cw;
                im.plane_rows[1] = im.plane_rows[2] = (int)ch;
                base += w * hh + 2 * cw * ch;
                if (p.ncomp == 4) {
                    im.plane_off[3] = base;
                    im.plane_stride[3] = 8 * p.h[3] * p.mxx;
                    im.plane_rows[3] = 8 * p.v[3] * p.myy;
                    base += (size_t)im.plane_stride[3] * im.plane_rows[3];
                }
            }
            total = align_up(base, 256);
            if (late) {
                pl.late.push_back((uint32_t)k);
                pl.late_max_blocks = std::max<int>(pl.late_max_blocks, (int)nblocks);
            }
        }
        if (!im.fused) {
            pl.generic.push_back((uint32_t)k);
            pl.generic_max_blocks = std::max<int>(pl.generic_max_blocks, (int)nblocks);
            pl.generic_max_pixels = std::max<size_t>(pl.generic_max_pixels, (size_t)p.width * p.height);
        } else {
            // tiles of the fused kernel: runs of MCUs inside one MCU row, sized so that one tile is
            // (close to) one block per thread of a 256-thread CTA
            const int nc = p.ncomp == 1 ? 1 : 3;
            const int key = p.h[0] << 8 | p.v[0] << 4 | nc;
            auto it = group_ix.find(key);
            if (it == group_ix.end()) {
                FusedGroup g;
                g.h = p.h[0];
                g.v = p.v[0];
                g.nc = nc;
                pl.groups.push_back(g);
                it = group_ix.emplace(key, (int)pl.groups.size() - 1).first;
            }
            FusedGroup& g = pl.groups[it->second];
            const int tcap = std::max(1, k2_fused_threads(p.h[0], p.v[0], nc) / bpm);
            const int per_row = (p.mxx + tcap - 1) / tcap;
            const int tn = (p.mxx + per_row - 1) / per_row;
            g.tmax = std::max(g.tmax, tn);
            if (nc == 1 && p.mxx * 2 <= tcap) {
                // narrow gray image: a tile is several whole MCU rows (contiguous in the coefficient stream)
                const int nr_max = std::min(16, tcap / p.mxx);
                g.tmax = std::max(g.tmax, nr_max * p.mxx);
                for (int my = 0; my < p.myy; my += nr_max) {
                    const int nr = std::min(nr_max, p.myy - my);
                    ZpxTileDev t;
                    t.img = (uint32_t)k;
                    t.my = (uint16_t)my;
                    t.n = (uint16_t)(nr * p.mxx);
                    t.mx0 = 0;
                    t.pad = (uint32_t)nr | (uint32_t)p.mxx << 16 | (im.qidx[0] < 255 ? (uint32_t)(im.qidx[0] + 1) << 8 : 0u);
                    g.tiles.push_back(t);
                }
            } else
            for (int my = 0; my < p.myy; my++)
                for (int mx0 = 0; mx0 < p.mxx; mx0 += tn) {
                    ZpxTileDev t;
                    t.img = (uint32_t)k;
                    t.my = (uint16_t)my;
                    t.n = (uint16_t)std::min(tn, p.mxx - mx0);
                    t.mx0 = (uint32_t)mx0;
                    // colour tiles: the components' quantiser indices, so that the kernel's loads of the tables do not
                    // wait for the image descriptor first (gray tiles: rows | width << 16, above; gray has one table)
                    t.pad = nc == 3 && im.qidx[0] < 1023 && im.qidx[1] < 1023 && im.qidx[2] < 1023
                                ? (uint32_t)(im.qidx[0] + 1) | (uint32_t)(im.qidx[1] + 1) << 10 | (uint32_t)(im.qidx[2] + 1) << 20 | 1u << 31
                                : nc == 1 && im.qidx[0] < 255 ? (uint32_t)(im.qidx[0] + 1) << 8 : 0u;
                    g.tiles.push_back(t);
                }
            g.bytes += nblocks * 128 + (uint64_t)4 * p.width * p.height;
        }
        // scans and intervals
        // Progressive scans are launched by dependency level, not by ordinal: a scan has to wait only for
        // earlier scans that touch the same coefficients (a common component and overlapping spectral
        // bands), e.g. libjpeg's 10-scan script needs 3 launches instead of 10.
        std::vector<int> level(p.scans.size(), 0);
        if (p.progressive) {
            for (size_t a = 0; a < p.scans.size(); a++) {
                for (size_t t = 0; t < a; t++) {
                    const ZpxScanHost &x = p.scans[a], &y = p.scans[t];
                    bool share = false;
                    for (int i = 0; i < x.ncomp; i++)
                        for (int j = 0; j < y.ncomp; j++) share = share || x.comp[i] == y.comp[j];
                    if (share && x.ss <= y.se && y.ss <= x.se) level[a] = std::max(level[a], level[t] + 1);
                }
            }
        }
        // Progressive frames with an ordinary scan script take the lane-per-interval kernels.  Their scratch lives
        // right after the coefficients: 16 bytes of non-zero / sign maps per block; per AC scan one 32-bit stream
        // position per coded block (+ one progress counter per interval); per coded block of the AC refinement scans
        // of one level a 64-byte list of zero positions (levels reuse the space).
        std::vector<uint32_t> pos_off(p.scans.size(), 0), zl_off(p.scans.size(), 0);
        auto coded_w = [&](int c) { return std::min(im.comp_bw[c], (p.width + 7) / 8); };
        auto coded_h = [&](int c) { return std::min(im.comp_bh[c], (p.height + 7) / 8); };
        if (p.progressive && b->ctx->opt_progressive_mode == 0 && lane_script_ok(p)) {
            uint64_t pos_entries = 0, zl_max = 0;
            std::vector<uint64_t> zl_level;
            bool fits = true;
            for (size_t a = 0; a < p.scans.size(); a++) {
                const ZpxScanHost& s = p.scans[a];
                for (const ZpxIntervalHost& iv : s.intervals) fits = fits && iv.limit - iv.start < (1u << 27);  // bit positions in 31 bits
                if (s.ss == 0) continue;
                const uint64_t coded = (uint64_t)coded_w(s.comp[0]) * coded_h(s.comp[0]);
                pos_off[a] = (uint32_t)pos_entries;
                pos_entries += coded + s.intervals.size();
                if (s.ah == 0) continue;
                if (zl_level.size() <= (size_t)level[a]) zl_level.resize(level[a] + 1, 0);
                zl_off[a] = (uint32_t)zl_level[level[a]];
                zl_level[level[a]] += coded;
                zl_max = std::max(zl_max, zl_level[level[a]]);
            }
            const uint64_t pos_blocks = (pos_entries * 4 + 127) / 128, zl_blocks = (zl_max * 64 + 127) / 128;
            // (zpx_batch_open budgets 128 bytes of scratch per block; scripts that refine many bands side by side
            // go to the warp-per-interval kernel)
            if (fits && pos_blocks + zl_blocks <= nblocks + 16 && pos_entries < (1ull << 31)) {
                lanep = true;
                zero_pos_blocks = pos_blocks;  // (a first pass leaves the entries of blocks without bits at zero)
                map_blocks = (nblocks + 7) / 8;
                im.pmask_base = pl.coef_blocks;
                im.ppos_base = im.pmask_base + map_blocks;
                im.pzl_base = im.ppos_base + pos_blocks;
                pl.coef_blocks += map_blocks + pos_blocks + zl_blocks;
            }
        }
        int scan_index = 0;
        for (const ZpxScanHost& s : p.scans) {
            ZpxScanDev sd;
            memset(&sd, 0, sizeof(sd));
            sd.img = (uint32_t)k;
            sd.ncomp = s.ncomp;
            sd.interleaved = s.ncomp > 1 ? 1 : 0;
            sd.ss = s.ss;
            sd.se = s.se;
            sd.ah = s.ah;
            sd.al = s.al;
            sd.total_mcu = p.mxx * p.myy;
            sd.restart_interval = s.restart_interval;
            sd.pos_off = pos_off[scan_index];
            sd.zl_off = zl_off[scan_index];
            sd.scan_index = scan_index++;
            // A sequential frame may code a component in more than one scan (only broken files do): the reference
            // reconstructs every scan as it comes, so the last one wins.  Here all scans of an image decode side by
            // side: the earlier ones are still decoded (their errors count) but do not store their blocks.
            bool superseded[ZPX_MAX_COMP] = {false, false, false, false};
            if (!p.progressive)
                for (int i = 0; i < s.ncomp; i++)
                    for (const ZpxScanHost* t = &s + 1; t <= &p.scans.back(); t++)
                        for (int j = 0; j < t->ncomp; j++) superseded[i] = superseded[i] || t->comp[j] == s.comp[i];
            int nb = 0;
            for (int i = 0; i < s.ncomp; i++) {
                const int c = s.comp[i];
                const int dci = dd.add_huff(s.dc[i], false, pl.huff), aci = dd.add_huff(s.ac[i], true, pl.huff);
                for (int j = 0; j < p.h[c] * p.v[c]; j++) {
                    sd.blk_comp[nb] = (uint8_t)c;
                    sd.blk_hx[nb] = (uint8_t)(j % p.h[c]);
                    sd.blk_vy[nb] = (uint8_t)(j / p.h[c]);
                    sd.blk_slot[nb] = (uint8_t)(im.blk_off[c] + j);
                    sd.blk_dc[nb] = (uint16_t)dci;
                    sd.blk_ac[nb] = (uint16_t)aci;
                    sd.blk_pack[nb][0] = (uint32_t)dci;
                    sd.blk_pack[nb][1] = (uint32_t)aci;
                    sd.blk_pack[nb][2] = (uint32_t)c | (uint32_t)(j % p.h[c]) << 8 | (uint32_t)(j / p.h[c]) << 16 |
                                         (uint32_t)(im.blk_off[c] + j) << 24;
                    sd.blk_pack[nb][3] = (uint32_t)p.h[c] | (uint32_t)p.v[c] << 8 | (s.dc[i].defined ? 0u : 1u << 16) |
                                         (s.ac[i].defined ? 0u : 1u << 17) | (superseded[i] ? 1u << 18 : 0u);
                    nb++;
                }
            }
            sd.nblk = nb;
            sd.rotate = 0;
            if (s.ncomp > 1 && nb == s.ncomp) {
                bool same = true;
                for (int i = 1; i < nb; i++) same = same && sd.blk_dc[i] == sd.blk_dc[0] && sd.blk_ac[i] == sd.blk_ac[0];
                sd.rotate = same ? 1 : 0;
            }
            if (s.ncomp == 1) {
                const int c = s.comp[0];
                sd.cw = coded_w(c);
                sd.ch = coded_h(c);
            }
            // An AC table that holds a symbol (r, 0) with 0 < r < 15 can start an End-Of-Band run inside a
            // sequential scan; the self-synchronising decoder does not model that state, so such
            // scans always take one lane per interval.
            bool eob_capable = false;
            if (!p.progressive)
                for (int i = 0; i < s.ncomp; i++)
                    for (int v = 0; v < s.ac[i].num_codes; v++) {
                        const int sym = s.ac[i].vals[v];
                        eob_capable = eob_capable || ((sym & 15) == 0 && (sym >> 4) >= 1 && (sym >> 4) <= 14);
                    }
            const uint32_t scan_ix = (uint32_t)pl.scans.size();
            pl.scans.push_back(sd);
            if (s.intervals.empty()) continue;
            // one contiguous copy per scan: [first interval start, last interval limit)
            const size_t src0 = s.intervals.front().start, src1 = s.intervals.back().limit;
            const size_t dst0 = pl.blob_bytes;
            pl.copies.push_back({bi, src0, src1 - src0, dst0});
            pl.blob_bytes = align_up(dst0 + (src1 - src0) + 8, 16);
            uint32_t ord = 0;
            for (const ZpxIntervalHost& iv : s.intervals) {
                ZpxIntervalDev d;
                d.start = dst0 + (iv.start - src0);
                d.len = (uint32_t)(iv.limit - iv.start);
                d.scan = scan_ix;
                d.first_mcu = iv.first_mcu;
                d.n_mcu = iv.n_mcu;
                d.ordinal = ord++;
                d.flags = iv.eof_limit ? 1u : 0u;
                // bit 1: last interval of a scan that another scan follows (End-Of-Band run carry check)
                if (&iv == &s.intervals.back() && &s != &p.scans.back()) d.flags |= 2u;
                if (s.ncomp > 1) {
                    d.first_block = iv.first_mcu * (uint32_t)nb;
                    d.n_blocks = iv.n_mcu * (uint32_t)nb;
                } else {
                    // non-interleaved: h*v linearised blocks per MCU iteration over the padded grid, only
                    // those intersecting the image carry data (decoder.zig:1331-1336)
                    const int c = s.comp[0];
                    const uint64_t hv = (uint64_t)p.h[c] * p.v[c], bw = (uint64_t)im.comp_bw[c];
                    auto coded_before = [&](uint64_t x) {
                        const uint64_t by = x / bw, rem = x % bw;
                        return std::min<uint64_t>(by, (uint64_t)sd.ch) * sd.cw + (by < (uint64_t)sd.ch ? std::min<uint64_t>(rem, (uint64_t)sd.cw) : 0);
                    };
                    const uint64_t x0 = iv.first_mcu * hv, x1 = ((uint64_t)iv.first_mcu + iv.n_mcu) * hv;
                    d.first_block = (uint32_t)coded_before(x0);
                    d.n_blocks = (uint32_t)(coded_before(x1) - coded_before(x0));
                }
                d.sub_first = d.nsub = d.sub_bytes = 0;
                d.ulen = 0;
                d.ustart = 0;
                if (!p.progressive || lanep) {
                    // the interval's place in the unstuffed blob and the pieces k0_unstuff works on
                    d.ulen = d.len - iv.n_stuffed;
                    d.ustart = pl.ublob_bytes;
                    pl.ublob_bytes += align_up(d.ulen, 16);
                    for (uint32_t q = 0; q < iv.n_segs; q++) {
                        const ZpxSegHost& sh = s.segs[iv.seg_first + q];
                        ZpxSegDev sgd;
                        sgd.src = dst0 + (sh.src - src0);
                        sgd.dst = d.ustart + sh.uoff;
                        sgd.len = sh.len;
                        sgd.flags = q + 1 == iv.n_segs ? 1u : 0u;
                        pl.segs.push_back(sgd);
                    }
                    if (!p.progressive) pl.iv_lane_only.push_back(eob_capable ? 1 : 0);
                }
                if (p.progressive) {
                    const size_t lv = (size_t)level[sd.scan_index];
                    if (pl.prog_lists.size() < 3 * (lv + 1)) pl.prog_lists.resize(3 * (lv + 1));
                    if (lane_groups.size() < 3 * (lv + 1)) lane_groups.resize(3 * (lv + 1));
                    const uint32_t self = (uint32_t)pl.ivs_prog.size();  // (indices are fixed up below)
                    if (!lanep) pl.prog_lists[3 * lv].push_back(self);
                    else if (s.ss == 0 && s.ah != 0) pl.prog_lists[3 * lv + 2].push_back(self);
                    else lane_groups[3 * lv + (s.ss == 0 ? 0 : s.ah == 0 ? 1 : 2)].push_back(self);
                    pl.ivs_prog.push_back(d);
                } else {
                    pl.ivs.push_back(d);
                }
            }
        }
        if (p.progressive) pl.prog_zero.push_back({im.coef_base, nblocks + map_blocks + zero_pos_blocks});
        pl.imgs.push_back(im);
    }
    // progressive intervals go after the sequential ones
    pl.n_seq = pl.ivs.size();
    // longest scans first: a launch is as long as its slowest warp, so the big ones must not start last
    auto by_len = [&](uint32_t a, uint32_t b) { return pl.ivs_prog[a].len > pl.ivs_prog[b].len; };
    for (auto& l : pl.prog_lists) std::stable_sort(l.begin(), l.end(), by_len);
    // lane-per-interval lists: the level's DC-first, AC-first and AC-refinement intervals, each group sorted the same
    // way (the 32 lanes of a warp then finish together) and padded to whole warps
    for (size_t g = 0; g < lane_groups.size(); g++) {
        std::vector<uint32_t>& l = lane_groups[g];
        if (l.empty()) continue;
        std::stable_sort(l.begin(), l.end(), by_len);
        if (g % 3 == 1) {
            // AC first passes: 16 streams per warp (k3l_level keeps a 10-bit table per lane for them)
            std::vector<uint32_t> w;
            for (size_t i = 0; i < l.size(); i += 16) {
                w.insert(w.end(), l.begin() + (ptrdiff_t)i, l.begin() + (ptrdiff_t)std::min(i + 16, l.size()));
                w.resize(align_up(w.size(), 32), 0xffffffffu);
            }
            l.swap(w);
        }
        l.resize(align_up(l.size(), 32), 0xffffffffu);
        std::vector<uint32_t>& dst = pl.prog_lists[g / 3 * 3 + 1];
        if (pl.prog_grp.size() <= g) pl.prog_grp.resize(g + 1);
        DevicePlan::LaneGroup& a = pl.prog_grp[g];
        a.first = (uint32_t)dst.size();
        a.count = (uint32_t)l.size();
        for (uint32_t ix : l)
            if (ix != 0xffffffffu) a.max_blocks = std::max(a.max_blocks, pl.ivs_prog[ix].n_blocks);
        dst.insert(dst.end(), l.begin(), l.end());
    }
    pl.prog_grp.resize(pl.prog_lists.size());
    for (auto& l : pl.prog_lists)
        for (uint32_t& ix : l)
            if (ix != 0xffffffffu) ix += (uint32_t)pl.n_seq;
    pl.ivs.insert(pl.ivs.end(), pl.ivs_prog.begin(), pl.ivs_prog.end());
    // entropy mode: with enough restart intervals to fill the GPU, one lane per interval decodes each
    // once, serially; otherwise the self-synchronising decoder parallelises inside the intervals
    const int64_t mode = b->ctx->opt_entropy_mode;
    {
        // auto: estimated time of either decoder (measured on B200, cfg2-like data).  One lane per interval is bound
        // by its longest interval while there are few of them (a lane decodes about 2.1 MB/s) and by instruction
        // issue once the GPU is full (125 GB/s); the self-synchronising decoder parallelises inside intervals but
        // decodes everything twice (45 GB/s) and has a fixed cost of a few launches.
        uint64_t total = 0, longest = 0;
        for (size_t k = 0; k < pl.n_seq; k++) {
            total += pl.ivs[k].len;
            longest = std::max<uint64_t>(longest, pl.ivs[k].len);
        }
        const double lane_s = std::max((double)longest / 2.1e6, (double)total / 125e9);
        const double sub_s = 0.3e-3 + (double)total / (total < (100u << 20) ? 38e9 : 45e9);
        pl.sub_mode = mode == 2 || (mode == 0 && sub_s < lane_s);
    }
    pl.n_sub_iv = 0;
    if (pl.sub_mode) {
        // lane-only intervals (End-Of-Band-run capable tables) to the back of the sequential range
        std::vector<ZpxIntervalDev> keep, lane;
        for (size_t k = 0; k < pl.n_seq; k++) (pl.iv_lane_only[k] ? lane : keep).push_back(pl.ivs[k]);
        pl.n_sub_iv = keep.size();
        std::copy(keep.begin(), keep.end(), pl.ivs.begin());
        std::copy(lane.begin(), lane.end(), pl.ivs.begin() + (ptrdiff_t)keep.size());
        for (size_t k = 0; k < pl.n_sub_iv; k++) {
            ZpxIntervalDev& d = pl.ivs[k];
            const uint32_t span = d.ulen;
            // sub-sequence size: long streams re-synchronise more slowly (more blocks per MCU, longer runs past the
            // boundary) and have lanes to spare: measured on B200, 384 bytes for 512x512 files, 768 for 2160p ones
            const uint32_t submax = b->ctx->opt_subseq > 0 ? (uint32_t)align_up((size_t)b->ctx->opt_subseq, 4)
                                    : span < (256u << 10) ? 384u : span < (1u << 20) ? 512u : 768u;
            uint32_t sub = (uint32_t)align_up((span + 31) / 32, 4);
            sub = std::max(32u, std::min(submax, sub));
            d.sub_bytes = sub;
            d.nsub = std::max(1u, (span + sub - 1) / sub);
            d.sub_first = (uint32_t)pl.n_subs;
            pl.n_subs += d.nsub;
            for (uint32_t f = 0; f < d.nsub; f += 32) pl.warps.push_back({(uint32_t)k, f});
        }
    }
    // descriptor buffer layout
    size_t off = 0;
    auto place = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    pl.off_imgs = place(pl.imgs.size() * sizeof(ZpxImageDev));
    pl.off_scans = place(pl.scans.size() * sizeof(ZpxScanDev));
    pl.off_ivs = place(pl.ivs.size() * sizeof(ZpxIntervalDev));
    pl.off_huff = place(pl.huff.size() * sizeof(ZpxHuffDev));
    pl.off_quant = place(pl.quant.size() * sizeof(ZpxQuantDev));
    pl.off_generic = place(pl.generic.size() * sizeof(uint32_t));
    pl.off_warps = place(pl.warps.size() * sizeof(ZpxWarpDev));
    pl.off_segs = place(pl.segs.size() * sizeof(ZpxSegDev));
    pl.prog_off.clear();
    for (auto& l : pl.prog_lists) pl.prog_off.push_back(place(l.size() * sizeof(uint32_t)));
    for (FusedGroup& g : pl.groups) g.tiles_off = place(g.tiles.size() * sizeof(ZpxTileDev));
    pl.desc_bytes = off;
    memset(&pl.timing, 0, sizeof(pl.timing));
    pl.timing.images = (int32_t)pl.images.size();
    for (const ZpxImageDev& im : pl.imgs) {
        pl.timing.pixels += (uint64_t)im.width * im.height;
        pl.timing.rgba_bytes += (uint64_t)4 * im.width * im.height;
    }
    pl.timing.coef_bytes = pl.coef_blocks * 128;
    for (const ZpxIntervalDev& d : pl.ivs) pl.timing.entropy_bytes_in += d.len;
    for (const FusedGroup& g : pl.groups) pl.timing.idct_fused_bytes += g.bytes;
}

int decode_on_device(zpx_batch* b, int di, cudaStream_t user_stream, bool host_sweeps = false) {
    zpx_ctx* ctx = b->ctx;
    DeviceCtx& dc = ctx->devs[di];
    DevicePlan& pl = b->plans[di];
    if (pl.images.empty()) return ZPX_OK;
    CU(ctx, cudaSetDevice(dc.dev));
    cudaStream_t st = user_stream ? user_stream : dc.stream;
    uint8_t* desc = (uint8_t*)dc.desc.p;

    CU(ctx, cudaEventRecord(dc.ev[0], st));
    CU(ctx, cudaMemsetAsync(dc.status.p, 0xff, pl.imgs.size() * sizeof(unsigned long long), st));
    uint32_t* img_flags = (uint32_t*)((uint8_t*)dc.status.p + align_up(pl.imgs.size() * sizeof(unsigned long long), 256));
    CU(ctx, cudaMemsetAsync(img_flags, 0, pl.imgs.size() * sizeof(uint32_t), st));
    if (!pl.inject_flags.empty())  // coefficients came from zpx_batch_set_coefficients, not from the entropy kernels
        CU(ctx, cudaMemcpyAsync(img_flags, pl.inject_flags.data(), pl.imgs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    // (the unfused kernels leave the planes of components no scan covers at makeImg's zeros; the fused kernel
    // writes every byte of its images' planes)
    if (pl.plane_bytes && !pl.generic.empty()) CU(ctx, cudaMemsetAsync(dc.planes.p, 0, pl.plane_bytes, st));

    int k1_launches = 0, k2_launches = 0;
    // ---- K1: entropy decode ----
    K1Params k1;
    k1.blob = (const uint8_t*)dc.blob.p;
    k1.ublob = (const uint8_t*)dc.ublob.p;
    k1.dbg = getenv("ZPX_K1_DBG") ? atoi(getenv("ZPX_K1_DBG")) : 0;
    k1.ivs = (const ZpxIntervalDev*)(desc + pl.off_ivs);
    k1.n_iv = (int)pl.n_seq;
    k1.scans = (const ZpxScanDev*)(desc + pl.off_scans);
    k1.imgs = (const ZpxImageDev*)(desc + pl.off_imgs);
    k1.huff = (const ZpxHuffDev*)(desc + pl.off_huff);
    k1.coef = (uint4*)dc.coef.p;
    k1.status = (unsigned long long*)dc.status.p;
    k1.img_flags = img_flags;
    k1.eob_in = nullptr;
    k1.eob_out = nullptr;
    if (!pl.segs.empty()) {
        CU(ctx, k0_launch_unstuff(k1.blob, (uint8_t*)dc.ublob.p, (const ZpxSegDev*)(desc + pl.off_segs), (int)pl.segs.size(), st));
        k1_launches++;
    }
    if (k1.n_iv > 0 && !pl.sub_mode) {
        CU(ctx, k1_launch_lane_per_interval(k1, dc.sm_count, st));
        k1_launches++;
    } else if (k1.n_iv > 0) {
        if (pl.n_sub_iv < pl.n_seq) {  // scans the self-synchronising decoder does not take
            K1Params kl = k1;
            kl.ivs = k1.ivs + pl.n_sub_iv;
            kl.n_iv = (int)(pl.n_seq - pl.n_sub_iv);
            CU(ctx, k1_launch_lane_per_interval(kl, dc.sm_count, st));
            k1_launches++;
        }
        k1.n_iv = (int)pl.n_sub_iv;
    }
    if (k1.n_iv > 0 && pl.sub_mode) {
        K1SParams ks;
        ks.k1 = k1;
        ks.warps = (const ZpxWarpDev*)(desc + pl.off_warps);
        ks.n_warps = (int)pl.warps.size();
        ks.n_iv = k1.n_iv;
        uint8_t* sb = (uint8_t*)dc.subs.p;
        const size_t S = align_up(pl.n_subs, 64);
        ks.s_dc = (int4*)sb;
        ks.s_in = (unsigned long long*)(sb + S * 16);
        ks.s_out = (unsigned long long*)(sb + S * 24);
        ks.s_n = (int*)(sb + S * 32);
        ks.s_bad = (int*)(sb + S * 36);
        ks.changed = (int*)(sb + S * 40);  // (256 bytes: up to 64 flags)
        ks.gate = nullptr;
        int* hflag = (int*)dc.hflag.p;
        pl.conv_flag = nullptr;
        // sweep 0 decodes every sub-sequence from its guess and settles each warp; later sweeps carry
        // end states across warp boundaries until nothing changes (exact fixed point)
        CU(ctx, k1s_launch_sync(ks, 0, st));
        k1_launches++;
        bool multi = false;
        for (size_t k = 0; k < pl.n_sub_iv; k++) multi = multi || pl.ivs[k].nsub > 32;
        if (multi && user_stream != nullptr && !host_sweeps) {
            // On the caller's stream nothing is read back: the boundary pass and ZPX_OPT_GATED_SWEEPS further rounds of
            // (sweep, boundary pass) are enqueued now, each launch gated on the flag the pass before it wrote -- a
            // round with nothing to do costs two empty launches.  Streams that need more rounds than that (tiny
            // sub-sequences, adversarial data) leave the last flag set: the status call sees it and decodes the
            // device's share again with the host loop below (finalize_status).
            const int ZPX_GATED_SWEEPS = std::max(0, (int)ctx->opt_gated_sweeps);  // (ZPX_OPT_GATED_SWEEPS, default 2)
            int* F = ks.changed;
            CU(ctx, cudaMemsetAsync(F, 0, (2 * ZPX_GATED_SWEEPS + 1) * sizeof(int), st));
            ks.changed = F;
            CU(ctx, k1s_launch_fix(ks, st));
            k1_launches++;
            for (int r = 1; r <= ZPX_GATED_SWEEPS; r++) {
                ks.gate = F + 2 * r - 2;
                ks.changed = F + 2 * r - 1;
                CU(ctx, k1s_launch_sync(ks, r, st));
                ks.gate = F + 2 * r - 1;
                ks.changed = F + 2 * r;
                CU(ctx, k1s_launch_fix(ks, st));
                k1_launches += 2;
            }
            ks.gate = nullptr;
            pl.conv_flag = F + 2 * ZPX_GATED_SWEEPS;
            if (ctx->opt_gated_sweeps < 0) CU(ctx, cudaMemsetAsync(F, 1, 1, st));  // test hook: exercise the repair path
            multi = false;
        }
        for (int sweep = 1; multi; sweep++) {
            // warp boundaries first (cheap); a full sweep only if one of them moved
            CU(ctx, cudaMemsetAsync(ks.changed, 0, sizeof(int), st));
            CU(ctx, k1s_launch_fix(ks, st));
            k1_launches++;
            CU(ctx, cudaMemcpyAsync(hflag, ks.changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
            if (*hflag == 0) break;
            CU(ctx, cudaMemsetAsync(ks.changed, 0, sizeof(int), st));
            CU(ctx, k1s_launch_sync(ks, sweep, st));
            k1_launches++;
            CU(ctx, cudaMemcpyAsync(hflag, ks.changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
            if (*hflag == 0) break;
        }
        CU(ctx, k1s_launch_scan(ks, st));
        CU(ctx, k1s_launch_write(ks, st));
        k1_launches += 2;
    }
    // ---- K3: progressive frames, launches by dependency level over zeroed coefficient grids (and bit maps) ----
    for (const auto& z : pl.prog_zero)
        CU(ctx, cudaMemsetAsync((uint8_t*)dc.coef.p + z.first * 128, 0, z.second * 128, st));
    for (size_t k = 0; k < pl.prog_lists.size(); k++) {
        if (pl.prog_lists[k].empty()) continue;
        const uint32_t* list = (const uint32_t*)(desc + pl.prog_off[k]);
        const int nl = (int)pl.prog_lists[k].size();
        if (k % 3 == 0) {
            CU(ctx, k3_launch_progressive(k1, list, nl, st));
        } else if (k % 3 == 1) {
            // the serial lanes only find where the blocks of an AC pass start; the writes follow in parallel.  AC
            // refinement passes: zero-position lists first
            const DevicePlan::LaneGroup &f = pl.prog_grp[k / 3 * 3 + 1], &a = pl.prog_grp[k / 3 * 3 + 2];
            if (a.count) {
                CU(ctx, k3l_launch_refine_prep(k1, list + a.first, (int)a.count, a.max_blocks, st));
                k1_launches++;
            }
            CU(ctx, k3l_launch_level(k1, list, nl, st));
            if (f.count) {
                CU(ctx, k3l_launch_first_apply(k1, list + f.first, (int)f.count, f.max_blocks, st));
                k1_launches++;
            }
            if (a.count) {
                CU(ctx, k3l_launch_refine_apply(k1, list + a.first, (int)a.count, a.max_blocks, st));
                k1_launches++;
            }
        } else {
            uint32_t mb = 0;
            for (uint32_t ix : pl.prog_lists[k]) mb = std::max(mb, pl.ivs[ix].n_blocks);
            CU(ctx, k3l_launch_dc_refine(k1, list, nl, mb, st));
        }
        k1_launches++;
    }
    CU(ctx, cudaEventRecord(dc.ev[1], st));

    // ---- K2: fused kernel, one launch per sampling format ----
    for (const FusedGroup& g : pl.groups) {
        K2Params k2;
        k2.coef = (const int16_t*)dc.coef.p;
        k2.out = pl.native == 2 ? nullptr : (uint8_t*)dc.out.p;
        k2.planes = pl.native != 0 ? (uint8_t*)dc.planes.p : nullptr;
        k2.imgs = k1.imgs;
        k2.tiles = (const ZpxTileDev*)(desc + g.tiles_off);
        k2.quant = (const ZpxQuantDev*)(desc + pl.off_quant);
        k2.img_flags = img_flags;
        k2.ntiles = (int)g.tiles.size();
        k2.tmax = g.tmax;
        k2.dense_only = ctx->opt_k2_dense != 0;
        // at least four warps: phase 2 hands whole item rows to warps, eight row(-pair)s per MCU row (measured: 4:4:4
        // tiles of 32 MCUs, 96 blocks, +3.6 % with a fourth warp that idles in phase 1); ZPX_K2_MIN_NT: experiments
        static const int min_nt = getenv("ZPX_K2_MIN_NT") ? atoi(getenv("ZPX_K2_MIN_NT")) : 128;
        k2.nt = std::max(std::max(96, min_nt), (g.tmax * k2_fused_bpm(g.h, g.v, g.nc) + 31) / 32 * 32);
        CU(ctx, k2_launch_fused(g.h, g.v, g.nc, k2, dc.sm_count, st));
        k2_launches++;
    }
    CU(ctx, cudaEventRecord(dc.ev[2], st));
    // ---- K2 generic ----
    if (!pl.generic.empty()) {
        K2GParams kg;
        kg.coef = (const int16_t*)dc.coef.p;
        kg.planes = (uint8_t*)dc.planes.p;
        kg.out = (uint8_t*)dc.out.p;
        kg.imgs = k1.imgs;
        kg.quant = (const ZpxQuantDev*)(desc + pl.off_quant);
        const uint32_t* list = (const uint32_t*)(desc + pl.off_generic);
        const int total = (int)pl.generic.size();
        for (int i0 = 0; i0 < total; i0 += 32768) {
            kg.list = list + i0;
            CU(ctx, k2g_launch(kg, std::min(32768, total - i0), pl.generic_max_blocks, pl.generic_max_pixels, st));
            k2_launches += 2;
        }
    }
    CU(ctx, cudaEventRecord(dc.ev[3], st));
    ctx->launches += (uint64_t)(k1_launches + k2_launches);
    pl.timing.entropy_launches = k1_launches;
    pl.timing.idct_launches = k2_launches;
    pl.decoded = true;
    pl.late_done = false;
    return ZPX_OK;
}

int collect_timing(zpx_batch* b, int di) {
    zpx_ctx* ctx = b->ctx;
    DeviceCtx& dc = ctx->devs[di];
    DevicePlan& pl = b->plans[di];
    if (pl.images.empty() || !pl.decoded) return ZPX_OK;
    CU(ctx, cudaSetDevice(dc.dev));
    CU(ctx, cudaEventSynchronize(dc.ev[3]));
    float a = 0, c = 0, d = 0, t = 0;
    cudaEventElapsedTime(&a, dc.ev[0], dc.ev[1]);
    cudaEventElapsedTime(&c, dc.ev[1], dc.ev[2]);
    cudaEventElapsedTime(&d, dc.ev[2], dc.ev[3]);
    cudaEventElapsedTime(&t, dc.ev[0], dc.ev[3]);
    pl.timing.entropy_ms = a;
    pl.timing.idct_fused_ms = c;
    pl.timing.idct_ms = c + d;
    pl.timing.total_ms = t;
    return ZPX_OK;
}

// Results are read on the context's own stream; the kernels may have been launched on the caller's
// (zpx_batch_decode(b, stream)): order the former after the latter's last launch (event recorded there).
cudaError_t after_decode(DeviceCtx& dc) { return cudaStreamWaitEvent(dc.stream, dc.ev[3], 0); }

// combine header status, host-side pending errors and the device error records
// Frames in which a scan ended inside an End-Of-Band run (corrupt streams only; in sequential frames the symbol itself is
// already out of place).  The reference keeps the
// run in the decoder (decoder.zig:144, reset only at a restart marker :1451), so the next scan starts by skipping
// blocks; the batch decode runs the scans of a frame level by level, side by side, each from a run of zero, and flags
// such a frame instead.  Here the flagged frames of one device are decoded again the reference's way: coefficients
// zeroed, scans launched ONE BY ONE in file order (k3_progressive; k1_lane_per_interval for sequential frames), the run handed from a scan's
// last interval to the next scan's first through two words per image (read / written alternately), then reconstructed
// by the unfused kernels.  Slow (a launch per scan) and rare.
int rescue_eob_carry(zpx_batch* b, size_t di, const std::vector<uint32_t>& slots) {
    zpx_ctx* ctx = b->ctx;
    DeviceCtx& dc = ctx->devs[di];
    DevicePlan& pl = b->plans[di];
    cudaStream_t st = dc.stream;
    uint8_t* desc = (uint8_t*)dc.desc.p;
    const size_t nimg = pl.imgs.size();
    std::vector<char> flagged(nimg, 0);
    for (uint32_t k : slots) flagged[k] = 1;
    std::vector<std::vector<uint32_t>> per_scan;  // [scan ordinal] -> intervals of the flagged progressive frames
    // sequential frames (several scans, End-Of-Band-run symbols in their tables: k1_lane_per_interval): per scan the
    // contiguous range of its intervals, one launch each
    struct Range { uint32_t first, last, count; };
    std::vector<std::map<uint32_t, Range>> seq_scan;  // [scan ordinal][image slot]
    for (uint32_t i = 0; i < (uint32_t)pl.ivs.size(); i++) {
        const ZpxScanDev& sc = pl.scans[pl.ivs[i].scan];
        if (!flagged[sc.img]) continue;
        const size_t so = (size_t)sc.scan_index;
        if (pl.imgs[sc.img].progressive) {
            if (per_scan.size() <= so) per_scan.resize(so + 1);
            per_scan[so].push_back(i);
        } else {
            if (seq_scan.size() <= so) seq_scan.resize(so + 1);
            auto it = seq_scan[so].find(sc.img);
            if (it == seq_scan[so].end()) seq_scan[so][sc.img] = Range{i, i, 1};
            else { it->second.first = std::min(it->second.first, i); it->second.last = std::max(it->second.last, i); it->second.count++; }
        }
    }
    for (const auto& m : seq_scan)
        for (const auto& kv : m)
            if (kv.second.last - kv.second.first + 1 != kv.second.count) return ZPX_OK;  // (not contiguous: leave the frames refused)
    size_t max_list = slots.size();
    for (const auto& l : per_scan) max_list = std::max(max_list, l.size());
    CU(ctx, dc.carry.ensure(2 * nimg * sizeof(uint32_t)));
    CU(ctx, dc.late_list.ensure(max_list * sizeof(uint32_t)));
    CU(ctx, cudaMemsetAsync(dc.carry.p, 0, 2 * nimg * sizeof(uint32_t), st));
    int max_blocks = 0;
    size_t max_pixels = 0;
    std::vector<uint32_t> late, eager;
    for (uint32_t k : slots) {
        const ZpxImageDev& im = pl.imgs[k];
        const ZpxParsed& p = b->parsed[pl.images[k]];
        const size_t nb = (size_t)p.mxx * p.myy * im.bpm;
        CU(ctx, cudaMemsetAsync((uint8_t*)dc.coef.p + im.coef_base * 128, 0, nb * 128, st));
        CU(ctx, cudaMemsetAsync((unsigned long long*)dc.status.p + k, 0xff, sizeof(unsigned long long), st));
        max_blocks = std::max(max_blocks, (int)nb);
        max_pixels = std::max(max_pixels, (size_t)p.width * p.height);
        (im.fused && pl.native == 0 ? late : eager).push_back(k);
    }
    K1Params k1;
    memset(&k1, 0, sizeof(k1));
    k1.blob = (const uint8_t*)dc.blob.p;
    k1.ublob = (const uint8_t*)dc.ublob.p;
    k1.ivs = (const ZpxIntervalDev*)(desc + pl.off_ivs);
    k1.scans = (const ZpxScanDev*)(desc + pl.off_scans);
    k1.imgs = (const ZpxImageDev*)(desc + pl.off_imgs);
    k1.huff = (const ZpxHuffDev*)(desc + pl.off_huff);
    k1.coef = (uint4*)dc.coef.p;
    k1.status = (unsigned long long*)dc.status.p;
    k1.img_flags = (uint32_t*)((uint8_t*)dc.status.p + align_up(nimg * sizeof(unsigned long long), 256));
    uint32_t* cw = (uint32_t*)dc.carry.p;
    for (size_t s = 0; s < std::max(per_scan.size(), seq_scan.size()); s++) {
        k1.eob_in = cw + (s & 1) * nimg;
        k1.eob_out = cw + ((s + 1) & 1) * nimg;
        if (s < per_scan.size() && !per_scan[s].empty()) {
            CU(ctx, cudaMemcpyAsync(dc.late_list.p, per_scan[s].data(), per_scan[s].size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            CU(ctx, k3_launch_progressive(k1, (const uint32_t*)dc.late_list.p, (int)per_scan[s].size(), st));
            ctx->launches++;
        }
        if (s < seq_scan.size()) {
            for (const auto& kv : seq_scan[s]) {
                K1Params kl = k1;
                kl.ivs = k1.ivs + kv.second.first;
                kl.n_iv = (int)kv.second.count;
                CU(ctx, k1_launch_lane_per_interval(kl, dc.sm_count, st));
                ctx->launches++;
            }
        }
    }
    // reconstruction (dequantise + IDCT -> planes -> colour) of those frames
    K2GParams kg{};
    kg.coef = (const int16_t*)dc.coef.p;
    kg.out = (uint8_t*)dc.out.p;
    kg.imgs = k1.imgs;
    kg.quant = (const ZpxQuantDev*)(desc + pl.off_quant);
    kg.list = (const uint32_t*)dc.late_list.p;
    for (int pass = 0; pass < 2; pass++) {
        const std::vector<uint32_t>& l = pass == 0 ? eager : late;
        if (l.empty()) continue;
        if (pass == 1) {
            // frames that took the fused kernel have no planes of their own: the on-demand plane buffer serves as
            // scratch (zpx_batch_fetch_native makes every frame's planes again, from the coefficients written above)
            CU(ctx, dc.planes_late.ensure(pl.late_plane_bytes + 256));
            for (uint32_t k : l) {
                zpx_image_info info;
                zpx_fill_info(b->parsed[pl.images[k]], &info);
                CU(ctx, cudaMemsetAsync((uint8_t*)dc.planes_late.p + pl.imgs[k].plane_off[0], 0, info.native_len, st));
            }
            pl.late_done = false;
        }
        kg.planes = (uint8_t*)(pass == 0 ? dc.planes.p : dc.planes_late.p);
        CU(ctx, cudaMemcpyAsync(dc.late_list.p, l.data(), l.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (pl.native == 2) CU(ctx, k2g_launch_planes(kg, (int)l.size(), max_blocks, st));
        else CU(ctx, k2g_launch(kg, (int)l.size(), max_blocks, max_pixels, st));
        ctx->launches += pl.native == 2 ? 1 : 2;
    }
    CU(ctx, cudaMemcpyAsync(dc.hstatus.p, dc.status.p, nimg * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return ZPX_OK;
}

int finalize_status(zpx_batch* b) {
    if (b->status_ready) return ZPX_OK;
    zpx_ctx* ctx = b->ctx;
    for (size_t di = 0; di < b->plans.size(); di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty() || !pl.decoded) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, dc.hstatus.ensure(pl.imgs.size() * sizeof(unsigned long long)));
        CU(ctx, after_decode(dc));
        if (pl.conv_flag != nullptr) {
            // a decode on the caller's stream whose synchronisation sweeps were enqueued blind: did they converge?
            int* hflag = (int*)dc.hflag.p;
            CU(ctx, cudaMemcpyAsync(hflag, pl.conv_flag, sizeof(int), cudaMemcpyDeviceToHost, dc.stream));
            CU(ctx, cudaStreamSynchronize(dc.stream));
            if (*hflag != 0) {
                int e = decode_on_device(b, (int)di, nullptr, true);  // (rare: more sweeps were needed; the host loop does them)
                if (e) return e;
                CU(ctx, after_decode(dc));
            }
            pl.conv_flag = nullptr;
        }
        CU(ctx, cudaMemcpyAsync(dc.hstatus.p, dc.status.p, pl.imgs.size() * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, dc.stream));
        CU(ctx, cudaStreamSynchronize(dc.stream));
        const unsigned long long* hs = (const unsigned long long*)dc.hstatus.p;
        int failed = 0;
        // first: progressive frames that were flagged because a scan ended inside an End-Of-Band run are decoded again,
        // scan by scan with the run carried (the device statuses in hs are read again afterwards)
        {
            std::vector<uint32_t> carry;
            for (size_t k = 0; k < pl.images.size(); k++)
                if (hs[k] != ZPX_STATUS_NONE && (int)(hs[k] & 0xff) == ZPX_E_UNSUPPORTED_STREAM) carry.push_back((uint32_t)k);
            if (!carry.empty() && !getenv("ZPX_NO_EOB_RESCUE")) {
                int e = rescue_eob_carry(b, di, carry);
                if (e) return e;
            }
        }
        for (size_t k = 0; k < pl.images.size(); k++) {
            const int bi = pl.images[k];
            const ZpxParsed& p = b->parsed[bi];
            int st = 0;
            if (hs[k] != ZPX_STATUS_NONE) st = (int)(hs[k] & 0xff);
            if (hs[k] != ZPX_STATUS_NONE && getenv("ZPX_DEBUG_STATUS"))
                fprintf(stderr, "zpx: image %d device status: scan %llu block %llu code %d\n", bi, hs[k] >> 48,
                        (hs[k] >> 8) & 0xffffffffffull, st);
            // errors the host queued (found after the entropy data they follow) come after every real device error;
            // ZPX_E_COEF_RANGE is not an error of the reference, which would go on and return the queued one
            // (only the sequential kernels' deferred report, scan field 0x7fff; after a wrapped coefficient in a
            // progressive frame nothing later can be trusted and ZPX_E_COEF_RANGE stands)
            if (st == 0 || (st == ZPX_E_COEF_RANGE && (hs[k] >> 48) == 0x7fffull)) {
                int host = 0;
                for (const ZpxScanHost& s : p.scans)
                    if (s.pending_err) { host = s.pending_err; break; }
                if (host == 0) host = p.trailing_err;
                if (host) st = host;
            }
            b->status[bi] = st;
            if (st) failed++;
        }
        pl.timing.images_failed = failed;
    }
    b->status_ready = true;
    return ZPX_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------
extern "C" {

int32_t zpx_abi_version(void) { return ZPX_ABI_VERSION; }

const char* zpx_error_name(int32_t code) {
    if (code >= 0 && code <= ZPX_E_REF_LAST) return kErrNames[code];
    switch (code) {
        case ZPX_E_CUDA: return "CudaFailure";
        case ZPX_E_NO_DEVICE: return "NoCudaDevice";
        case ZPX_E_INVALID_ARG: return "InvalidArgument";
        case ZPX_E_BAD_STATE: return "BadState";
        case ZPX_E_COEF_RANGE: return "CoefficientOutOfRange";
        case ZPX_E_UNSUPPORTED_STREAM: return "UnsupportedStream";
        case ZPX_E_MALFORMED_TABLE: return "MalformedHuffmanTable";
    }
    return "?";
}

int32_t zpx_probe(const uint8_t* buf, size_t len, zpx_image_info* out) {
    if (!out || (!buf && len)) return ZPX_E_INVALID_ARG;
    ZpxParsed p;
    zpx_parse_jpeg(buf, len, true, &p);
    zpx_fill_info(p, out);
    // decodeConfig reports 4-component frames as YCbCr too (decoder.zig:210-215); variant stays what load returns
    return p.status;
}

int32_t zpx_parse_report_of(const uint8_t* buf, size_t len, zpx_image_info* info, zpx_parse_report* rep) {
    if (!rep || (!buf && len)) return ZPX_E_INVALID_ARG;
    ZpxParsed p;
    zpx_parse_jpeg(buf, len, false, &p);
    if (info) zpx_fill_info(p, info);
    memset(rep, 0, sizeof(*rep));
    rep->status = p.status;
    rep->n_scans = (int32_t)p.scans.size();
    rep->pending_after_interval = -1;
    rep->lane_script = p.progressive && !p.scans.empty() && lane_script_ok(p) ? 1 : 0;
    for (const ZpxScanHost& s : p.scans) {
        rep->n_intervals += (int32_t)s.intervals.size();
        if (s.pending_err) {
            rep->pending_err = s.pending_err;
            rep->pending_after_interval = s.err_after_interval;
        }
        if (!s.intervals.empty()) rep->entropy_bytes += s.intervals.back().limit - s.intervals.front().start;
        rep->pieces_ok = 1;
        for (const ZpxIntervalHost& iv : s.intervals) {
            rep->stuffed_bytes += iv.n_stuffed;
            rep->unstuffed_bytes += iv.limit - iv.start - iv.n_stuffed;
            size_t at = iv.start;
            uint32_t u = 0;
            for (uint32_t q = 0; q < iv.n_segs; q++) {
                const ZpxSegHost& sg = s.segs[iv.seg_first + q];
                rep->n_pieces++;
                rep->max_piece = std::max<int32_t>(rep->max_piece, (int32_t)sg.len);
                bool ok = sg.src == at && sg.uoff == u && sg.len > 0;
                // inside an interval every 0xFF opens a pair: a piece that starts with 0x00 after 0xFF splits one
                if (sg.src > iv.start && buf[sg.src] == 0x00 && buf[sg.src - 1] == 0xff) ok = false;
                uint32_t pairs = 0;
                for (size_t k = sg.src; k + 1 < sg.src + sg.len; k++)
                    if (buf[k] == 0xff && buf[k + 1] == 0x00) { pairs++; k++; }
                at += sg.len;
                u += sg.len - pairs;
                if (!ok) rep->pieces_ok = 0;
            }
            if (at != iv.limit || u != iv.limit - iv.start - iv.n_stuffed) rep->pieces_ok = 0;
        }
    }
    rep->trailing_err = p.trailing_err;
    rep->fused = (p.status == 0 && fused_eligible(p, recon_mask_of(p), 0)) ? 1 : 0;  // (for RGBA output, 8-bit quantisers)
    rep->mode = p.mode;
    return ZPX_OK;
}

int32_t zpx_partition(const uint64_t* weights, int32_t n, int32_t n_devices, int32_t* device_of) {
    if (n < 0 || n_devices <= 0 || (n > 0 && (!weights || !device_of))) return ZPX_E_INVALID_ARG;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) total += weights[i];
    uint64_t acc = 0;
    for (int i = 0; i < n; i++) {
        if (weights[i] == 0) {
            device_of[i] = -1;
            continue;
        }
        // the device whose share of the total weight contains this image's midpoint: contiguous, monotone
        device_of[i] = (int32_t)std::min<uint64_t>((uint64_t)n_devices - 1, (acc + weights[i] / 2) * (uint64_t)n_devices / total);
        acc += weights[i];
    }
    return ZPX_OK;
}

int32_t zpx_ctx_create(const int32_t* device_ids, int32_t n_devices, zpx_ctx** out) {
    if (!out) return ZPX_E_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return ZPX_E_NO_DEVICE;
    zpx_ctx* c = new (std::nothrow) zpx_ctx();
    if (!c) return ZPX_E_OutOfMemory;
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_devices);
    for (int id : ids) {
        if (id < 0 || id >= count) {
            zpx_ctx_destroy(c);
            return ZPX_E_INVALID_ARG;
        }
        DeviceCtx d;
        d.dev = id;
        if ((e = cudaSetDevice(id)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess) {
            zpx_ctx_destroy(c);
            return ZPX_E_CUDA;
        }
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, id);
        {
            size_t fr = 0, tot = 0;
            if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) d.total_mem = tot;
            else (void)cudaGetLastError();
        }
        for (auto& ev : d.ev) cudaEventCreate(&ev);
        c->devs.push_back(d);
    }
    *out = c;
    return ZPX_OK;
}

void zpx_ctx_destroy(zpx_ctx* c) {
    if (!c) return;
    for (zpx_ctx* sh : c->shadows) zpx_ctx_destroy(sh);
    for (DeviceCtx& d : c->devs) {
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        d.blob.release();
        d.ublob.release();
        d.coef.release();
        d.out.release();
        d.planes.release();
        d.desc.release();
        d.status.release();
        d.subs.release();
        d.scratch.release();
        d.planes_late.release();
        d.late_list.release();
        d.carry.release();
        d.hflag.release();
        d.stage.release();
        d.hdesc.release();
        d.hstatus.release();
        for (auto& ev : d.ev)
            if (ev) cudaEventDestroy(ev);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
}

int32_t zpx_ctx_num_devices(const zpx_ctx* c) { return c ? (int32_t)c->devs.size() : 0; }
int32_t zpx_last_cuda_error(const zpx_ctx* c) { return c ? c->last_cuda : 0; }
const char* zpx_last_cuda_error_string(const zpx_ctx* c) { return c ? c->last_cuda_str.c_str() : ""; }
uint64_t zpx_ctx_kernel_launches(const zpx_ctx* c) { return c ? c->launches.load() : 0; }

int32_t zpx_ctx_set_option(zpx_ctx* c, int32_t option, int64_t value) {
    if (!c) return ZPX_E_INVALID_ARG;
    switch (option) {
        case ZPX_OPT_ENTROPY_MODE: c->opt_entropy_mode = value; return ZPX_OK;
        case ZPX_OPT_FORCE_GENERIC: c->opt_force_generic = value; return ZPX_OK;
        case ZPX_OPT_SUBSEQ_BYTES: c->opt_subseq = value; return ZPX_OK;
        case ZPX_OPT_PIPELINE_CHUNK: c->opt_pipeline_chunk = value; return ZPX_OK;
        case ZPX_OPT_PIPELINE_RAMP: c->opt_pipeline_ramp = value; return ZPX_OK;
        case ZPX_OPT_TEST_WIDE: c->opt_test_wide = value; return ZPX_OK;
        case ZPX_OPT_PROGRESSIVE_MODE: c->opt_progressive_mode = value; return ZPX_OK;
        case ZPX_OPT_K2_DENSE: c->opt_k2_dense = value; return ZPX_OK;
        case ZPX_OPT_GATED_SWEEPS:
            if (value < -1 || value > 8) return ZPX_E_INVALID_ARG;  // (-1: test hook, the rounds count as not converged)
            c->opt_gated_sweeps = value;
            return ZPX_OK;
        case ZPX_OPT_NATIVE_PLANES:
            if (value < 0 || value > 2) return ZPX_E_INVALID_ARG;
            c->opt_native = value;
            return ZPX_OK;
        case ZPX_OPT_PIPELINE_WORKERS:
            if (value < 1 || value > 8) return ZPX_E_INVALID_ARG;
            c->opt_pipeline_workers = value;
            return ZPX_OK;
    }
    return ZPX_E_INVALID_ARG;
}

void* zpx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void zpx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int32_t zpx_batch_open(zpx_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, int32_t n, zpx_batch** out) {
    if (!ctx || !out || n < 0 || (n > 0 && (!bufs || !lens))) return ZPX_E_INVALID_ARG;
    *out = nullptr;
    zpx_batch* b = new (std::nothrow) zpx_batch();
    if (!b) return ZPX_E_OutOfMemory;
    b->ctx = ctx;
    b->n = n;
    b->bufs.assign(bufs, bufs + n);
    b->lens.assign(lens, lens + n);
    b->parsed.resize(n);
    b->dev_of.assign(n, -1);
    b->slot_of.assign(n, -1);
    b->status.assign(n, 0);
    // per-image device memory budget: half of the smallest device (ZPX_IMAGE_BUDGET_MB overrides, tests)
    uint64_t image_budget = ~0ull;
    for (const DeviceCtx& d : ctx->devs)
        if (d.total_mem) image_budget = std::min<uint64_t>(image_budget, d.total_mem / 2);
    if (const char* e = getenv("ZPX_IMAGE_BUDGET_MB")) image_budget = (uint64_t)atoll(e) << 20;
    std::vector<uint64_t> need_of((size_t)n, 0);
    // header parse, parallel over host cores (decodeInner's marker loop; no entropy decode)
    parallel_for((size_t)n, 4, [&](size_t i) {
        if (!b->bufs[i] && b->lens[i]) {
            b->parsed[i].status = ZPX_E_INVALID_ARG;
            return;
        }
        zpx_parse_jpeg(b->bufs[i], b->lens[i], false, &b->parsed[i]);
        // An image whose buffers alone would not fit the device (a few bytes can declare 65535 x 65535) gets the
        // reference's own answer to a failed allocation (makeImg, decoder.zig:1708-1783: error.OutOfMemory) and
        // stays out of the plan, instead of failing the upload of the whole batch.
        ZpxParsed& p = b->parsed[i];
        if (p.status == 0 && p.mxx > 0 && p.myy > 0) {
            uint64_t bpm = 0;
            for (int c = 0; c < p.ncomp; c++) bpm += (uint64_t)p.h[c] * p.v[c];
            const uint64_t blocks = (uint64_t)p.mxx * p.myy * bpm;
            // coefficients (+ a progressive frame's bit maps and scratch) + planes + RGBA
            need_of[i] = blocks * (128 + 64 + (p.progressive ? 16 + 128 : 0)) + (p.progressive ? 8192 : 0) + (uint64_t)4 * p.width * p.height;
            if (need_of[i] > image_budget) p.status = ZPX_E_OutOfMemory;
        }
    });
    // schedule: contiguous index ranges balanced by entropy-coded bytes (images are independent,
    // nothing is exchanged between devices)
    const int nd = (int)ctx->devs.size();
    b->plans.resize(nd);
    std::vector<uint64_t> w(n, 0);
    uint64_t total = 0;
    for (int i = 0; i < n; i++) {
        b->status[i] = b->parsed[i].status;
        if (b->parsed[i].status != 0) continue;
        for (const ZpxScanHost& s : b->parsed[i].scans)
            if (!s.intervals.empty()) w[i] += s.intervals.back().limit - s.intervals.front().start;
        w[i] += 1024;
        total += w[i];
    }
    (void)total;
    zpx_partition(w.data(), n, nd, b->dev_of.data());
    // the same budget for a device's whole share: images that no longer fit are refused one by one (the caller can
    // submit them again in a later batch), the others decode
    std::vector<uint64_t> used((size_t)nd, 0);
    for (int i = 0; i < n; i++) {
        const int d = b->dev_of[i];
        if (d < 0) continue;
        const uint64_t cap = getenv("ZPX_IMAGE_BUDGET_MB") ? image_budget : (ctx->devs[d].total_mem ? ctx->devs[d].total_mem / 10 * 8 : ~0ull);
        if (used[(size_t)d] + need_of[i] > cap) {
            b->status[i] = b->parsed[i].status = ZPX_E_OutOfMemory;
            b->dev_of[i] = -1;
            continue;
        }
        used[(size_t)d] += need_of[i];
    }
    for (int i = 0; i < n; i++) {
        const int d = b->dev_of[i];
        if (d < 0) continue;
        b->slot_of[i] = (int)b->plans[d].images.size();
        b->plans[d].images.push_back(i);
    }
    for (int d = 0; d < nd; d++) build_plan(b, d);
    *out = b;
    return ZPX_OK;
}

int32_t zpx_batch_size(const zpx_batch* b) { return b ? b->n : 0; }

int32_t zpx_batch_info(const zpx_batch* b, int32_t i, zpx_image_info* out) {
    if (!b || !out || i < 0 || i >= b->n) return ZPX_E_INVALID_ARG;
    zpx_fill_info(b->parsed[i], out);
    out->device = b->dev_of[i];
    return ZPX_OK;
}

int32_t zpx_batch_upload(zpx_batch* b) {
    if (!b) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    const int nd = (int)ctx->devs.size();
    // a context has ONE set of device buffers: uploading a batch evicts the previous one
    if (ctx->resident && ctx->resident != b) {
        for (DeviceCtx& dc : ctx->devs) {
            cudaSetDevice(dc.dev);
            cudaStreamSynchronize(dc.stream);
        }
    }
    ctx->resident = b;
    // allocate and stage
    for (int di = 0; di < nd; di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, dc.blob.ensure(pl.blob_bytes + 64));
        if (pl.ublob_bytes) CU(ctx, dc.ublob.ensure(pl.ublob_bytes + 256));
        CU(ctx, dc.coef.ensure(pl.coef_blocks * 128 + 256));
        CU(ctx, dc.out.ensure(pl.out_bytes + 256));
        if (pl.plane_bytes) CU(ctx, dc.planes.ensure(pl.plane_bytes + 256));
        CU(ctx, dc.desc.ensure(pl.desc_bytes + 256));
        // per image: error key (u64), then flags (u32)
        CU(ctx, dc.status.ensure(align_up(pl.imgs.size() * sizeof(unsigned long long), 256) + pl.imgs.size() * sizeof(uint32_t) + 256));
        CU(ctx, dc.stage.ensure(pl.blob_bytes + 64));
        CU(ctx, dc.hdesc.ensure(pl.desc_bytes + 256));
        // descriptors
        uint8_t* hd = (uint8_t*)dc.hdesc.p;
        memcpy(hd + pl.off_imgs, pl.imgs.data(), pl.imgs.size() * sizeof(ZpxImageDev));
        memcpy(hd + pl.off_scans, pl.scans.data(), pl.scans.size() * sizeof(ZpxScanDev));
        memcpy(hd + pl.off_ivs, pl.ivs.data(), pl.ivs.size() * sizeof(ZpxIntervalDev));
        memcpy(hd + pl.off_huff, pl.huff.data(), pl.huff.size() * sizeof(ZpxHuffDev));
        memcpy(hd + pl.off_quant, pl.quant.data(), pl.quant.size() * sizeof(ZpxQuantDev));
        memcpy(hd + pl.off_generic, pl.generic.data(), pl.generic.size() * sizeof(uint32_t));
        memcpy(hd + pl.off_warps, pl.warps.data(), pl.warps.size() * sizeof(ZpxWarpDev));
        memcpy(hd + pl.off_segs, pl.segs.data(), pl.segs.size() * sizeof(ZpxSegDev));
        for (size_t k = 0; k < pl.prog_lists.size(); k++)
            memcpy(hd + pl.prog_off[k], pl.prog_lists[k].data(), pl.prog_lists[k].size() * sizeof(uint32_t));
        if (pl.sub_mode) {
            CU(ctx, dc.subs.ensure(align_up(pl.n_subs, 64) * 40 + 256));
            CU(ctx, dc.hflag.ensure(64));
        }
        for (const FusedGroup& g : pl.groups) memcpy(hd + g.tiles_off, g.tiles.data(), g.tiles.size() * sizeof(ZpxTileDev));
        // entropy-coded segments into pinned staging, parallel over host cores
        uint8_t* stg = (uint8_t*)dc.stage.p;
        parallel_for(pl.copies.size(), 8, [&](size_t k) {
            const DevicePlan::Copy& c = pl.copies[k];
            memcpy(stg + c.dst, b->bufs[c.img] + c.src, c.len);
            memset(stg + c.dst + c.len, 0, 8);
        });
        CU(ctx, cudaEventRecord(dc.ev[4], dc.stream));
        CU(ctx, cudaMemcpyAsync(dc.desc.p, dc.hdesc.p, pl.desc_bytes, cudaMemcpyHostToDevice, dc.stream));
        if (pl.blob_bytes) CU(ctx, cudaMemcpyAsync(dc.blob.p, dc.stage.p, pl.blob_bytes, cudaMemcpyHostToDevice, dc.stream));
        CU(ctx, cudaEventRecord(dc.ev[5], dc.stream));
    }
    for (int di = 0; di < nd; di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, cudaStreamSynchronize(dc.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, dc.ev[4], dc.ev[5]);
        pl.timing.h2d_ms = ms;
        pl.uploaded = true;
    }
    return ZPX_OK;
}

int32_t zpx_batch_decode(zpx_batch* b, void* stream) {
    if (!b) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    const int nd = (int)ctx->devs.size();
    if (stream && nd != 1) return ZPX_E_INVALID_ARG;
    if (b->n > 0 && ctx->resident != b) return ZPX_E_BAD_STATE;  // another batch was uploaded on this context since
    for (int di = 0; di < nd; di++)
        if (!b->plans[di].images.empty() && !b->plans[di].uploaded) return ZPX_E_BAD_STATE;
    b->status_ready = false;
    if (nd == 1) {
        int e = decode_on_device(b, 0, (cudaStream_t)stream);
        if (e) return e;
    } else {
        // one host thread per device: the self-synchronising decoder reads a flag back between sweeps, which
        // would otherwise serialise the devices
        std::vector<int> rc((size_t)nd, 0);
        std::vector<std::thread> th;
        for (int di = 1; di < nd; di++) th.emplace_back([&, di] { rc[(size_t)di] = decode_on_device(b, di, nullptr); });
        rc[0] = decode_on_device(b, 0, nullptr);
        for (std::thread& t : th) t.join();
        for (int di = 0; di < nd; di++)
            if (rc[(size_t)di]) return rc[(size_t)di];
    }
    if (stream) return ZPX_OK;  // caller synchronises its own stream
    for (int di = 0; di < nd; di++) {
        int e = collect_timing(b, di);
        if (e) return e;
    }
    return ZPX_OK;
}

int32_t zpx_batch_status(zpx_batch* b, int32_t* status) {
    if (!b) return ZPX_E_INVALID_ARG;
    int e = finalize_status(b);
    if (e) return e;
    if (status) memcpy(status, b->status.data(), sizeof(int32_t) * b->n);
    return ZPX_OK;
}

int32_t zpx_batch_timing(const zpx_batch* b, int32_t di, zpx_timing* out) {
    if (!b || !out || di < 0 || di >= (int)b->plans.size()) return ZPX_E_INVALID_ARG;
    // timing of a user-stream decode is collected lazily
    if (b->plans[di].decoded) collect_timing(const_cast<zpx_batch*>(b), di);
    *out = b->plans[di].timing;
    return ZPX_OK;
}

const void* zpx_batch_device_rgba(const zpx_batch* b, int32_t i) {
    if (!b || i < 0 || i >= b->n || b->dev_of[i] < 0) return nullptr;
    const DevicePlan& pl = b->plans[b->dev_of[i]];
    if (!pl.decoded || b->ctx->resident != b) return nullptr;  // (evicted by a later upload: no stale pointers)
    return (const uint8_t*)b->ctx->devs[b->dev_of[i]].out.p + pl.out_off[b->slot_of[i]];
}

int32_t zpx_batch_fetch_rgba(zpx_batch* b, uint8_t* const* out, const size_t* out_stride, int32_t* status) {
    if (!b || (b->n > 0 && !out)) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    const int nd = (int)ctx->devs.size();
    if (b->n > 0 && ctx->resident != b) {
        bool any = false;
        for (int di = 0; di < nd; di++) any = any || !b->plans[di].images.empty();
        if (any) return ZPX_E_BAD_STATE;  // its results were evicted by a later upload on this context
    }
    for (int di = 0; di < nd; di++)
        if (!b->plans[di].images.empty() && !b->plans[di].decoded) return ZPX_E_BAD_STATE;
    int e = finalize_status(b);
    if (e) return e;
    for (int di = 0; di < nd; di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, after_decode(dc));
        CU(ctx, cudaEventRecord(dc.ev[6], dc.stream));
        size_t k = 0;
        while (k < pl.images.size()) {
            const int bi = pl.images[k];
            const ZpxParsed& p = b->parsed[bi];
            const size_t row = (size_t)4 * p.width;
            const size_t len = row * p.height;
            if (!out[bi] || b->status[bi] != 0) { k++; continue; }
            if (pl.native == 2 && pl.imgs[k].fused) return ZPX_E_BAD_STATE;  // decoded with ZPX_OPT_NATIVE_PLANES = 2: no RGBA
            const uint8_t* src = (const uint8_t*)dc.out.p + pl.out_off[k];
            if (out_stride && out_stride[bi] != 0 && out_stride[bi] != row) {
                CU(ctx, cudaMemcpy2DAsync(out[bi], out_stride[bi], src, row, row, p.height, cudaMemcpyDeviceToHost, dc.stream));
                k++;
                continue;
            }
            // merge runs that are contiguous on both sides into one copy
            size_t run = len, k2 = k + 1;
            while (k2 < pl.images.size()) {
                const int bj = pl.images[k2];
                const ZpxParsed& pj = b->parsed[bj];
                if (!out[bj] || b->status[bj] != 0) break;
                if (out_stride && out_stride[bj] != 0 && out_stride[bj] != (size_t)4 * pj.width) break;
                if (pl.out_off[k2] != pl.out_off[k] + run || out[bj] != out[bi] + run) break;
                run += (size_t)4 * pj.width * pj.height;
                k2++;
            }
            CU(ctx, cudaMemcpyAsync(out[bi], src, run, cudaMemcpyDeviceToHost, dc.stream));
            k = k2;
        }
        CU(ctx, cudaEventRecord(dc.ev[7], dc.stream));
    }
    for (int di = 0; di < nd; di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, cudaStreamSynchronize(dc.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, dc.ev[6], dc.ev[7]);
        pl.timing.d2h_ms = ms;
    }
    if (status) memcpy(status, b->status.data(), sizeof(int32_t) * b->n);
    return ZPX_OK;
}

int32_t zpx_batch_fetch_native(zpx_batch* b, uint8_t* const* out, int32_t* status) {
    if (!b || (b->n > 0 && !out)) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    if (b->n > 0 && ctx->resident != b) {
        for (const DevicePlan& pl : b->plans)
            if (!pl.images.empty()) return ZPX_E_BAD_STATE;  // its results were evicted by a later upload on this context
    }
    int e = finalize_status(b);
    if (e) return e;
    for (size_t di = 0; di < b->plans.size(); di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        if (!pl.decoded) return ZPX_E_BAD_STATE;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, after_decode(dc));
        // planes of fused images that were decoded without ZPX_OPT_NATIVE_PLANES: the coefficients are still resident,
        // the unfused IDCT kernel writes the planes now (makeImg's layout, zeros where the reference reconstructs nothing)
        bool want_late = false;
        for (uint32_t k : pl.late) want_late = want_late || (out[pl.images[k]] && b->status[pl.images[k]] == 0);
        if (want_late && !pl.late_done) {
            CU(ctx, dc.planes_late.ensure(pl.late_plane_bytes + 256));
            CU(ctx, dc.late_list.ensure(pl.late.size() * sizeof(uint32_t)));
            CU(ctx, cudaMemsetAsync(dc.planes_late.p, 0, pl.late_plane_bytes, dc.stream));
            CU(ctx, cudaMemcpyAsync(dc.late_list.p, pl.late.data(), pl.late.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, dc.stream));
            K2GParams kg{};
            kg.coef = (const int16_t*)dc.coef.p;
            kg.planes = (uint8_t*)dc.planes_late.p;
            kg.imgs = (const ZpxImageDev*)((const uint8_t*)dc.desc.p + pl.off_imgs);
            kg.quant = (const ZpxQuantDev*)((const uint8_t*)dc.desc.p + pl.off_quant);
            const int total = (int)pl.late.size();
            for (int i0 = 0; i0 < total; i0 += 32768) {
                kg.list = (const uint32_t*)dc.late_list.p + i0;
                CU(ctx, k2g_launch_planes(kg, std::min(32768, total - i0), pl.late_max_blocks, dc.stream));
                ctx->launches++;
            }
            pl.late_done = true;
        }
        CU(ctx, cudaEventRecord(dc.ev[6], dc.stream));
        size_t k = 0;
        while (k < pl.images.size()) {
            const int i = pl.images[k];
            const ZpxImageDev& im = pl.imgs[k];
            const ZpxParsed& p = b->parsed[i];
            if (!out[i] || b->status[i] != 0) { k++; continue; }
            zpx_image_info info;
            zpx_fill_info(p, &info);
            const bool late = im.fused && pl.native == 0;
            if (p.variant == ZPX_VARIANT_RGBA) {
                CU(ctx, cudaMemcpyAsync(out[i], (const uint8_t*)dc.out.p + im.out_off, info.native_len, cudaMemcpyDeviceToHost, dc.stream));
                k++;
            } else if (p.variant == ZPX_VARIANT_CMYK) {
                // Image{.CMYK}: applyBlack's interleave, built from the planes on demand (4-component frames
                // always take the unfused path)
                CU(ctx, dc.scratch.ensure(info.native_len));
                K2GParams kg{};
                kg.planes = (uint8_t*)dc.planes.p;
                kg.imgs = (const ZpxImageDev*)((const uint8_t*)dc.desc.p + pl.off_imgs);
                CU(ctx, k2g_launch_cmyk_native(kg, (uint32_t)k, (size_t)p.width * p.height, (uint8_t*)dc.scratch.p, dc.stream));
                ctx->launches++;
                CU(ctx, cudaMemcpyAsync(out[i], dc.scratch.p, info.native_len, cudaMemcpyDeviceToHost, dc.stream));
                CU(ctx, cudaStreamSynchronize(dc.stream));  // scratch is reused by the next image
                k++;
            } else {
                // Gray / YCbCr planes: merge runs that are contiguous on both sides into one copy
                size_t run = info.native_len, k2 = k + 1;
                while (k2 < pl.images.size()) {
                    const int j = pl.images[k2];
                    const ZpxParsed& pj = b->parsed[j];
                    const ZpxImageDev& imj = pl.imgs[k2];
                    if (!out[j] || b->status[j] != 0) break;
                    if (pj.variant != ZPX_VARIANT_GRAY && pj.variant != ZPX_VARIANT_YCBCR) break;
                    if ((imj.fused && pl.native == 0) != late) break;
                    if (imj.plane_off[0] != im.plane_off[0] + run || out[j] != out[i] + run) break;
                    zpx_image_info ij;
                    zpx_fill_info(pj, &ij);
                    run += ij.native_len;
                    k2++;
                }
                CU(ctx, cudaMemcpyAsync(out[i], (const uint8_t*)(late ? dc.planes_late.p : dc.planes.p) + im.plane_off[0], run, cudaMemcpyDeviceToHost, dc.stream));
                k = k2;
            }
        }
        CU(ctx, cudaEventRecord(dc.ev[7], dc.stream));
    }
    for (size_t di = 0; di < b->plans.size(); di++) {
        DevicePlan& pl = b->plans[di];
        if (pl.images.empty()) continue;
        DeviceCtx& dc = ctx->devs[di];
        CU(ctx, cudaSetDevice(dc.dev));
        CU(ctx, cudaStreamSynchronize(dc.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, dc.ev[6], dc.ev[7]);
        pl.timing.d2h_ms = ms;
    }
    if (status) memcpy(status, b->status.data(), sizeof(int32_t) * b->n);
    return ZPX_OK;
}

int32_t zpx_batch_fetch_coefficients(zpx_batch* b, int32_t i, int16_t* out, size_t cap_blocks, size_t* n_blocks) {
    if (!b || i < 0 || i >= b->n || b->dev_of[i] < 0) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    const int di = b->dev_of[i];
    DevicePlan& pl = b->plans[di];
    if (!pl.decoded || ctx->resident != b) return ZPX_E_BAD_STATE;
    DeviceCtx& dc = ctx->devs[di];
    const ZpxImageDev& im = pl.imgs[b->slot_of[i]];
    const ZpxParsed& p = b->parsed[i];
    const size_t nb = (size_t)p.mxx * p.myy * im.bpm;
    if (n_blocks) *n_blocks = nb;
    if (!out) return ZPX_OK;
    if (cap_blocks < nb) return ZPX_E_INVALID_ARG;
    CU(ctx, cudaSetDevice(dc.dev));
    CU(ctx, after_decode(dc));
    CU(ctx, cudaStreamSynchronize(dc.stream));
    CU(ctx, cudaMemcpy(out, (const uint8_t*)dc.coef.p + im.coef_base * 128, nb * 128, cudaMemcpyDeviceToHost));
    // undo the row swizzle: block with component-x index bx stores row r at slot r ^ (bx & 7)
    std::vector<int16_t> tmp(64);
    for (size_t k = 0; k < nb; k++) {
        int bx;
        if (im.layout == ZPX_LAYOUT_INTERLEAVED) {
            const size_t mcu = k / im.bpm;
            const int slot = (int)(k % im.bpm);
            int c = 0;
            for (int cc = 0; cc < im.ncomp; cc++)
                if ((int)im.blk_off[cc] <= slot) c = cc;
            const int j = slot - (int)im.blk_off[c];
            bx = im.h[c] * (int)(mcu % p.mxx) + j % im.h[c];
        } else {
            size_t kk = k;
            int c = 0;
            for (; c < im.ncomp; c++) {
                const size_t cnt = (size_t)im.comp_bw[c] * im.comp_bh[c];
                if (kk < cnt) break;
                kk -= cnt;
            }
            bx = (int)(kk % im.comp_bw[c]);
        }
        int16_t* blk = out + k * 64;
        memcpy(tmp.data(), blk, 128);
        for (int r = 0; r < 8; r++) memcpy(blk + r * 8, tmp.data() + ((r ^ (bx & 7)) * 8), 16);
    }
    return ZPX_OK;
}

// ---- test hooks: the reconstruction kernels without the entropy stage ----
int32_t zpx_batch_open_synthetic(zpx_ctx* ctx, int32_t width, int32_t height, int32_t ncomp, const uint8_t* comp_hv,
                                 const uint16_t* quant, int32_t mode, zpx_batch** out) {
    if (!ctx || !out || !comp_hv || !quant || width <= 0 || height <= 0 || width > 65535 || height > 65535) return ZPX_E_INVALID_ARG;
    if (ncomp != 1 && ncomp != 3 && ncomp != 4) return ZPX_E_INVALID_ARG;
    if (mode < ZPX_MODE_GRAY || mode > ZPX_MODE_YCCK || (ncomp == 1) != (mode == ZPX_MODE_GRAY) || (ncomp == 4) != (mode >= ZPX_MODE_CMYK))
        return ZPX_E_INVALID_ARG;
    *out = nullptr;
    zpx_batch* b = new (std::nothrow) zpx_batch();
    if (!b) return ZPX_E_OutOfMemory;
    b->ctx = ctx;
    b->n = 1;
    b->bufs.assign(1, nullptr);
    b->lens.assign(1, 0);
    b->parsed.resize(1);
    b->dev_of.assign(1, 0);
    b->slot_of.assign(1, 0);
    b->status.assign(1, 0);
    ZpxParsed& p = b->parsed[0];
    p.width = width;
    p.height = height;
    p.ncomp = ncomp;
    p.baseline = true;
    int total_hv = 0;
    for (int c = 0; c < ncomp; c++) {
        p.h[c] = ncomp == 1 ? 1 : comp_hv[c] >> 4;   // gray forces h = v = 1 (decoder.zig:559-560)
        p.v[c] = ncomp == 1 ? 1 : comp_hv[c] & 15;
        p.tq[c] = c;
        p.cid[c] = (uint8_t)(c + 1);
        total_hv += p.h[c] * p.v[c];
        const bool ok = (p.h[c] == 1 || p.h[c] == 2 || p.h[c] == 4) && (p.v[c] == 1 || p.v[c] == 2 || p.v[c] == 4);
        if (!ok || (c > 0 && (p.h[0] % p.h[c] || p.v[0] % p.v[c]))) {
            delete b;
            return ZPX_E_INVALID_ARG;
        }
    }
    if (ncomp > 1 && total_hv > 10) {
        delete b;
        return ZPX_E_INVALID_ARG;
    }
    p.jfif = mode == ZPX_MODE_YCBCR;
    p.adobe_valid = mode == ZPX_MODE_RGB || mode >= ZPX_MODE_CMYK;
    p.adobe_transform = mode == ZPX_MODE_YCCK ? 2 : 0;
    p.mxx = (width + 8 * p.h[0] - 1) / (8 * p.h[0]);
    p.myy = (height + 8 * p.v[0] - 1) / (8 * p.v[0]);
    p.saw_sos = true;
    ZpxScanHost sc;
    sc.ncomp = ncomp;
    for (int c = 0; c < ncomp; c++) {
        sc.comp[c] = c;
        for (int z = 0; z < 64; z++) {
            sc.quant[c][z] = quant[c * 64 + z];
            p.final_quant[c][z] = quant[c * 64 + z];
        }
    }
    p.scans.push_back(sc);  // no intervals: the entropy stage has nothing to do
    zpx_derive(&p);
    b->plans.resize(ctx->devs.size());
    b->plans[0].images.push_back(0);
    for (size_t d = 0; d < ctx->devs.size(); d++) build_plan(b, (int)d);
    *out = b;
    return ZPX_OK;
}

int32_t zpx_batch_set_coefficients(zpx_batch* b, int32_t i, const int16_t* blocks, size_t n_blocks) {
    if (!b || !blocks || i < 0 || i >= b->n || b->dev_of[i] < 0) return ZPX_E_INVALID_ARG;
    zpx_ctx* ctx = b->ctx;
    const int di = b->dev_of[i];
    DevicePlan& pl = b->plans[di];
    if (!pl.uploaded || ctx->resident != b) return ZPX_E_BAD_STATE;
    DeviceCtx& dc = ctx->devs[di];
    const ZpxImageDev& im = pl.imgs[b->slot_of[i]];
    const ZpxParsed& p = b->parsed[i];
    const size_t nb = (size_t)p.mxx * p.myy * im.bpm;
    if (n_blocks != nb) return ZPX_E_INVALID_ARG;
    std::vector<int16_t> sw(nb * 64);
    bool wide = ctx->opt_test_wide != 0;
    for (size_t k = 0; k < nb; k++) {
        int bx;
        if (im.layout == ZPX_LAYOUT_INTERLEAVED) {
            const size_t mcu = k / im.bpm;
            const int slot = (int)(k % im.bpm);
            int c = 0;
            for (int cc = 0; cc < im.ncomp; cc++)
                if ((int)im.blk_off[cc] <= slot) c = cc;
            bx = im.h[c] * (int)(mcu % p.mxx) + (slot - (int)im.blk_off[c]) % im.h[c];
        } else {
            size_t kk = k;
            int c = 0;
            for (; c < im.ncomp; c++) {
                const size_t cnt = (size_t)im.comp_bw[c] * im.comp_bh[c];
                if (kk < cnt) break;
                kk -= cnt;
            }
            bx = (int)(kk % im.comp_bw[c]);
        }
        const int16_t* src = blocks + k * 64;
        for (int r = 0; r < 8; r++) memcpy(&sw[k * 64 + (size_t)((r ^ (bx & 7)) * 8)], src + r * 8, 16);
        for (int z = 0; z < 64; z++) wide = wide || src[z] < -4096 || src[z] > 4095;
    }
    CU(ctx, cudaSetDevice(dc.dev));
    CU(ctx, cudaStreamSynchronize(dc.stream));
    // (on the context's stream and waited for: a plain cudaMemcpy from pageable memory returns once the bytes are
    // staged, and nothing orders its DMA before kernels on the non-blocking streams the decode uses -- seen as a rare
    // stale-coefficient mismatch in the block tests)
    CU(ctx, cudaMemcpyAsync((uint8_t*)dc.coef.p + im.coef_base * 128, sw.data(), nb * 128, cudaMemcpyHostToDevice, dc.stream));
    CU(ctx, cudaStreamSynchronize(dc.stream));
    if (pl.inject_flags.size() != pl.imgs.size()) pl.inject_flags.assign(pl.imgs.size(), 0);
    pl.inject_flags[b->slot_of[i]] = wide ? 1u : 0u;
    return ZPX_OK;
}

int32_t zpx_test_colour(zpx_ctx* ctx, int32_t mode, const uint8_t* samples, size_t n, uint8_t* rgba) {
    if (!ctx || !samples || !rgba || (mode != ZPX_MODE_YCBCR && mode != ZPX_MODE_CMYK && mode != ZPX_MODE_YCCK)) return ZPX_E_INVALID_ARG;
    if (n == 0) return ZPX_OK;
    DeviceCtx& dc = ctx->devs[0];
    const size_t in_bytes = n * (mode == ZPX_MODE_YCBCR ? 3 : 4);
    CU(ctx, cudaSetDevice(dc.dev));
    CU(ctx, cudaStreamSynchronize(dc.stream));
    CU(ctx, dc.scratch.ensure(align_up(in_bytes, 256) + n * 4));
    uint8_t* din = (uint8_t*)dc.scratch.p;
    uint8_t* dout = din + align_up(in_bytes, 256);
    CU(ctx, cudaMemcpyAsync(din, samples, in_bytes, cudaMemcpyHostToDevice, dc.stream));
    CU(ctx, k2_launch_test_colour(mode, din, n, dout, dc.stream));
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(rgba, dout, n * 4, cudaMemcpyDeviceToHost, dc.stream));
    CU(ctx, cudaStreamSynchronize(dc.stream));
    return ZPX_OK;
}

void zpx_batch_close(zpx_batch* b) {
    if (!b) return;
    if (b->ctx->resident == b) b->ctx->resident = nullptr;
    for (DeviceCtx& dc : b->ctx->devs) {
        cudaSetDevice(dc.dev);
        cudaStreamSynchronize(dc.stream);
    }
    delete b;
}

static int32_t decode_range_rgba(zpx_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, int32_t n,
                                 uint8_t* const* out, const size_t* out_stride, int32_t* status, bool native) {
    zpx_batch* b = nullptr;
    int e = zpx_batch_open(ctx, bufs, lens, n, &b);
    if (e) return e;
    e = zpx_batch_upload(b);
    if (!e) e = zpx_batch_decode(b, nullptr);
    if (!e) e = native ? zpx_batch_fetch_native(b, out, status) : zpx_batch_fetch_rgba(b, out, out_stride, status);
    zpx_batch_close(b);
    return e;
}

// Large batches are cut into chunks that flow through two sets of device buffers and streams
// (context + shadow contexts, one host thread each): while one chunk's RGBA travels back over PCIe,
// the next chunk is parsed, uploaded and decoded.  Images are independent, so chunking changes
// nothing in the results.
static int32_t decode_batch_pipelined(zpx_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, int32_t n,
                                      uint8_t* const* out, const size_t* out_stride, int32_t* status, bool native) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !lens || !out))) return ZPX_E_INVALID_ARG;
    int32_t chunk = (int32_t)ctx->opt_pipeline_chunk;
    if (chunk == 0) {
        // auto: about 28 MB of compressed input per chunk (64 images of cfg2; roughly 0.5 GB of RGBA), which keeps
        // the device->host copy of one chunk around 10 ms -- long enough to amortise launches, short enough
        // for the pipeline to fill quickly
        uint64_t total = 0;
        for (int32_t i = 0; i < n; i++) total += lens[i];
        const uint64_t avg = std::max<uint64_t>(1, total / (uint64_t)std::max(n, 1));
        chunk = (int32_t)std::min<uint64_t>(2048, std::max<uint64_t>(16, (28u << 20) / avg));
    }
    if (ctx->opt_pipeline_chunk < 0 || n < 2 * chunk) return decode_range_rgba(ctx, bufs, lens, n, out, out_stride, status, native);
    const int n_workers = (int)std::min<int64_t>(ctx->opt_pipeline_workers, (n + chunk - 1) / chunk);
    while ((int)ctx->shadows.size() < n_workers - 1) {
        std::vector<int32_t> ids;
        for (const DeviceCtx& d : ctx->devs) ids.push_back(d.dev);
        zpx_ctx* sh = nullptr;
        int e = zpx_ctx_create(ids.data(), (int32_t)ids.size(), &sh);
        if (e) return e;
        ctx->shadows.push_back(sh);
    }
    for (zpx_ctx* sh : ctx->shadows) {
        sh->opt_entropy_mode = ctx->opt_entropy_mode;
        sh->opt_force_generic = ctx->opt_force_generic;
        sh->opt_subseq = ctx->opt_subseq;
        sh->opt_native = ctx->opt_native;
        sh->opt_progressive_mode = ctx->opt_progressive_mode;
        sh->opt_k2_dense = ctx->opt_k2_dense;
        sh->opt_gated_sweeps = ctx->opt_gated_sweeps;
    }
    // chunk list: the first two chunks are a quarter and a half of the regular size, so that the first
    // device->host copy starts early (the pipeline is bound by that copy; its fill time is pure loss)
    std::vector<std::pair<int32_t, int32_t>> chunks;
    for (int32_t i0 = 0; i0 < n;) {
        int32_t cnt = chunk;
        if (ctx->opt_pipeline_ramp && chunks.size() == 0) cnt = std::max(1, chunk / 4);
        else if (ctx->opt_pipeline_ramp && chunks.size() == 1) cnt = std::max(1, chunk / 2);
        cnt = std::min(cnt, n - i0);
        chunks.push_back({i0, cnt});
        i0 += cnt;
    }
    const int32_t n_chunks = (int32_t)chunks.size();
    std::atomic<int32_t> next(0);
    std::vector<int32_t> rc((size_t)n_workers, 0);
    std::vector<int32_t> st(status ? 0 : n);
    int32_t* stp = status ? status : st.data();
    const size_t par_each = std::max<size_t>(1, usable_cpus() / (size_t)n_workers);
    auto worker = [&](int w) {
        zpx_ctx* c = w == 0 ? ctx : ctx->shadows[(size_t)w - 1];
        t_par_limit = par_each;  // the workers' parse / staging loops share the host cores
        for (;;) {
            const int32_t k = next.fetch_add(1);
            if (k >= n_chunks) break;
            const int32_t i0 = chunks[k].first, cnt = chunks[k].second;
            const int e = decode_range_rgba(c, bufs + i0, lens + i0, cnt, out + i0, out_stride ? out_stride + i0 : nullptr, stp + i0, native);
            if (e) {
                // the chunk's images did not get a result: say so per image, then stop this worker
                for (int32_t i = i0; i < i0 + cnt; i++) stp[i] = e;
                rc[(size_t)w] = e;
                break;
            }
        }
    };
    std::vector<std::thread> threads;
    for (int w = 1; w < n_workers; w++) threads.emplace_back(worker, w);
    worker(0);
    t_par_limit = 0;
    for (std::thread& t : threads) t.join();
    int32_t ret = rc[0];
    for (int w = 1; w < n_workers; w++) {
        zpx_ctx* c = ctx->shadows[(size_t)w - 1];
        ctx->launches += c->launches.exchange(0);
        if (!ret && rc[(size_t)w]) {  // surface the CUDA error text on the caller's context
            ret = rc[(size_t)w];
            ctx->last_cuda = c->last_cuda;
            ctx->last_cuda_str = c->last_cuda_str;
        }
    }
    return ret;
}

int32_t zpx_decode_batch_rgba(zpx_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, int32_t n,
                              uint8_t* const* out, const size_t* out_stride, int32_t* status) {
    return decode_batch_pipelined(ctx, bufs, lens, n, out, out_stride, status, false);
}

// The same pipeline with jpeg.load's own return value as the result: the native Image variant's .pixels buffer
// (planar, MCU-padded Y/Cb/Cr or Gray: 1.5 bytes per pixel for 4:2:0 instead of RGBA's 4).  The fused kernel writes
// planes only for the duration of the call.
int32_t zpx_decode_batch_native(zpx_ctx* ctx, const uint8_t* const* bufs, const size_t* lens, int32_t n,
                                uint8_t* const* out, int32_t* status) {
    if (!ctx) return ZPX_E_INVALID_ARG;
    const int64_t saved = ctx->opt_native;
    ctx->opt_native = 2;
    const int32_t e = decode_batch_pipelined(ctx, bufs, lens, n, out, nullptr, status, true);
    ctx->opt_native = saved;
    return e;
}

}  // extern "C"
