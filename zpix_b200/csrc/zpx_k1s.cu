// zpx_k1s.cu -- self-synchronising speculative Huffman decode for entropy-coded segments that have
// no (or too few) restart markers.  Same reference code as zpx_k1.cu (processSos MCU loop,
// src/jpeg/decoder.zig:1294-1452, and the bit/Huffman layer :909-1134), parallelised INSIDE a segment.
//
// A segment ("domain": a restart interval, or the whole scan when DRI = 0) is cut into sub-sequences
// of sub_bytes raw bytes; one lane per sub-sequence, 32 consecutive sub-sequences per warp.
// Decoder state at a symbol boundary: (raw bit position, block index inside the MCU, zig-zag index).
//
//   k1s_sync   every lane decodes its sub-sequence (no output) from a start state and records the
//              state in which the decoder crosses into the next sub-sequence, the number of blocks it
//              started and the sum of the DC differences it decoded, per component.  Sub-sequence 0
//              starts in the true state, the others from a guess (boundary byte, block 0, DC).  Lanes
//              hand their end state to the next lane (shuffle inside the warp, global memory across
//              warps) and re-decode while their start state changes.  After sweep 0 only the first
//              sub-sequence of each warp can be stale: k1s_fix repairs those, one lane per warp boundary,
//              and another full sweep runs only if a repaired sub-sequence ends in a new state.
//              Fixed point => every start state is the true one, by induction from sub-sequence 0
//              (F_i(true start of i) = true start of i+1); Huffman streams re-synchronise after a
//              few symbols, so this takes 2-3 rounds in practice but is exact for any input.
//   k1s_scan   exclusive prefix sums, per domain, of the block counts and DC sums (wrapping int32
//              adds: identical to the sequential accumulation of decoder.zig:1374).
//   k1s_write  (zpx_k1.cu) every lane decodes again from its true start state and writes the blocks that
//              START inside its sub-sequence (it skips the tail of a block begun earlier and runs past
//              its boundary to finish its last block), absolute DC included.
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpx_k1_common.cuh"

namespace zpx {

constexpr int K1S_NT = 128;     // threads per CTA (4 warps), one stream per thread
constexpr int K1S_RS = K1S_NT * 4;  // ring word stride

__device__ __forceinline__ unsigned long long pack_state(uint32_t pos, int c, int z) {
    return (unsigned long long)pos | ((unsigned long long)(uint32_t)c << 32) | ((unsigned long long)(uint32_t)z << 40);
}

struct LaneCtx {
    const ZpxIntervalDev* iv;
    const ZpxScanDev* sc;
    uint32_t li;   // sub-sequence index inside the domain
    uint32_t bnd;  // byte offset (unstuffed, from the domain's start) of this sub-sequence's end boundary
};

__device__ __forceinline__ void lane_setup(const K1Params& P, const ZpxWarpDev w, int lane, LaneCtx& L) {
    L.iv = &P.ivs[w.iv];
    L.sc = &P.scans[L.iv->scan];
    L.li = w.first + lane;
    L.bnd = (L.li + 1) * L.iv->sub_bytes;
}

// Decode (without output) from state `in` until the first symbol that starts at or after the
// sub-sequence boundary, or the data runs out.  Only the bit count and the zig-zag advance of each
// symbol matter here (one 32-bit ZPX_FE entry); values are extracted for DC symbols only,
// to accumulate the per-component sums of DC differences.  Invalid codes advance one bit (any
// deterministic rule works: true states of a well-formed stream never meet one; errors are reported
// by k1s_write).  Called by the lanes in `mask` together; the vote keeps them in lock step.
//
// Checkpoints (CK): a re-decode from a corrected start state joins the trajectory of the lane's previous
// decode after a few dozen symbols (that is what self-synchronisation means), and from there on repeats
// it.  Each decode therefore records its state at the first symbol boundary at or after 1/4, 1/2 and 3/4
// of the sub-sequence -- (bits past that byte offset, block phase, zig-zag index), which does not
// depend on where the decode started -- together with the counts so far; a later decode that reaches a
// checkpoint in the same state stops there and completes its counts from the previous decode's
// (exact: equal state at equal position means identical continuation).  n_out / dc_out / old_out
// carry the previous decode's results in.
struct CkStore {          // this lane's slots in shared memory, [3] strided by K1S_NT
    uint32_t* st;         // rel | c << 16 | k << 24 ; 0xffffffff = none
    int* n;
    int4* dc;
};

template <bool CK, bool SMEM>
__device__ __forceinline__ unsigned long long sync_decode(const K1Params& P, const LaneCtx& L, unsigned long long in,
                                                          unsigned mask, int& n_out, int4& dc_out, int& bad_out,
                                                          const uint32_t sdesc, const uint32_t tb, const uint32_t ring_col,
                                                          unsigned long long old_out = 0, CkStore ck = CkStore{nullptr, nullptr, nullptr}) {
    const ZpxScanDev* __restrict__ sc = L.sc;
    // thresholds T_m = bnd - (3 - m) * step, m = 0..2, then the sub-sequence boundary itself (m == 3)
    const uint32_t step = (L.iv->sub_bytes >> 2) & ~3u;
    const bool use_ck = CK && sc->rotate == 0 && step >= 32;
    uint32_t m = 3;
    if (use_ck) {
        // arm the first threshold that lies after the start position
        m = 0;
        while (m < 3 && (L.bnd - (3 - m) * step) * 8u <= (uint32_t)in) {
            ck.st[m * K1S_NT] = 0xffffffffu;
            m++;
        }
    }
    uint32_t thr = (L.bnd - (3 - m) * step) * 8u;  // m == 3: the boundary
    RingReader<K1S_RS> rd;
    rd.init(ring_col, P.ublob, L.iv->ustart, L.iv->ulen, (uint32_t)in);
    int c = (int)((in >> 32) & 0xff), k = (int)((in >> 40) & 0xff);
    const int nblk = sc->interleaved ? sc->nblk : 1;
    const uint4* __restrict__ bpack = reinterpret_cast<const uint4*>(sc->blk_pack);
    // rotate mode: the phase is not part of the state; c counts blocks relative to this sub-sequence and
    // the DC sums are kept per relative phase (k1s_scan rotates them into components)
    const bool rotate = sc->rotate != 0;
    if (rotate) c = 0;
    int n = 0, d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    uint4 bi = SMEM ? lds_u128(sdesc + c * 16) : bpack[c];
    const uint32_t* __restrict__ fdc = SMEM ? nullptr : P.huff[bi.x].fast;
    const uint32_t* __restrict__ fac = SMEM ? nullptr : P.huff[bi.y].fast;
    bool go = true, merged = false, skipping = false;
    int bad = 0;  // an invalid code was skipped (one bit further).  If this decode started in the true state, the
                  // reference fails there (diagnostic only: see `skipping` below for how the write pass finds it)
    int it = 0;
    while (__any_sync(mask, go)) {
        if (it == K1_TOPUP) {  // uniform over the lanes in mask
            rd.topup();
            it = 0;
        }
        it++;
        if (go) {
            if (rd.bitpos >= rd.endbits) {
                go = false;  // the data ran out
            } else if (rd.bitpos >= thr) {
                if (!CK || m >= 3) {
                    go = false;  // first symbol boundary at or after the end of the sub-sequence
                } else {
                    const uint32_t cur = (rd.bitpos - thr) | (uint32_t)c << 16 | (uint32_t)k << 24;
                    const int4 odc = ck.dc[m * K1S_NT];
                    const int on = ck.n[m * K1S_NT];
                    if (ck.st[m * K1S_NT] == cur) {
                        // joined the previous trajectory: the rest is known
                        const int dn = n - on;
                        const int e0 = d0 - odc.x, e1 = d1 - odc.y, e2 = d2 - odc.z, e3 = d3 - odc.w;
                        for (uint32_t q = m; q < 3; q++) {  // re-base the stored counts on this decode's start
                            ck.n[q * K1S_NT] += dn;
                            int4 t = ck.dc[q * K1S_NT];
                            t.x += e0; t.y += e1; t.z += e2; t.w += e3;
                            ck.dc[q * K1S_NT] = t;
                        }
                        n = n_out + dn;
                        d0 = dc_out.x + e0;
                        d1 = dc_out.y + e1;
                        d2 = dc_out.z + e2;
                        d3 = dc_out.w + e3;
                        merged = true;
                        go = false;
                    } else {
                        ck.st[m * K1S_NT] = cur;
                        ck.n[m * K1S_NT] = n;
                        ck.dc[m * K1S_NT] = make_int4(d0, d1, d2, d3);
                        m++;
                        thr = (L.bnd - (3 - m) * step) * 8u;
                    }
                }
            } else {
                const uint32_t hi = rd.peek();
                const bool isdc = k == 0;
                uint32_t e;
                if (SMEM) e = lds_u32(isdc ? bi.x + ((hi >> (32 - K1_DLB)) << 2) : bi.y + ((hi >> (32 - K1_ALB)) << 2));
                else e = __ldg((isdc ? fdc : fac) + (hi >> (32 - ZPX_LUT_BITS)));
                if (SMEM && e == 0 && !isdc) e = k1_long_ac_sm(tb, (bi.w >> 24) & 15u, hi);  // the common rare case, inline
                if ((int)e <= 0) {
                    // longer code / invalid code / EOB run / DC category > 16
                    unsigned long long r;
                    if (SMEM) r = k1_slow_symbol_sm(tb, isdc ? (bi.w >> 20) & 15u : (bi.w >> 24) & 15u, hi, isdc, e);
                    else r = k1_slow_symbol(&P.huff[isdc ? bi.x : bi.y], hi, isdc, e);
                    e = (uint32_t)r;
                    if ((int)(r >> 32) == ZPX_E_BadHuffmanCode) {
                        bad = 1;
                        e = 1u | 32u << 16;  // invalid: one bit further, same state (tot = 1, no value bits, adv = 0)
                        // An invalid code where a block should start: the block that fails to start is counted (once
                        // per run of skipped bits), so that the lane that reaches it in the true state decodes it in
                        // the write pass -- and reports the reference's error there -- and every block found
                        // further on gets a strictly larger ordinal: nothing decoded after the first error of the
                        // reference's order can tie with it in the status key.
                        if (k == 0 && !skipping) {
                            skipping = true;
                            n++;
                        }
                    } else if (e >> 31) {
                        // EOB run: r more bits belong to the symbol
                        const uint32_t len = (e >> 8) & 0xffu, rr = (e >> 16) & 0xffu;
                        e = ZPX_FE(len + rr, len, 0, 64, 0);
                    }
                }
                const int len = fe_len(e), s32 = fe_s32(e);
                int tot = fe_tot(e);
                const int adv = fe_adv(e);
                if (adv) skipping = false;  // a valid symbol
                if (isdc && adv) {
                    const int v = fe_extend(hi << len, s32);
                    const int comp = rotate ? c : (int)(bi.z & 0xff);
                    if (comp == 0) d0 += v; else if (comp == 1) d1 += v; else if (comp == 2) d2 += v; else d3 += v;
                    n++;
                } else if (k + adv - 1 > 63 && s32 != 32) {
                    tot = len;  // decoder.zig:1393-1395: run past the block end, the value bits stay unread
                }
                k += adv;
                rd.bitpos += tot;
                if (k > 63) {
                    k = 0;
                    c = c + 1 == nblk ? 0 : c + 1;
                    if (SMEM) {
                        bi = lds_u128(sdesc + c * 16);
                    } else {
                        bi = bpack[c];
                        fdc = P.huff[bi.x].fast;
                        fac = P.huff[bi.y].fast;
                    }
                }
            }
        }
    }
    n_out = n;
    dc_out = make_int4(d0, d1, d2, d3);
    bad_out = (CK && merged) ? (bad_out | bad) : bad;  // merged: the rest of the previous decode still counts
    if (CK && merged) return old_out;
    if (CK && use_ck)  // checkpoints this decode did not reach (data ran out) must not match later
        for (uint32_t q = m; q < 3; q++) ck.st[q * K1S_NT] = 0xffffffffu;
    // a DC symbol decoded in rotate mode belongs to relative phase c *before* the block ends; when the
    // sub-sequence ends inside a block, that block's phase is (c) and was counted -- nothing to fix
    return pack_state(rd.bitpos, rotate ? 0 : c, k);
}

// ---------------------------------------------------------------------------
// sweep kernel
// ---------------------------------------------------------------------------
struct K1SyncSmem {
    uint32_t ring[K1_RW * K1S_NT];
    uint32_t ck[3 * K1S_NT];
    int cn[3 * K1S_NT];
    int4 cdc[3 * K1S_NT];
    K1Tables tab;
};

__global__ void __launch_bounds__(K1S_NT) k1s_sync(const K1SParams P, const int sweep) {
    if (P.gate != nullptr && *P.gate == 0) return;  // (enqueued ahead of time: the previous pass changed nothing)
    extern __shared__ __align__(16) uint8_t k1s_smem[];
    K1SyncSmem& S = *reinterpret_cast<K1SyncSmem*>(k1s_smem);
    const CkStore ck{S.ck + threadIdx.x, S.cn + threadIdx.x, S.cdc + threadIdx.x};
    for (int q = 0; q < 3; q++) S.ck[q * K1S_NT + threadIdx.x] = 0xffffffffu;
    const int wid = blockIdx.x * (K1S_NT / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const bool wv = wid < P.n_warps;
    const ZpxWarpDev w = P.warps[wv ? wid : P.n_warps - 1];
    LaneCtx L;
    lane_setup(P.k1, w, lane, L);
    S.tab.lane_scan[threadIdx.x] = L.iv->scan;
    __syncthreads();
    const bool cached = k1_tables_setup(P.k1, S.tab, K1S_NT);
    if (!wv) return;
    uint32_t sdesc = 0;
    if (cached)
        for (int i = 0; i < S.tab.nscan; i++)
            if (S.tab.scan_id[i] == L.iv->scan) sdesc = smem_addr(&S.tab.desc[i][0]);
    const uint32_t tb = smem_addr(&S.tab);
    const uint32_t rc = smem_addr(S.ring) + threadIdx.x * 4;

    const bool active = L.li < L.iv->nsub;
    const uint32_t t = L.iv->sub_first + L.li;
    const bool last_active = active && (lane == 31 || L.li + 1 == L.iv->nsub);
    const bool more_warps = w.first + 32 < L.iv->nsub;

    unsigned long long in = 0, out = 0, old_out = 0;
    int n = 0, bad = 0;
    int4 dc = make_int4(0, 0, 0, 0);
    bool dirty = false;
    if (sweep == 0) {
        if (active) {
            in = pack_state(L.li * L.iv->sub_bytes * 8u, 0, 0);
            dirty = true;
        }
    } else {
        if (active) {
            in = P.s_in[t];
            out = P.s_out[t];
            old_out = out;
            n = P.s_n[t];
            dc = P.s_dc[t];
            bad = P.s_bad[t];
            if (lane == 0 && w.first > 0) {
                const unsigned long long ni = P.s_out[t - 1];
                if (ni != in) {
                    in = ni;
                    dirty = true;
                }
            }
        }
        // a lane whose predecessor's end state is not its start state (k1s_fix may have changed lane 0's)
        const unsigned long long po0 = __shfl_up_sync(0xffffffffu, out, 1);
        if (lane > 0 && active && po0 != in) {
            in = po0;
            dirty = true;
        }
        if (!__any_sync(0xffffffffu, dirty)) return;
    }
    bool touched = false;
    do {
        const unsigned dm = __ballot_sync(0xffffffffu, dirty);
        if (dirty) {
            if (cached) out = sync_decode<true, true>(P.k1, L, in, dm, n, dc, bad, sdesc, tb, rc, out, ck);
            else out = sync_decode<true, false>(P.k1, L, in, dm, n, dc, bad, 0, tb, rc, out, ck);
            touched = true;
        }
        const unsigned long long po = __shfl_up_sync(0xffffffffu, out, 1);
        bool nd = false;
        if (lane > 0 && active && po != in) {
            in = po;
            nd = true;
        }
        dirty = nd;
    } while (__any_sync(0xffffffffu, dirty));
    if (active && touched) {
        P.s_in[t] = in;
        P.s_out[t] = out;
        P.s_n[t] = n;
        P.s_dc[t] = dc;
        P.s_bad[t] = bad;
    }
    if (last_active && more_warps && (sweep == 0 || out != old_out)) atomicExch(P.changed, 1);
}

// ---------------------------------------------------------------------------
// boundary kernel: after a sweep only the FIRST sub-sequence of every warp can still have a stale start
// state (it comes from the previous warp).  One lane per warp boundary re-decodes that sub-sequence from
// the previous warp's end state; the lanes of a warp here are 32 different boundaries, so this costs
// 1/32 of the warp instructions of another full sweep.  If such a re-decode ends in a new state, the rest
// of that warp is stale: `changed` asks the host for another k1s_sync sweep.  The lanes of a CTA belong to
// up to 128 different scans here: tables are read in HBM.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(K1S_NT) k1s_fix(const K1SParams P) {
    if (P.gate != nullptr && *P.gate == 0) return;
    __shared__ uint32_t s_ring[K1_RW * K1S_NT];
    const int g = blockIdx.x * K1S_NT + threadIdx.x;
    const bool valid = g < P.n_warps;
    const ZpxWarpDev w = P.warps[valid ? g : 0];
    LaneCtx L;
    lane_setup(P.k1, w, 0, L);
    bool dirty = valid && w.first > 0;
    const uint32_t t = L.iv->sub_first + L.li;
    unsigned long long in = 0, old_out = 0;
    if (dirty) {
        const unsigned long long ni = P.s_out[t - 1];
        in = P.s_in[t];
        old_out = P.s_out[t];
        dirty = ni != in;
        in = ni;
    }
    const unsigned dm = __ballot_sync(0xffffffffu, dirty);
    if (dm == 0) return;
    if (dirty) {
        int n = 0, bad = 0;
        int4 dc = make_int4(0, 0, 0, 0);
        const unsigned long long out = sync_decode<false, false>(P.k1, L, in, dm, n, dc, bad, 0, 0, smem_addr(s_ring) + threadIdx.x * 4);
        P.s_in[t] = in;
        P.s_out[t] = out;
        P.s_n[t] = n;
        P.s_dc[t] = dc;
        P.s_bad[t] = bad;
        if (out != old_out && L.li + 1 < L.iv->nsub) atomicExch(P.changed, 1);
    }
}

// ---------------------------------------------------------------------------
// per-domain exclusive scan of (blocks started, DC sums); one warp per domain
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(K1S_NT) k1s_scan(const K1SParams P) {
    const int d = blockIdx.x * (K1S_NT / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (d >= P.n_iv) return;
    const ZpxIntervalDev* iv = &P.k1.ivs[d];
    const ZpxScanDev* sc = &P.k1.scans[iv->scan];
    const bool rotate = sc->rotate != 0;
    const int nblk = sc->nblk;
    const uint32_t nsub = iv->nsub, base = iv->sub_first;
    int cn = 0;
    int4 cd = make_int4(0, 0, 0, 0);
    for (uint32_t i0 = 0; i0 < nsub; i0 += 32) {
        const uint32_t i = i0 + lane;
        int n = 0;
        int4 v = make_int4(0, 0, 0, 0);
        int kstart = 0;
        if (i < nsub) {
            n = P.s_n[base + i];
            v = P.s_dc[base + i];
            kstart = (int)((P.s_in[base + i] >> 40) & 0xff);
        }
        int sn = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tn = __shfl_up_sync(0xffffffffu, sn, o);
            if (lane >= o) sn += tn;
        }
        const int excl_n = cn + sn - n;
        if (rotate) {
            // v holds the DC sums by phase relative to the sub-sequence (relative block 0 is the block in
            // flight at its start); the first block STARTED here is block excl_n of the segment
            const int t0 = kstart != 0 ? 1 : 0;
            const int rot = ((excl_n - t0) % nblk + nblk) % nblk;
            int r[4] = {0, 0, 0, 0};
            const int src[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (t < nblk) {
                    const int p = (t + rot) % nblk;
                    const int comp = sc->blk_comp[p];
                    if (comp == 0) r[0] += src[t]; else if (comp == 1) r[1] += src[t]; else if (comp == 2) r[2] += src[t]; else r[3] += src[t];
                }
            v = make_int4(r[0], r[1], r[2], r[3]);
        }
        int4 sv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tx = __shfl_up_sync(0xffffffffu, sv.x, o), ty = __shfl_up_sync(0xffffffffu, sv.y, o);
            const int tz = __shfl_up_sync(0xffffffffu, sv.z, o), tw = __shfl_up_sync(0xffffffffu, sv.w, o);
            if (lane >= o) {
                sv.x += tx;
                sv.y += ty;
                sv.z += tz;
                sv.w += tw;
            }
        }
        if (i < nsub) {
            P.s_n[base + i] = excl_n;
            P.s_dc[base + i] = make_int4(cd.x + sv.x - v.x, cd.y + sv.y - v.y, cd.z + sv.z - v.z, cd.w + sv.w - v.w);
        }
        cn += __shfl_sync(0xffffffffu, sn, 31);
        cd.x += __shfl_sync(0xffffffffu, sv.x, 31);
        cd.y += __shfl_sync(0xffffffffu, sv.y, 31);
        cd.z += __shfl_sync(0xffffffffu, sv.z, 31);
        cd.w += __shfl_sync(0xffffffffu, sv.w, 31);
    }
}

// the write pass (k1s_write) lives in zpx_k1.cu: it is the block-synchronous loop of the lane-per-interval
// kernel, started from each sub-sequence's true state

cudaError_t k1s_launch_sync(const K1SParams& P, int sweep, cudaStream_t s) {
    if (P.n_warps <= 0) return cudaSuccess;
    const int wpc = K1S_NT / 32;
    cudaError_t e = cudaFuncSetAttribute(k1s_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SyncSmem));
    if (e != cudaSuccess) return e;
    k1s_sync<<<(P.n_warps + wpc - 1) / wpc, K1S_NT, sizeof(K1SyncSmem), s>>>(P, sweep);
    return cudaGetLastError();
}
cudaError_t k1s_launch_fix(const K1SParams& P, cudaStream_t s) {
    if (P.n_warps <= 0) return cudaSuccess;
    k1s_fix<<<(P.n_warps + K1S_NT - 1) / K1S_NT, K1S_NT, 0, s>>>(P);
    return cudaGetLastError();
}
cudaError_t k1s_launch_scan(const K1SParams& P, cudaStream_t s) {
    if (P.n_iv <= 0) return cudaSuccess;
    const int wpc = K1S_NT / 32;
    k1s_scan<<<(P.n_iv + wpc - 1) / wpc, K1S_NT, 0, s>>>(P);
    return cudaGetLastError();
}
}  // namespace zpx
