// zpx_k0.cu -- byte-stuffing removal for the sequential entropy kernels.
//
// Replaces readByteStuffedByte / unreadByteStuffedByte (src/jpeg/decoder.zig:712-749, 479-487) for SOF0/SOF1
// scans: inside a restart interval the host's limit search guarantees that every 0xFF is followed by 0x00
// (zpx_parse.cpp find_limit), so removing the stuffing means dropping each 0x00 whose predecessor is 0xFF.
// Doing it once, in a streaming pass over the uploaded bytes, takes the FF 00 test and the byte-granular
// position bookkeeping out of every Huffman symbol step of K1 (zpx_k1_common.cuh RingReader) and makes bit
// positions of the unstuffed stream canonical for the self-synchronising decoder.
//
// Work unit: one warp per piece (ZpxSegDev: about 16 KB of raw bytes of one interval, cut by the host where no
// FF 00 pair is split; the host also knows how many pairs precede the piece, i.e. where its output starts).
// A round moves 512 raw bytes: coalesced 16-byte loads into shared memory, then 4 steps in which lane l looks
// at word 32 k + l and the word before it (conflict-free; the previous round's last word sits in the word in front
// of the staging area).  A byte is dropped iff (byte | ~predecessor) == 0 -- one exact zero-byte test per word,
// byte-parallel.  Three steps out of four have nothing to drop (a warp vote) and store their four bytes per lane
// straight; otherwise ballots + popc give each lane its output position.  The kept bytes go to a LINEAR buffer:
// complete 16-byte vectors are stored to HBM after every round and the < 16 bytes left over move to its start, so a
// step's byte stores share one address register and there is no ring mask.  Whether a step lies entirely inside the
// piece is decided per round for interior rounds and per 128-byte step otherwise (a 6 KB interval has two boundary
// steps, not eight).  Head and tail bytes that share a vector with the neighbouring piece are stored one by one.
// Reads and writes every entropy-coded byte once.  (The first form of this kernel -- 1 KB output ring, two
// zero-byte tests per word, per-round boundary handling -- took 0.32 ms on cfg2, this one 0.24 ms; 0.14 ms is the
// traffic at the measured HBM peak.)  tools/k0_model.py is a lane-by-lane Python model of the index logic, checked
// against a plain unstuffing on random pieces (tests/test_host.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "zpx_internal.h"
#include "zpx_kernels.h"

namespace zpx {

#ifndef ZPX_K0_PREFETCH
#define ZPX_K0_PREFETCH 1
#endif
#ifndef ZPX_K0_WARPS
#define ZPX_K0_WARPS 4
#endif

constexpr int K0_WARPS = ZPX_K0_WARPS;

__global__ void __launch_bounds__(K0_WARPS * 32) k0_unstuff(const uint8_t* __restrict__ blob, uint8_t* __restrict__ ublob,
                                                             const ZpxSegDev* __restrict__ segs, const int n_segs) {
    __shared__ __align__(16) uint8_t s_in[K0_WARPS][16 + 512];
    __shared__ __align__(16) uint8_t s_out[K0_WARPS][16 + 512 + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = blockIdx.x * K0_WARPS + warp;
    if (gid >= n_segs) return;  // whole warps
    const ZpxSegDev sg = segs[gid];
    uint8_t* in = s_in[warp] + 16;  // in[-4 .. -1]: last raw word of the round before
    uint8_t* out = s_out[warp];
    const uint64_t a0 = sg.src & ~(uint64_t)15;
    const uint32_t m = (uint32_t)(sg.dst & 15u);
    uint8_t* gptr = ublob + (sg.dst - m);  // 16-byte aligned; out[p] <-> gptr[p]
    uint32_t tot = m;                      // bytes in `out` (the first m of the first vector are the neighbour's)
    bool head_shared = m != 0;             // vector 0 of `out` is still the one shared with the piece before
    const int len = (int)sg.len;
    const uint8_t* gin = blob + a0 + 16u * (uint32_t)lane;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t* my_in = reinterpret_cast<const uint32_t*>(in) + lane;
    uint8_t* my_out = out + 4u * (uint32_t)lane;
    uint32_t carry = 0;  // lane 31: last raw word of the round before

    // one step: 128 raw bytes, lane l looks at word 32 k + l (conflict-free) and the word before it
    auto step = [&](auto full, const int k, const int rel0) {
        const uint32_t x = my_in[32 * k];
        const uint32_t pw = my_in[32 * k - 1];
        const uint32_t t = x | ~__funnelshift_l(pw, x, 8);  // byte i = byte i of x | ~(the byte before it)
        // 0x80 in every byte of t that is zero: a 0x00 whose predecessor is 0xFF (exact per byte, no carries)
        const uint32_t drop = ~((((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) | 0x7f7f7f7fu);
        const int rs = rel0 + 128 * k;  // (offset of the step's first byte) - src
        if (decltype(full)::value || (rs >= 1 && rs + 128 <= len)) {
            uint8_t* o = my_out + tot;
            if (!__any_sync(0xffffffffu, drop != 0)) {
                // nothing to remove in these 128 bytes (three steps out of four)
                o[0] = (uint8_t)x;
                o[1] = (uint8_t)(x >> 8);
                o[2] = (uint8_t)(x >> 16);
                o[3] = (uint8_t)(x >> 24);
                tot += 128u;
            } else {
                // some words of these 128 bytes lose one or two bytes
                const int ngone = __popc(drop);
                const uint32_t b1 = __ballot_sync(0xffffffffu, ngone >= 1), b2 = __ballot_sync(0xffffffffu, ngone >= 2);
                o -= __popc(b1 & lt) + __popc(b2 & lt);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (!(drop & (0x80u << (8 * i)))) *o++ = (uint8_t)(x >> (8 * i));
                tot += 128u - (uint32_t)(__popc(b1) + __popc(b2));
            }
        } else if (rs < len) {
            // a step at the piece's first or last byte: bytes outside [0, len) are not kept either; the piece's
            // first byte is never a stuffed 0x00 (its predecessor is not part of the piece)
            const int rel = rs + 4 * lane;
            uint32_t gone = drop;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (rel + i == 0) gone &= ~(0x80u << (8 * i));
                if (rel + i < 0 || rel + i >= len) gone |= 0x80u << (8 * i);
            }
            const int ngone = __popc(gone);
            // kept bytes before this lane's word = 4 * lane - bytes gone in the lanes below
            const uint32_t b1 = __ballot_sync(0xffffffffu, ngone >= 1), b2 = __ballot_sync(0xffffffffu, ngone >= 2);
            const uint32_t b3 = __ballot_sync(0xffffffffu, ngone >= 3), b4 = __ballot_sync(0xffffffffu, ngone >= 4);
            uint8_t* o = my_out + tot - (__popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt) + __popc(b4 & lt));
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (!(gone & (0x80u << (8 * i)))) *o++ = (uint8_t)(x >> (8 * i));
            tot += 128u - (uint32_t)(__popc(b1) + __popc(b2) + __popc(b3) + __popc(b4));
        }
        // (else: a step after the piece's last byte)
    };

    // the next round's 16 bytes are loaded while this round is worked on (one round of latency hidden per warp)
    uint4 vn = make_uint4(0u, 0u, 0u, 0u);
    if (ZPX_K0_PREFETCH && -(int)(sg.src - a0) + 16 * lane < len) vn = __ldg(reinterpret_cast<const uint4*>(gin));
    for (int rel0 = -(int)(sg.src - a0); rel0 < max(len, 1); rel0 += 512, gin += 512) {
        uint4 v = vn;
        if (ZPX_K0_PREFETCH) {
            vn = make_uint4(0u, 0u, 0u, 0u);
            if (rel0 + 512 + 16 * lane < len) vn = __ldg(reinterpret_cast<const uint4*>(gin + 512));
        } else if (rel0 + 16 * lane < len) {
            v = __ldg(reinterpret_cast<const uint4*>(gin));
        }
        reinterpret_cast<uint4*>(in)[lane] = v;
        if (lane == 31) reinterpret_cast<uint32_t*>(in)[-1] = carry;
        carry = v.w;
        __syncwarp();
        // a round whose 512 bytes are all inside the piece (and not its first byte) needs no per-step range test
        if (rel0 >= 1 && rel0 + 512 <= len) {
#pragma unroll
            for (int k = 0; k < 4; k++) step(std::true_type{}, k, rel0);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) step(std::false_type{}, k, rel0);
        }
        if (rel0 + 512 >= len && (sg.flags & 1u)) {
            // last round of the last piece of its interval: zeros up to the next 16-byte boundary (K1 reads whole
            // 16-byte chunks)
            const uint32_t padn = (16u - (tot & 15u)) & 15u;
            __syncwarp();
            if ((uint32_t)lane < padn) out[tot + lane] = 0;
            tot += padn;
        }
        __syncwarp();
        // complete vectors to HBM (at most 33: 15 + 512 + 15 bytes), the rest to the start of the buffer
        const uint32_t nvec = tot >> 4, rest = tot & 15u;
        if ((uint32_t)lane < nvec && !(head_shared && lane == 0))
            *reinterpret_cast<uint4*>(gptr + 16 * lane) = *reinterpret_cast<const uint4*>(out + 16 * lane);
        if (nvec > 32u && lane == 0) *reinterpret_cast<uint4*>(gptr + 512) = *reinterpret_cast<const uint4*>(out + 512);
        if (head_shared && nvec > 0u) {
            // the first vector is shared with the piece before this one: only bytes m..15 are ours
            if ((uint32_t)lane >= m && lane < 16) gptr[lane] = out[lane];
            head_shared = false;
        }
        uint8_t keep = 0;
        if ((uint32_t)lane < rest) keep = out[16u * nvec + lane];
        __syncwarp();
        if ((uint32_t)lane < rest && nvec > 0u) out[lane] = keep;
        gptr += 16u * nvec;
        tot = rest;
        __syncwarp();
    }
    // bytes of the last, incomplete vector (the piece after this one owns the rest of it)
    if ((uint32_t)lane < tot && !(head_shared && (uint32_t)lane < m)) gptr[lane] = out[lane];
}

cudaError_t k0_launch_unstuff(const uint8_t* blob, uint8_t* ublob, const ZpxSegDev* segs, int n_segs, cudaStream_t s) {
    if (n_segs <= 0) return cudaSuccess;
    k0_unstuff<<<(n_segs + K0_WARPS - 1) / K0_WARPS, K0_WARPS * 32, 0, s>>>(blob, ublob, segs, n_segs);
    return cudaGetLastError();
}

}  // namespace zpx
