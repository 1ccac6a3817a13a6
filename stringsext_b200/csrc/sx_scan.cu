// sx_scan.cu -- sm_100a kernels of the scanner hot path + the C ABI (include/stringsext_b200.h).
//
// sx_prefilter_kernel<FAMILY>  one streaming pass over the whole input (the HBM-bound kernel):
//     tiles of 256 windows staged into swizzled shared memory, one lane per window, SWAR block
//     classification -> 128-bit "good byte" mask -> INTERESTING flag; output: a bit mask of the
//     windows the exact kernel has to look at (interesting windows, their neighbours, tile edges).
// sx_list_scan_kernel / sx_list_expand_kernel  turn the per-tile masks into an ordered window list.
// sx_exact_kernel<Dec>  one list entry per thread: decoder + SplitStr automaton (sx_core.cuh),
//     transfer-function carry resolution along runs of adjacent windows, count / reserve (one
//     atomicAdd per block) / write of the finding records in stream order.
// sx_materialize_kernel  transcodes each record's input range to UTF-8 text.
//
// There is no CPU scanning path in this file: without a CUDA device every call fails.
#include "../../include/stringsext_b200.h"
#include "sx_exact.cuh"
#include "sx_sparse_utf8.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sx {

// ---------------------------------------------------------------------------------------------
// Prefilter: one streaming pass over the input (HBM bound).  256 windows per tile, one lane per
// window; each lane classifies its window's bytes with SWAR block tests, packs the "good byte"
// flags into a 128-bit mask and decides INTERESTING from the mask (sx_core.cuh, pref_*_ref is the
// byte-wise specification of exactly what is computed here).
// ---------------------------------------------------------------------------------------------
// Every prefilter CTA owns a contiguous range of tiles and compacts the windows it keeps into its own
// region of `list` (region of CTA b starts at b * region_stride); cta_count[b] = how many it kept.
struct PrefOut {
    uint32_t* list;        // indexed by tile: the region of a CTA starts at its first tile * 256
    uint32_t* cta_count;   // of this launch (piece)
    long long tiles_per_cta;
    long long tile0, tile_end;  // tiles of this launch: one piece of the call's window range
    long long w_first;          // first window of the piece: always listed (its carry-in is given, sx_range_carry_kernel)
    long long w_lo, w_hi;       // the call's window range: windows of the boundary tiles outside it are not this call's
    uint32_t piece_ctas;        // > 0: every piece_ctas-th CTA starts a piece of the exact stage; its first window is listed too
};

struct PrefK {
    // runtime block-function coefficients (0 or 0xFFFFFFFF): K[k] = block k may be good
    uint32_t ka[4];  // bytes < 0x80, blocks 0..3
    uint32_t kh[4];  // bytes >= 0x80, blocks 4..7 (PF_UTF8: lead blocks, only [2],[3] used)
    uint32_t multi;  // PF_UTF8
    uint32_t hi_mask[4];   // PF_UNIT: bit i of the 128-bit mask is the most significant byte of a unit
    uint32_t edge_mask[4]; // PF_UNIT: bytes of units whose tested byte lies outside the window
    uint32_t spread_left;  // PF_UNIT: spread the tested flag towards lower addresses (LE) or higher (BE)
};

#define LOP3_SEL 0xCA  // a ? b : c

__device__ __forceinline__ uint32_t lop3_sel(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// g(b6,b5) with b6 at bit 7 of x1 and b5 at bit 7 of x2; k0..k3 = value for (b6,b5) = 00,01,10,11
__device__ __forceinline__ uint32_t blk2(uint32_t x1, uint32_t x2, const uint32_t* k) {
    return lop3_sel(x1, lop3_sel(x2, k[3], k[2]), lop3_sel(x2, k[1], k[0]));
}
__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// a[0..3]: dp4a accumulators, each holding 8 byte flags at bits S..S+7 -> the 32 flags as one mask word.
// The multiply-adds run on the FMA pipe; one shift on the ALU pipe.
__device__ __forceinline__ uint32_t mad_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
template <int S>
__device__ __forceinline__ uint32_t pack_plane(const uint32_t* a) {
    const uint32_t t = mad_u(a[2], 65536u, mad_u(a[1], 256u, a[0]));
    return mad_u(a[3], 1u << (24 - S), t >> S);
}

// 128-bit helpers on uint32_t m[4] (bit i of the mask = byte i of the window)
__device__ __forceinline__ void shr128(const uint32_t* m, uint32_t s, uint32_t* o) {
    const uint32_t wsft = s >> 5, bs = s & 31;
    uint32_t t[8] = {m[0], m[1], m[2], m[3], 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // static indexing only
            if ((uint32_t)j == wsft) { lo = t[k + j]; hi = t[k + j + 1]; }
        }
        o[k] = __funnelshift_r(lo, hi, bs);
    }
}
__device__ __forceinline__ uint32_t ctz128(const uint32_t* m) {  // 128 when zero
    if (m[0]) return __ffs(m[0]) - 1;
    if (m[1]) return 32 + __ffs(m[1]) - 1;
    if (m[2]) return 64 + __ffs(m[2]) - 1;
    if (m[3]) return 96 + __ffs(m[3]) - 1;
    return 128;
}
__device__ __forceinline__ uint32_t clz128(const uint32_t* m) {
    if (m[3]) return __clz(m[3]);
    if (m[2]) return 32 + __clz(m[2]);
    if (m[1]) return 64 + __clz(m[1]);
    if (m[0]) return 96 + __clz(m[0]);
    return 128;
}
template <int S>
__device__ __forceinline__ void and_shr128_static(uint32_t* r) {  // r &= r >> S, S a power of two < 128
    if (S < 32) {
        r[0] &= __funnelshift_r(r[0], r[1], S);
        r[1] &= __funnelshift_r(r[1], r[2], S);
        r[2] &= __funnelshift_r(r[2], r[3], S);
        r[3] &= r[3] >> S;
    } else if (S == 32) {
        r[0] &= r[1]; r[1] &= r[2]; r[2] &= r[3]; r[3] = 0;
    } else {
        r[0] &= r[2]; r[1] &= r[3]; r[2] = 0; r[3] = 0;
    }
}
// r = positions where a run of >= T set bits starts (1 <= T <= 128, warp-uniform); returns r != 0
__device__ __forceinline__ bool has_run128(const uint32_t* m, uint32_t T, uint32_t* r) {
    r[0] = m[0]; r[1] = m[1]; r[2] = m[2]; r[3] = m[3];
    uint32_t have = 1;
    if (T >= 2) { and_shr128_static<1>(r); have = 2; }
    if (T >= 4) { and_shr128_static<2>(r); have = 4; }
    if (T >= 8) { and_shr128_static<4>(r); have = 8; }
    if (T >= 16) { and_shr128_static<8>(r); have = 16; }
    if (T >= 32) { and_shr128_static<16>(r); have = 32; }
    if (T >= 64) { and_shr128_static<32>(r); have = 64; }
    if (T >= 128) { and_shr128_static<64>(r); have = 128; }
    if (T > have) {
        uint32_t s[4];
        shr128(r, T - have, s);
        r[0] &= s[0]; r[1] &= s[1]; r[2] &= s[2]; r[3] &= s[3];
    }
    return (r[0] | r[1] | r[2] | r[3]) != 0;
}

// Refinement of a long good-byte run (PF_UTF8, rare): a run of >= T good bytes can hold n chars only if it
// has >= n non-continuation bytes.  Rebuilds the continuation-byte mask of the (full, 16-byte aligned) window at
// `wbytes` and walks its long runs.  m = good-byte mask of the window.
__device__ __noinline__ bool pref_refine_utf8(const uint8_t* wbytes, uint32_t nchunk, const uint32_t* m, uint32_t T,
                                              uint32_t n_chars) {
    uint32_t rs[4];
    if (!has_run128(m, T, rs)) return false;
    uint32_t cnm[4] = {0, 0, 0, 0};
    {
        uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if ((uint32_t)c < nchunk) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wbytes) + c);
                const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = (4 * c + j) & 7;
                    acc[k >> 1] = dp4a_u(xs[j] & ~(xs[j] << 1) & 0x80808080u, (k & 1) ? 0x80402010u : 0x08040201u, acc[k >> 1]);
                    if (k == 7) {
                        cnm[(4 * c + j) >> 3] = (acc[0] >> 7) | (acc[1] << 1) | (acc[2] << 9) | (acc[3] << 17);
                        acc[0] = acc[1] = acc[2] = acc[3] = 0;
                    }
                }
                if ((c & 1) == 0 && (uint32_t)c == nchunk - 1) cnm[c >> 1] = (acc[0] >> 7) | (acc[1] << 1);
            }
        }
    }
    const uint32_t nm2[4] = {~m[0], ~m[1], ~m[2], ~m[3]};
    const uint32_t mc[4] = {m[0] & ~cnm[0], m[1] & ~cnm[1], m[2] & ~cnm[2], m[3] & ~cnm[3]};
    bool ok = false;
    for (int guard = 0; guard < 16 && !ok && (rs[0] | rs[1] | rs[2] | rs[3]); ++guard) {
        const uint32_t st0 = ctz128(rs);  // start of a maximal run of >= T good bytes
        uint32_t t4[4];
        shr128(nm2, st0, t4);
        uint32_t rl = ctz128(t4);  // its length (the run may reach bit 127)
        if (rl > 128 - st0) rl = 128 - st0;
        shr128(mc, st0, t4);
        uint32_t chars = 0;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const int lo_bit = q4 * 32;
            uint32_t wv = t4[q4];
            if ((int)rl <= lo_bit) wv = 0;
            else if ((int)rl < lo_bit + 32) wv &= (1u << (rl - lo_bit)) - 1u;
            chars += __popc(wv);
        }
        ok = chars >= n_chars;
        const uint32_t e = st0 + rl;  // drop the run-start bits of this run
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const int lo_bit = q4 * 32;
            if ((int)e >= lo_bit + 32) rs[q4] = 0;
            else if ((int)e > lo_bit) rs[q4] &= ~((1u << (e - lo_bit)) - 1u);
        }
    }
    return ok;
}

// TMA helpers (cp.async.bulk.tensor + mbarrier), sm_90+/sm_100a PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

constexpr uint32_t kPrefStageBytes = 32768;
constexpr uint32_t kPrefQueueCap = 448;  // flushed as soon as fewer than one tile's worth of slots is free
constexpr size_t kPrefSmemBytes = 2 * kPrefStageBytes + 1024 + 64 + 64 + 32 + kPrefQueueCap * 20 + 256;

// DEFSHAPE (PF_UTF8 only): the default filter shape -- ASCII blocks 1..3 may pass, only 2-byte leads
// (block 6) may pass -- with the block functions folded into single LOP3s.
// FAST (W == 128, T <= 32): bit-plane classification.  The top three bits of every byte are gathered into three
// 128-bit planes with AND + dp4a (the dp4a and the plane assembly run on the FMA pipe, which the old per-word
// SWAR left idle while the ALU pipe was saturated: ncu, profiles/r01_final_ncu_summary.json), the class logic then
// costs a few LOP3 per 32 bytes, and the run test works on the window's good-byte mask extended by the last
// T - 1 flags of the previous window (no lead/trail exchange).  Everything else takes the per-word path.
template <int FAMILY, bool DEFSHAPE, bool FAST>
__global__ void __launch_bounds__(kPrefThreads, 3)
sx_prefilter_kernel(const __grid_constant__ ScanParams P, const PrefCfg C, const __grid_constant__ PrefK K, const PrefOut O,
                    long long total_windows, const __grid_constant__ CUtensorMap tmap, uint32_t use_tma) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // two 32 KiB stages (TMA SWIZZLE_128B layout == swz()), then the per-tile exchange arrays and the mbarriers
    uint32_t* s_trail = reinterpret_cast<uint32_t*>(smem_raw + 2 * kPrefStageBytes);  // 256
    uint32_t* s_iw = s_trail + 256;                                                   // 8 words
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_iw + 16);                         // 2 mbarriers
    uint32_t* s_qw = s_iw + 24;                                                       // queue: window index | candidate flag
    uint32_t* s_qm = s_qw + kPrefQueueCap;                                            // queue: good-byte masks of candidates
    uint8_t* s_lut = reinterpret_cast<uint8_t*>(s_qm + 4 * kPrefQueueCap);             // PF_PAIR: class of every byte value
    if (FAMILY == PF_PAIR) {
        s_lut[threadIdx.x] = C.pair_cls[threadIdx.x];  // kPrefThreads == 256
        __syncthreads();
    }
    const bool do_refine = (FAMILY == PF_UTF8) && C.refine != 0;
    uint32_t qn = 0;  // queued windows (uniform across the block)
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = P.W, nchunk = W >> 4;
    const uint32_t tile_bytes = kPrefTileWin * W;
    const long long t_begin = O.tile0 + (long long)blockIdx.x * O.tiles_per_cta;
    const long long t_end = (t_begin + O.tiles_per_cta) < O.tile_end ? (t_begin + O.tiles_per_cta) : O.tile_end;
    uint32_t* const my_list = O.list + (size_t)t_begin * kPrefTileWin;
    const bool piece_start = O.piece_ctas != 0 && (blockIdx.x % O.piece_ctas) == 0;
    uint32_t kept = 0;  // windows this CTA has listed so far (uniform across the block)
    const int64_t full_rows_bytes = (P.len >> 7) << 7;  // the tensor map covers whole 128-byte rows only

    if (use_tma && t_begin < t_end) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar[0], tile_bytes);
            tma_load_2d(smem_raw, &tmap, &s_bar[0], 0, (int32_t)((t_begin * (long long)tile_bytes) >> 7));
        }
        __syncthreads();
    }

    for (long long tile = t_begin; tile < t_end; ++tile) {
        const int64_t lo = (int64_t)tile * tile_bytes;
        const int64_t hi = (lo + tile_bytes) < P.len ? (lo + tile_bytes) : P.len;
        const uint32_t it = (uint32_t)(tile - t_begin);
        uint8_t* const sm = smem_raw + (it & 1u) * kPrefStageBytes;
        if (use_tma) {
            // ---- stage the tile with TMA, double buffered: the next tile is in flight while this one is classified
            if (tid == 0 && tile + 1 < t_end) {
                mbar_expect_tx(&s_bar[(it + 1) & 1u], tile_bytes);
                tma_load_2d(smem_raw + ((it + 1) & 1u) * kPrefStageBytes, &tmap, &s_bar[(it + 1) & 1u], 0,
                            (int32_t)(((tile + 1) * (long long)tile_bytes) >> 7));
            }
            mbar_wait(&s_bar[it & 1u], (it >> 1) & 1u);
            if (hi > full_rows_bytes && lo <= full_rows_bytes) {  // ragged last row of the stream: plain loads
                const int64_t o = full_rows_bytes + tid;
                if (tid < 128) sm[swz((uint32_t)(o - lo))] = o < P.len ? P.in[o] : (uint8_t)0;
                __syncthreads();
            }
        } else
        // ---- fallback: coalesced 16-byte streaming loads, swizzled shared stores --------------------------
        {
            const uint32_t nbytes = (uint32_t)(hi - lo);
            const uint32_t nfull = nbytes >> 4;
            const uint4* src = reinterpret_cast<const uint4*>(P.in + lo);
            for (uint32_t c = tid; c < nfull; c += kPrefThreads) *reinterpret_cast<uint4*>(sm + swz(c * 16u)) = ldg_stream(src + c);
            if ((nbytes & 15u) && tid < 16) {  // ragged stream tail: zero padded, byte stores
                const uint32_t o = nfull * 16u + tid;
                sm[swz(o)] = o < nbytes ? P.in[lo + o] : (uint8_t)0;
            }
            __syncthreads();
        }

        // ---- per-window classification -----------------------------------------------------------------
        const long long w = tile * kPrefTileWin + tid;
        const int64_t ws = lo + (int64_t)tid * W;
        bool valid = w < total_windows && w >= O.w_lo && w < O.w_hi;
        uint32_t wlen = 0;
        if (valid) wlen = (uint32_t)(((ws + W) < P.len ? (ws + W) : P.len) - ws);
        uint32_t m[4] = {0, 0, 0, 0};
        uint32_t mh[4] = {0, 0, 0, 0};  // PrefCfg::sb_rule: bytes >= 0x80 of the window
        bool sure = false, cand = false;
        if constexpr (FAST) {
            // ---- bit planes p7/p6/p5 of the window's 128 bytes (8 conflict-free LDS.128) ----------------------
            uint32_t p7[4], p6[4], p5[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t a7[4] = {0, 0, 0, 0}, a6[4] = {0, 0, 0, 0}, a5[4] = {0, 0, 0, 0};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sm + swz(tid * 128u + (2 * g + h) * 16u));
                    const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * h + j;  // word inside the 32-byte group; a pair of words shares an accumulator
                        const uint32_t wt = (k & 1) ? 0x80402010u : 0x08040201u;
                        a7[k >> 1] = dp4a_u(xs[j] & 0x80808080u, wt, a7[k >> 1]);  // 8 flags at bits 7..14
                        a6[k >> 1] = dp4a_u(xs[j] & 0x40404040u, wt, a6[k >> 1]);  //            bits 6..13
                        a5[k >> 1] = dp4a_u(xs[j] & 0x20202020u, wt, a5[k >> 1]);  //            bits 5..12
                    }
                }
                p7[g] = pack_plane<7>(a7);
                p6[g] = pack_plane<6>(a6);
                p5[g] = pack_plane<5>(a5);
            }
            // ---- class logic on the planes -> good-byte mask m (bit i = byte i of the window) -----------------
            if (FAMILY == PF_UTF8) {
                uint32_t cn[4], lp[4], lc[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    cn[g] = p7[g] & ~p6[g];                                                  // 10xxxxxx
                    lp[g] = DEFSHAPE ? (p7[g] & p6[g] & ~p5[g]) : (p7[g] & p6[g] & lop3_sel(p5[g], K.kh[3], K.kh[2]));
                    lc[g] = DEFSHAPE ? lp[g] : (lp[g] | (cn[g] & K.multi));
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t ap = DEFSHAPE ? (~p7[g] & (p6[g] | p5[g])) : (~p7[g] & blk2(p6[g], p5[g], K.ka));
                    // byte i: Cn(i + 1) / LC(i - 1); both window edges are favourable (pref_good)
                    const uint32_t ncn = __funnelshift_r(cn[g], g < 3 ? cn[g < 3 ? g + 1 : 3] : 0xFFFFFFFFu, 1);
                    const uint32_t pl = __funnelshift_l(g > 0 ? lc[g > 0 ? g - 1 : 0] : 0xFFFFFFFFu, lc[g], 1);
                    m[g] = ap | (lp[g] & ncn) | (cn[g] & pl);
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) m[g] = lop3_sel(p7[g], blk2(p6[g], p5[g], K.kh), blk2(p6[g], p5[g], K.ka));
                if (FAMILY == PF_UNIT) {
                    // keep the flag of each unit's most significant byte and spread it over the unit
                    uint32_t fh[4] = {m[0] & K.hi_mask[0], m[1] & K.hi_mask[1], m[2] & K.hi_mask[2], m[3] & K.hi_mask[3]};
                    uint32_t gm[4] = {fh[0], fh[1], fh[2], fh[3]};
#pragma unroll
                    for (int sft = 1; sft < 4; ++sft) {
                        if ((uint32_t)sft < C.unit) {
                            uint32_t t[4];
                            if (K.spread_left) {
                                t[0] = __funnelshift_r(fh[0], fh[1], sft);
                                t[1] = __funnelshift_r(fh[1], fh[2], sft);
                                t[2] = __funnelshift_r(fh[2], fh[3], sft);
                                t[3] = fh[3] >> sft;
                            } else {
                                t[0] = fh[0] << sft;
                                t[1] = __funnelshift_l(fh[0], fh[1], sft);
                                t[2] = __funnelshift_l(fh[1], fh[2], sft);
                                t[3] = __funnelshift_l(fh[2], fh[3], sft);
                            }
                            gm[0] |= t[0]; gm[1] |= t[1]; gm[2] |= t[2]; gm[3] |= t[3];
                        }
                    }
                    m[0] = gm[0] | K.edge_mask[0]; m[1] = gm[1] | K.edge_mask[1];
                    m[2] = gm[2] | K.edge_mask[2]; m[3] = gm[3] | K.edge_mask[3];
                }
            }
            // ---- run test on (last T-1 flags of the previous window : m), 160 bits ---------------------------
            // the window before the tile's first one is not known here: taken as all good (pref_interesting_ref)
            uint32_t prev_top = __shfl_up_sync(0xffffffffu, m[3], 1);
            if (lane == 31) s_trail[warp] = m[3];
            __syncthreads();
            if (lane == 0) prev_top = warp ? s_trail[warp - 1] : 0xFFFFFFFFu;
            const uint32_t T = C.T;
            uint32_t r[5] = {T > 1 ? (prev_top & ~(0xFFFFFFFFu >> (T - 1))) : 0u, m[0], m[1], m[2], m[3]};
            // r &= r << s marks the END of every run of ones at least (have + s) long
            auto and_shl = [&](uint32_t sft) {
                r[4] &= __funnelshift_l(r[3], r[4], sft);
                r[3] &= __funnelshift_l(r[2], r[3], sft);
                r[2] &= __funnelshift_l(r[1], r[2], sft);
                r[1] &= __funnelshift_l(r[0], r[1], sft);
                r[0] &= r[0] << sft;
            };
            uint32_t have = 1;
            if (T >= 2) { and_shl(1); have = 2; }
            if (T >= 4) { and_shl(2); have = 4; }
            if (T >= 8) { and_shl(4); have = 8; }
            if (T >= 16) { and_shl(8); have = 16; }
            if (T >= 32) { and_shl(16); have = 32; }
            if (T > have) and_shl(T - have);
            // a run that ends inside the first T-1 bytes of the window began in the previous window; one that ends at
            // byte T-1 covers the window's first T bytes: both touch the left boundary (pref_interesting_ref)
            const uint32_t cross_bits = 0xFFFFFFFFu >> (32 - T);
            const bool crossing = (r[1] & cross_bits) != 0;
            const bool inwin = ((r[1] & ~cross_bits) | r[2] | r[3] | r[4]) != 0;
            if (valid) {
                const bool forced = (w == O.w_first) || (piece_start && w == t_begin * (long long)kPrefTileWin) || (w == total_windows - 1) || (wlen < W) || (P.is_last && ws + wlen == P.len);
                sure = forced || crossing;
                if (!sure && inwin) { if (do_refine) cand = true; else sure = true; }
            }
        } else {
        if (valid) {
            // Fully unrolled over the (up to) 8 16-byte chunks of the window: all mask indices are static.
            // acc[p] collects the flags of data words 2p and 2p+1 of the current 32-byte group (dp4a,
            // flags sit at bit 7 of each byte, so acc[p] = (8 flags) << 7).
            uint32_t acc[4] = {0, 0, 0, 0};
            uint32_t acch[4] = {0, 0, 0, 0};
            auto puth = [&](int widx, uint32_t x) {  // like put(): bit 7 of every byte of data word widx -> mh
                const int k = widx & 7;
                acch[k >> 1] = dp4a_u(x & 0x80808080u, (k & 1) ? 0x80402010u : 0x08040201u, acch[k >> 1]);
                if (k == 7) {
                    mh[widx >> 3] = (acch[0] >> 7) | (acch[1] << 1) | (acch[2] << 9) | (acch[3] << 17);
                    acch[0] = acch[1] = acch[2] = acch[3] = 0;
                }
            };
            uint32_t pA = 0, pL = 0, pC = 0, ppLC = 0xFFFFFFFFu;  // PF_UTF8: classes of the word awaiting its right neighbour
            auto put = [&](int widx, uint32_t gflags) {  // widx is a compile-time constant at every call site
                const int k = widx & 7;
                acc[k >> 1] = dp4a_u(gflags & 0x80808080u, (k & 1) ? 0x80402010u : 0x08040201u, acc[k >> 1]);
                if (k == 7) {
                    m[widx >> 3] = (acc[0] >> 7) | (acc[1] << 1) | (acc[2] << 9) | (acc[3] << 17);
                    acc[0] = acc[1] = acc[2] = acc[3] = 0;
                }
            };
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if ((uint32_t)c < nchunk) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sm + swz(tid * W + c * 16u));
                    const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t x = xs[j];
                        const uint32_t x1 = x << 1, x2 = x << 2;
                        if (C.sb_rule) puth(4 * c + j, x);
                        if (FAMILY == PF_UTF8 || FAMILY == PF_PAIR) {
                            uint32_t cn, lp, ap;  // flags at bit 7 of every byte: trail candidate, lead candidate, passing single byte
                            if (FAMILY == PF_PAIR) {
                                const uint32_t cw = (uint32_t)s_lut[x & 0xFFu] | ((uint32_t)s_lut[(x >> 8) & 0xFFu] << 8) |
                                                    ((uint32_t)s_lut[(x >> 16) & 0xFFu] << 16) | ((uint32_t)s_lut[x >> 24] << 24);
                                ap = cw << 7; lp = cw << 6; cn = cw << 5;
                            } else {
                                cn = x & ~x1;  // 10xxxxxx
                                // 11xxxxxx in a passing lead block / ASCII in a passing block
                                lp = DEFSHAPE ? (x & x1 & ~x2) : (x & x1 & lop3_sel(x2, K.kh[3], K.kh[2]));
                                ap = DEFSHAPE ? (~x & (x1 | x2)) : (~x & blk2(x1, x2, K.ka));
                            }
                            if (c > 0 || j > 0) {
                                const uint32_t ncn = __funnelshift_r(pC, cn, 8);    // byte i: Cn(i + 1)
                                const uint32_t lcp = DEFSHAPE ? pL : (pL | (pC & K.multi));
                                const uint32_t pl = __funnelshift_l(ppLC, lcp, 8);  // byte i: LC(i - 1)
                                put(4 * c + j - 1, pA | (pL & ncn) | (pC & pl));
                                ppLC = lcp;
                            }
                            pA = ap; pL = lp; pC = cn;
                        } else {
                            put(4 * c + j, lop3_sel(x, blk2(x1, x2, K.kh), blk2(x1, x2, K.ka)));
                        }
                    }
                    if ((FAMILY == PF_UTF8 || FAMILY == PF_PAIR) && (uint32_t)c == nchunk - 1) {  // flush the last word: right edge favourable
                        const uint32_t ncn = __funnelshift_r(pC, 0xFFFFFFFFu, 8);
                        const uint32_t pl = __funnelshift_l(ppLC, DEFSHAPE ? pL : (pL | (pC & K.multi)), 8);
                        put(4 * c + 3, pA | (pL & ncn) | (pC & pl));
                    }
                    if ((c & 1) == 0 && (uint32_t)c == nchunk - 1) {  // odd number of chunks: half a mask word is pending
                        m[c >> 1] = (acc[0] >> 7) | (acc[1] << 1);
                        mh[c >> 1] = (acch[0] >> 7) | (acch[1] << 1);
                    }
                }
            }
            if (FAMILY == PF_UNIT) {
                // keep the flag of each unit's most significant byte and spread it over the unit
                uint32_t fh[4] = {m[0] & K.hi_mask[0], m[1] & K.hi_mask[1], m[2] & K.hi_mask[2], m[3] & K.hi_mask[3]};
                uint32_t gm[4] = {fh[0], fh[1], fh[2], fh[3]};
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    if ((uint32_t)s < C.unit) {
                        uint32_t t[4];
                        if (K.spread_left) {
                            t[0] = __funnelshift_r(fh[0], fh[1], s);
                            t[1] = __funnelshift_r(fh[1], fh[2], s);
                            t[2] = __funnelshift_r(fh[2], fh[3], s);
                            t[3] = fh[3] >> s;
                        } else {
                            t[0] = fh[0] << s;
                            t[1] = __funnelshift_l(fh[0], fh[1], s);
                            t[2] = __funnelshift_l(fh[1], fh[2], s);
                            t[3] = __funnelshift_l(fh[2], fh[3], s);
                        }
                        gm[0] |= t[0]; gm[1] |= t[1]; gm[2] |= t[2]; gm[3] |= t[3];
                    }
                }
                m[0] = gm[0] | K.edge_mask[0]; m[1] = gm[1] | K.edge_mask[1];
                m[2] = gm[2] | K.edge_mask[2]; m[3] = gm[3] | K.edge_mask[3];
            }
            // bits beyond the window (W < 128 or the ragged last window) are not bytes of the window
            if (wlen < 128) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int lo_bit = q4 * 32;
                    if ((int)wlen <= lo_bit) m[q4] = 0;
                    else if ((int)wlen < lo_bit + 32) m[q4] &= (1u << (wlen - lo_bit)) - 1u;
                }
            }
        }
        uint32_t lead = 0, trail = 0;
        bool longrun = false, sb_lead_hi = false;
        if (valid) {
            uint32_t nm[4] = {~m[0], ~m[1], ~m[2], ~m[3]};  // bad bytes of the window
            if (wlen < 128) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int lo_bit = q4 * 32;
                    if ((int)wlen <= lo_bit) nm[q4] = 0;
                    else if ((int)wlen < lo_bit + 32) nm[q4] &= (1u << (wlen - lo_bit)) - 1u;
                }
            }
            if ((nm[0] | nm[1] | nm[2] | nm[3]) == 0) { lead = wlen; trail = wlen; }
            else { lead = ctz128(nm); trail = wlen - 128u + clz128(nm); }
            uint32_t rs[4];
            longrun = has_run128(m, C.T, rs);
            if (C.kill_trail != 0) {
                // PrefWin::kill: a good run of >= kill_trail bytes, or one that covers the window from its first byte,
                // reaches one of the last 4 bytes
                for (uint32_t j = 0; j < 4 && j < wlen; ++j) {
                    const uint32_t p = wlen - 1u - j;                       // run must cover byte p
                    if ((m[p >> 5] >> (p & 31u)) & 1u) {
                        int hb = -1;                                        // highest bad byte below p
                        for (int q4 = (int)(p >> 5); q4 >= 0 && hb < 0; --q4) {
                            uint32_t bad = nm[q4];
                            if (q4 == (int)(p >> 5)) bad &= (p & 31u) == 31u ? 0xFFFFFFFFu : ((1u << ((p & 31u) + 1u)) - 1u);
                            if (bad) hb = q4 * 32 + 31 - __clz(bad);
                        }
                        if (hb < 0 || p - (uint32_t)hb >= C.kill_trail) trail |= 0x80000000u;
                    }
                }
            }
            if (C.sb_rule) {
                // PrefWin::lead_hi / trail_hi: the leading / trailing good run holds a byte >= 0x80 (any good byte for the
                // unit families), or the trailing run covers the whole window
                bool lh = false, th = false;
                const uint32_t trail_len = trail & 0x3FFFFFFFu;  // bit 31 may hold the kill flag (both rules on)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t hw = (FAMILY == PF_UNIT) ? m[q4] : (mh[q4] & m[q4]);
                    const int lo_bit = q4 * 32;
                    const int nl = (int)lead - lo_bit;             // bits of this word inside the leading run
                    const uint32_t lmask = nl <= 0 ? 0u : (nl >= 32 ? 0xFFFFFFFFu : ((1u << nl) - 1u));
                    const int t0 = (int)(wlen - trail_len) - lo_bit;   // first bit of the trailing run, relative to this word
                    const uint32_t tmask = t0 >= 32 ? 0u : (t0 <= 0 ? 0xFFFFFFFFu : ~((1u << t0) - 1u));
                    lh = lh || (hw & lmask) != 0;
                    th = th || (hw & tmask) != 0;               // bits beyond wlen are zero in m
                }
                if (trail_len > 0 && (th || trail_len == wlen)) trail |= 0x40000000u;
                sb_lead_hi = lead > 0 && lead < wlen && lh;
            }

        }
        s_trail[tid] = trail;
        __syncthreads();
        // sure: listed whatever the refinement says; cand: listed only if a long run holds enough chars
        if (valid) {
            const bool forced = (w == O.w_first) || (piece_start && w == t_begin * (long long)kPrefTileWin) || (w == total_windows - 1) || (wlen < W) || (P.is_last && ws + wlen == P.len);
            if (forced) sure = true;
            else if (tid == 0) sure = lead >= 1 || C.kill_trail != 0 || (trail & 0x40000000u) != 0;
            else {
                const uint32_t pt = s_trail[tid - 1];  // bit 31: the previous window's kill flag, bit 30: its trail_hi
                sure = (lead >= 1 && (pt & 0x3FFFFFFFu) + lead >= C.T) || (pt >> 31) != 0 ||
                       ((trail & 0x40000000u) != 0 && ((pt & 0x40000000u) != 0 || sb_lead_hi));  // PrefCfg::sb_rule
            }
            if (!sure && longrun) { if (do_refine) cand = true; else sure = true; }
        }
        }
        // queue the window (tile order = (warp, lane) order); candidates carry their good-byte mask
        const bool push = sure || cand;
        const uint32_t ib = __ballot_sync(0xffffffffu, push);
        if (lane == 0) s_iw[warp] = __popc(ib);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kPrefThreads / 32; ++k) {
            const uint32_t c = s_iw[k];
            if ((uint32_t)k < warp) before += c;
            total += c;
        }
        if (push) {
            const uint32_t qi = qn + before + __popc(ib & ((1u << lane) - 1u));
            s_qw[qi] = (uint32_t)w | (cand ? 0x80000000u : 0u);  // window indices stay below 2^31 (checked on the host)
            if (cand) { s_qm[qi * 4 + 0] = m[0]; s_qm[qi * 4 + 1] = m[1]; s_qm[qi * 4 + 2] = m[2]; s_qm[qi * 4 + 3] = m[3]; }
        }
        qn += total;
        __syncthreads();
        // flush when another tile might not fit: every thread settles up to two queued windows in parallel
        if (qn > kPrefQueueCap - kPrefTileWin || tile + 1 == t_end) {
#pragma unroll 1
            for (uint32_t base = 0; base < qn; base += kPrefThreads) {
                const uint32_t qi = base + tid;
                bool keep = false;
                uint32_t wq = 0;
                if (qi < qn) {
                    const uint32_t e = s_qw[qi];
                    wq = e & 0x7FFFFFFFu;
                    keep = true;
                    if (e & 0x80000000u) {
                        const uint32_t mm[4] = {s_qm[qi * 4], s_qm[qi * 4 + 1], s_qm[qi * 4 + 2], s_qm[qi * 4 + 3]};
                        keep = pref_refine_utf8(P.in + (size_t)wq * W, nchunk, mm, C.T, C.n_chars);
                    }
                }
                const uint32_t kb = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) s_iw[warp] = __popc(kb);
                __syncthreads();
                uint32_t bf = 0, tt = 0;
#pragma unroll
                for (int k = 0; k < kPrefThreads / 32; ++k) {
                    const uint32_t c = s_iw[k];
                    if ((uint32_t)k < warp) bf += c;
                    tt += c;
                }
                if (keep) my_list[kept + bf + __popc(kb & ((1u << lane) - 1u))] = wq;
                kept += tt;
                __syncthreads();
            }
            qn = 0;
        }
    }
    if (tid == 0) O.cta_count[blockIdx.x] = kept;
}

// exclusive scan of the per-CTA counts (at most 1024 CTAs) -> cta_off[0..ncta], total -> counters[2]
__global__ void __launch_bounds__(1024) sx_list_offsets_kernel(const uint32_t* __restrict__ cta_count, uint32_t* __restrict__ cta_off,
                                                               uint32_t ncta, unsigned long long* counters) {
    __shared__ uint32_t wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t v = tid < ncta ? cta_count[tid] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t x = wsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += t;
        }
        wsum[lane] = x;  // inclusive
    }
    __syncthreads();
    const uint32_t base = warp ? wsum[warp - 1] : 0u;
    if (tid < ncta) cta_off[tid] = base + inc - v;
    if (tid == 1023) { cta_off[ncta] = base + inc; counters[2] = base + inc; }
}

// Block-kernel path, direct host output: sx_exact_kernel reserves record ranges block by block in completion order
// (block_desc[b] = {first record, count}); stream order is the order of the blocks.  sx_order_scan_kernel turns the
// counts into stream-order positions, sx_order_write_kernel writes every record as a finding (C-ABI layout) to its
// position in the collection's pinned set -- same result as the sparse pipeline's gather kernel.
__global__ void __launch_bounds__(1024) sx_order_scan_kernel(const uint2* __restrict__ desc, unsigned long long* __restrict__ pos,
                                                             unsigned long long nb) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < nb; base += 1024) {
        const unsigned long long k = base + tid;
        const unsigned long long v = k < nb ? desc[k].y : 0ull;
        unsigned long long a = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, a, d);
            if (lane >= (uint32_t)d) a += t;
        }
        if (lane == 31) ws[warp] = a;
        __syncthreads();
        if (warp == 0) {
            unsigned long long x = ws[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += t;
            }
            ws[lane] = x;
        }
        __syncthreads();
        const unsigned long long o = carry + (warp ? ws[warp - 1] : 0ull);
        if (k < nb) pos[k] = o + a - v;
        __syncthreads();
        if (tid == 1023) carry = o + a;
        __syncthreads();
    }
}
// one warp per block descriptor; 32 findings at a time are assembled in shared memory and leave as contiguous
// 16-byte stores (512 bytes per warp instruction: large PCIe write transactions)
__global__ void __launch_bounds__(256)
sx_order_write_kernel(const __grid_constant__ ScanParams P, const ScanOut O, const uint2* __restrict__ desc,
                      const unsigned long long* __restrict__ pos, unsigned long long nb, unsigned long long nrec) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long b = (unsigned long long)blockIdx.x * 8 + warp; b < nb; b += (unsigned long long)gridDim.x * 8) {
        const uint2 d = desc[b];
        const unsigned long long p0 = pos[b];
        for (uint32_t k0 = 0; k0 < d.y; k0 += 32) {
            const uint32_t cnt = d.y - k0 < 32u ? d.y - k0 : 32u;
            if (lane < cnt) {
                const Record r = O.recs[(unsigned long long)d.x + k0 + lane];
                const unsigned long long idx = p0 + k0 + lane;
                write_host_finding(P, O.host_findings + idx, r);  // consecutive lanes, consecutive 16-byte records
                if (idx == 0) O.final_state->first_flags = r.flags;
                if (idx == nrec - 1) O.final_state->last_flags = r.flags;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
sx_materialize_kernel(const __grid_constant__ ScanParams P, const Record* __restrict__ recs, unsigned long long n,
                      uint8_t* __restrict__ text, unsigned long long text_cap) {
    const GlobalSrc g{P.in, P.pend};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const Record r = recs[i];
        if (r.text_off + r.text_len <= text_cap) transcode_record(P, g, r, text + r.text_off);
    }
}

__host__ __device__ inline uint64_t sx_mix(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// byte i of the stream = byte (i & 7) (little endian) of sx_mix(seed, i >> 3)
__global__ void __launch_bounds__(256) sx_fill_kernel(uint8_t* dst, unsigned long long len, uint64_t seed, uint64_t off0) {
    const unsigned long long nwords = (len + 7) / 8 + 1;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t gw = (off0 >> 3) + w;
        const uint64_t v = sx_mix(seed, gw);
        const long long base = (long long)(gw * 8 - off0);
        if (base >= 0 && base + 8 <= (long long)len && ((reinterpret_cast<uintptr_t>(dst) + base) & 7) == 0) {
            *reinterpret_cast<uint64_t*>(dst + base) = v;
        } else {
            for (int b = 0; b < 8; ++b) {
                const long long o = base + b;
                if (o >= 0 && o < (long long)len) dst[o] = (uint8_t)(v >> (8 * b));
            }
        }
    }
}

}  // namespace sx

// =============================================================================================
// Host side: C ABI
// =============================================================================================
using namespace sx;

#include "sx_mb_tables.inc"  // kSxBig5Index, kSxJis0208Index, kSxJis0212Index (generated, tools/gen_multibyte_tables.py)

static thread_local int g_err_code = SX_OK;
static thread_local std::string g_err_msg;

static void set_err(int code, const std::string& msg) {
    g_err_code = code;
    g_err_msg = msg;
}
static bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_err(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? SX_ERR_NO_DEVICE : SX_ERR_CUDA,
            std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}
#define CK(call)                                  \
    do {                                          \
        if (!cuda_ok((call), #call)) return fail; \
    } while (0)

constexpr int kMaxPieces = 32;
// Host-visible control block of a state (pinned, device-mapped): what the kernels of the sparse pipeline report per piece
// and the final state of the scan; the host reads it after an event, no small device-to-host copies.
struct HostCtl {
    PieceSummary piece[kMaxPieces];
    FinalState fin;
    uint8_t tail[8];  // the last bytes of a device-resident buffer (pending decoder bytes of the ScannerState)
};
struct PieceEvents {
    cudaEvent_t p0 = nullptr, p1 = nullptr;  // around the piece's prefilter launch
    cudaEvent_t b0 = nullptr, b1 = nullptr;  // around its exact stage
    cudaEvent_t sev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t sord[2] = {nullptr, nullptr};
    cudaEvent_t sc = nullptr;                // after the piece's sx_sp_scan_kernel (the next piece's scan builds on its totals)
    cudaEvent_t gp[kGatherParts] = {nullptr, nullptr, nullptr, nullptr};  // after every part of its gather
};

// Asynchronous boundary: the reference's scanner threads hand their collections to the merger through a channel
// (main.rs:98, :161); here every state owns one worker thread that runs its scans in call order and a handle per call
// that the consumer waits on (sx_fc_wait).
struct sx_pending {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    sx_finding_collection* fc = nullptr;
    int err = SX_OK;
    std::string msg;
};
struct ScanWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                job = std::move(q.front());
                q.pop_front();
            }
            job();
        }
    }
};

struct sx_scanner_state {
    sx_mission m;
    int device;
    // ScannerState (scanner.rs:40-69)
    uint64_t consumed;
    bool cut;
    std::vector<uint8_t> leftover;
    // raw bytes still inside the decoder (reproduce Decoder state by look-back on the device)
    uint8_t pend[8];
    int npend;
    // device resources, grown on demand
    uint8_t* d_in = nullptr; size_t d_in_cap = 0;
    Record* d_recs = nullptr; size_t rec_cap = 0;
    uint8_t* d_text = nullptr; size_t text_cap = 0;
    uint2* d_blocks = nullptr; size_t blocks_cap = 0;
    uint32_t* d_ccount = nullptr; size_t ccount_cap = 0;
    uint32_t* d_coff = nullptr; size_t coff_cap = 0;
    uint32_t* d_list = nullptr; size_t list_cap = 0;
    int use_prefilter = 1;
    int use_tma = 1;
    int use_sparse = 1;
    int use_direct = 1;  // findings written by the GPU in their C-ABI form (0: record download + host conversion)
    int pieces = 0;      // test / tuning hook: number of pieces of the sparse pipeline (0: automatic)
    // sparse pipeline: per-entry arrays (EntryHot, null carries, staged records), per-chunk totals, tables, queues,
    // compact list, per-piece control blocks, staging of the findings in their C-ABI form
    uint8_t* d_entries = nullptr; size_t entries_cap = 0;
    uint8_t* d_btot = nullptr; size_t btot_cap = 0;
    uint8_t* d_tables = nullptr; size_t tables_cap = 0;
    uint32_t* d_queue = nullptr; size_t queue_cap = 0;
    uint32_t* d_clist = nullptr; size_t clist_cap = 0;
    unsigned long long* d_findings = nullptr; size_t findings_cap = 0;  // staging: 8-byte wire records
    unsigned long long* d_pagebase = nullptr; size_t pagebase_cap = 0;
    PieceCtl* d_ctl = nullptr;
    HostCtl* h_ctl = nullptr;
    unsigned long long* d_bpos = nullptr; size_t bpos_cap = 0;  // block path, direct output: stream-order position of every block
    uint32_t last_ncta = 0; size_t last_region_stride = 0;
    bool last_list_compact = false; size_t last_list_len = 0;
    const uint32_t* last_list_base = nullptr;
    std::vector<std::pair<size_t, size_t>> last_piece_lists;  // (offset into d_clist, entries) per piece
    // pinned host staging for result downloads
    Record* h_recs = nullptr; size_t h_recs_cap = 0;
    uint8_t* h_text = nullptr; size_t h_text_cap = 0;
    uint2* h_blocks = nullptr; size_t h_blocks_cap = 0;
    unsigned long long* d_counters = nullptr;
    FinalState* d_final = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    std::vector<PieceEvents> pev;
    cudaStream_t side = nullptr;
    cudaStream_t sA = nullptr, sB = nullptr, sC = nullptr;  // prefilter / set-up of the exact stage / copies to the host
    static constexpr int kLanes = 4;                        // exact stages of consecutive pieces run side by side: their kernels are
    cudaStream_t sBk[kLanes] = {nullptr, nullptr, nullptr, nullptr};    // latency bound, a piece's chain takes longer than its prefilter
    cudaStream_t sidek[kLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_setup = nullptr;
    int num_sms = 0;
    double rec_per_byte = 1.0 / 1024, text_per_byte = 1.0 / 64;
    double listed_frac = 1.0 / 16;  // windows the prefilter kept in the previous scan (sizes the entry arrays of large calls)
    // direct host output: cap on the pinned set of a state's FIRST scan (the estimate from rec_per_byte is generous
    // until a scan has been seen; an overflow simply reruns the exact stage with the counted sizes)
    size_t host_rec_hint = 1u << 20, host_text_hint = 16u << 20;
    bool have_history = false;
    size_t last_nrec = 0, last_ntext = 0, last_len = 0;  // previous scan: sizes the next pinned set (+ 1/16)
    sx_scan_stats stats;
    struct ScanWorker* worker = nullptr;  // asynchronous calls (sx_scan_*_async): one thread per state, calls run in call order
    cudaStream_t own = nullptr;           // the streaming driver's scans of this state (sx_scan_reader)
};

// uninitialised storage: value-initialising tens of MB per call showed up in the whole-job time
template <class T>
struct RawVec {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool external = false;  // storage owned by someone else (a pinned set)
    ~RawVec() { if (!external) free(p); }
    void adopt(T* q, size_t count) { if (!external) free(p); p = q; n = count; cap = count; external = true; }
    void reserve(size_t c) { if (c > cap) { p = (T*)realloc(p, c * sizeof(T)); cap = c; } }
    void resize(size_t c) { reserve(c); n = c; }
    void push_back(const T& x) { if (n == cap) reserve(cap ? cap * 2 : 16); p[n++] = x; }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    const T* begin() const { return p; }
    const T* end() const { return p + n; }
};
// Pinned, device-mapped host memory the GPU writes a call's findings into (sx_sp_gather_kernel) and the finding
// text is downloaded to.  A collection built that way owns its set; sets are recycled through a small pool because
// pinning memory costs far more than a scan.
struct PinnedSet {
    uint8_t* f = nullptr; size_t fcap = 0;  // findings as they come off the wire: fcap BYTES (8-byte records + page bases of the
                                            // per-stage pipeline, 16-byte records of the block path)
    uint8_t* t = nullptr; size_t tcap = 0;  // text bytes
};
// 8-byte records + one text offset per page of kWirePage records
static size_t wire8_bytes(size_t cap) { return cap * 8 + (cap / kWirePage + 2) * 8; }
static size_t wire8_cap(size_t bytes) { return bytes < 64 ? 0 : (size_t)((double)(bytes - 32) * kWirePage / (8.0 * (kWirePage + 1))); }
static std::mutex g_pool_mu;
static std::vector<PinnedSet> g_pool;
static void pinned_free(PinnedSet& s) {
    if (s.f) cudaFreeHost(s.f);
    if (s.t) cudaFreeHost(s.t);
    s = PinnedSet();
}
static size_t pinned_bytes(const PinnedSet& s) { return s.fcap + s.tcap; }
// Pool policy: pinning memory costs far more than a scan (seconds per GB), so released sets are kept for reuse -- up
// to kPoolTotal bytes per process (SX_PINNED_POOL_MIB overrides; page-locked memory is not swappable).  A process
// typically runs several missions side by side, each returning a collection per call, so several large sets are live
// at once; a request is served by the smallest pooled set that fits, and when the pool is full the least recently
// released sets go first.
static size_t pool_total_cap() {
    static size_t cap = 0;
    if (!cap) {
        cap = 24ull << 30;
        if (const char* ev = getenv("SX_PINNED_POOL_MIB")) { const long long v = atoll(ev); if (v >= 0) cap = (size_t)v << 20; }
    }
    return cap;
}
static bool pinned_acquire(size_t fcap, size_t tcap, PinnedSet* out) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].fcap >= fcap && g_pool[i].tcap >= tcap && (best < 0 || pinned_bytes(g_pool[i]) < pinned_bytes(g_pool[best]))) best = (int)i;
        // a pooled set several times larger than needed stays for the caller that needs it
        if (best >= 0 && pinned_bytes(g_pool[best]) <= 4 * (fcap + tcap) + (64ull << 20)) {
            *out = g_pool[best];
            g_pool.erase(g_pool.begin() + best);
            return true;
        }
    }
    PinnedSet s;
    s.fcap = fcap + fcap / 16 + 16384;
    s.tcap = tcap + tcap / 16 + 65536;
    if (cudaHostAlloc((void**)&s.f, s.fcap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc((void**)&s.t, s.tcap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        pinned_free(s);
        // out of pinnable memory: drop the pool and try once more
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            for (auto& q : g_pool) pinned_free(q);
            g_pool.clear();
        }
        s.fcap = fcap + 16384;
        s.tcap = tcap + 65536;
        if (cudaHostAlloc((void**)&s.f, s.fcap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess ||
            cudaHostAlloc((void**)&s.t, s.tcap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            pinned_free(s);
            return false;
        }
    }
    *out = s;
    return true;
}
static void pinned_release(PinnedSet& s) {
    if (!s.f && !s.t) return;
    std::vector<PinnedSet> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool.push_back(s);  // most recently released last
        s = PinnedSet();
        size_t total = 0;
        for (const auto& q : g_pool) total += pinned_bytes(q);
        while (!g_pool.empty() && (total > pool_total_cap() || g_pool.size() > 64)) {
            total -= pinned_bytes(g_pool.front());
            drop.push_back(g_pool.front());
            g_pool.erase(g_pool.begin());
        }
    }
    for (auto& q : drop) pinned_free(q);
}

// Device copies of the Big5 / EUC-JP index tables, one set per device, made on first use.
struct MbDeviceTables { uint32_t* big5 = nullptr; uint32_t* jis0208 = nullptr; uint32_t* jis0212 = nullptr; };
static std::mutex g_mb_mu;
static std::vector<MbDeviceTables> g_mb_tables;
static bool mb_tables_for(int device, MbDeviceTables* out) {
    std::lock_guard<std::mutex> lk(g_mb_mu);
    if ((size_t)device >= g_mb_tables.size()) g_mb_tables.resize((size_t)device + 1);
    MbDeviceTables& t = g_mb_tables[device];
    if (!t.big5) {
        auto up = [](uint32_t** d, const uint32_t* h, size_t n) {
            return cudaMalloc(d, n * sizeof(uint32_t)) == cudaSuccess && cudaMemcpy(*d, h, n * sizeof(uint32_t), cudaMemcpyHostToDevice) == cudaSuccess;
        };
        if (!up(&t.big5, kSxBig5Index, sizeof kSxBig5Index / 4) || !up(&t.jis0208, kSxJis0208Index, sizeof kSxJis0208Index / 4) ||
            !up(&t.jis0212, kSxJis0212Index, sizeof kSxJis0212Index / 4)) {
            cudaGetLastError();
            t = MbDeviceTables();
            return false;
        }
    }
    *out = t;
    return true;
}

// A collection either holds its findings expanded (v, the record-download path) or as they came off the wire: 16-byte
// records + text in a pinned set, expanded to sx_finding on demand in pages of kPage findings (a consumer that walks a
// 100-million-finding collection pays for what it touches, the scan does not pay for any of it).
struct sx_finding_collection {
    static constexpr size_t kPage = kWirePage;  // expansion granularity = the wire format's page
    RawVec<sx_finding> v;
    RawVec<uint8_t> text;
    PinnedSet set;  // direct output: wire records and the finding text live here
    int wire = 0;                 // 0: expanded (v), 8: per-stage pipeline's records + page bases, 16: block path's records
    size_t wcap = 0;              // wire == 8: record capacity of the set (the page bases follow the records)
    size_t n = 0;                 // findings
    uint64_t base = 0;            // stream position of the call's first byte (wire positions are relative to it)
    int16_t file_id = -1;
    uint8_t mission_id = 0;
    bool have_first = false;      // the first finding's text is in `text` (host-carried leftover in front of it)
    std::vector<uint8_t> page_done;
    bool all_done = false;
    std::mutex mu;
    uint64_t first_byte_position = 0;
    int str_buf_overflow = 0;
    ~sx_finding_collection() { pinned_release(set); }
    const unsigned long long* page_bases() const { return reinterpret_cast<const unsigned long long*>(set.f + wcap * 8); }
    // text of wire record i (without a host-carried prefix)
    void wire_text(size_t i, const uint8_t** p, size_t* len) const {
        if (wire == 16) {
            const WireFinding* w = reinterpret_cast<const WireFinding*>(set.f);
            *p = set.t + (w[i].b & 0xFFFFFFFFFFull);
            *len = (size_t)((w[i].a >> 40) & 0x3FFFFFu);
            return;
        }
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(set.f);
        unsigned long long off = page_bases()[i / kWirePage];
        for (size_t j = (i / kWirePage) * kWirePage; j < i; ++j) off += (w[j] >> 40) & 0x1FFFFFu;
        *p = set.t + off;
        *len = (size_t)((w[i] >> 40) & 0x1FFFFFu);
    }
    void expand(size_t i0, size_t i1) {  // [i0, i1) lies inside one page
        sx_finding* const out = v.data();
        sx_finding f;
        f.input_file_id = file_id;
        f.mission_id = mission_id;
        f.in_start = 0;
        f.in_len = 0;
        if (wire == 16) {
            const WireFinding* const w = reinterpret_cast<const WireFinding*>(set.f);
            for (size_t i = i0; i < i1; ++i) {
                f.position = base + (w[i].a & 0xFFFFFFFFFFull);
                f.precision = (uint8_t)(w[i].a >> 62);
                f.completes_previous = (uint8_t)((w[i].b >> 40) & 1u);
                f.s = set.t + (w[i].b & 0xFFFFFFFFFFull);
                f.s_len = (uint32_t)((w[i].a >> 40) & 0x3FFFFFu);
                out[i] = f;
            }
        } else {
            const unsigned long long* const w = reinterpret_cast<const unsigned long long*>(set.f);
            unsigned long long off = i0 < i1 ? page_bases()[i0 / kWirePage] : 0ull;
            for (size_t i = i0; i < i1; ++i) {
                const unsigned long long a = w[i];
                f.position = base + (a & 0xFFFFFFFFFFull);
                f.precision = (uint8_t)((a >> 61) & 3u);
                f.completes_previous = (uint8_t)(a >> 63);
                f.s_len = (uint32_t)((a >> 40) & 0x1FFFFFu);
                f.s = set.t + off;
                off += f.s_len;
                out[i] = f;
            }
        }
        if (have_first && i0 == 0 && i1 > 0) { out[0].s = text.data(); out[0].s_len = (uint32_t)(text.size() - 1); }
    }
    void ensure_page(size_t i) {
        if (!wire || all_done) return;
        const size_t pg = i / kPage;
        std::lock_guard<std::mutex> lk(mu);
        if (v.cap < n) v.resize(n);
        if (page_done.empty()) page_done.assign((n + kPage - 1) / kPage, 0);
        if (!page_done[pg]) { expand(pg * kPage, std::min(n, (pg + 1) * kPage)); page_done[pg] = 1; }
    }
    void ensure_all() {
        if (!wire || all_done) return;
        std::lock_guard<std::mutex> lk(mu);
        if (all_done) return;
        if (v.cap < n) v.resize(n);
        if (page_done.empty()) page_done.assign((n + kPage - 1) / kPage, 0);
        const size_t npages = page_done.size();
        const unsigned nth = n > (1u << 20) ? std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
        auto job = [&](size_t p0, size_t p1) {
            for (size_t pg = p0; pg < p1; ++pg)
                if (!page_done[pg]) { expand(pg * kPage, std::min(n, (pg + 1) * kPage)); page_done[pg] = 1; }
        };
        if (nth <= 1) job(0, npages);
        else {
            std::vector<std::thread> th;
            for (unsigned k = 0; k < nth; ++k) th.emplace_back(job, npages * k / nth, npages * (k + 1) / nth);
            for (auto& t : th) t.join();
        }
        all_done = true;
    }
};

extern "C" {

int sx_last_error_code(void) { return g_err_code; }
const char* sx_last_error(void) { return g_err_msg.c_str(); }

int sx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

sx_scanner_state* sx_scanner_state_new(const sx_mission* m, int device) {
    sx_scanner_state* const fail = nullptr;
    if (!m) { set_err(SX_ERR_ARGUMENT, "mission is NULL"); return fail; }
    if (m->grep_char > 127) { set_err(SX_ERR_ARGUMENT, "grep_char must be an ASCII code (options.rs) or -1"); return fail; }
    if (m->output_line_char_nb_max < 6 || m->output_line_char_nb_max > 8192) { set_err(SX_ERR_UNSUPPORTED, "output_line_char_nb_max must be in 6..8192"); return fail; }
    if (m->chars_min_nb == 0) { set_err(SX_ERR_UNSUPPORTED, "chars_min_nb must be >= 1"); return fail; }
    if (m->encoding_id > SX_ENC_EUC_JP) { set_err(SX_ERR_ARGUMENT, "unknown encoding_id"); return fail; }
    int n = sx_device_count();
    if (n <= 0) { set_err(SX_ERR_NO_DEVICE, "no CUDA device: the scanner has no CPU fallback"); return fail; }
    if (device < 0 || device >= n) { set_err(SX_ERR_ARGUMENT, "device ordinal out of range"); return fail; }
    CK(cudaSetDevice(device));
    sx_scanner_state* ss = new sx_scanner_state();
    ss->m = *m;
    ss->device = device;
    ss->consumed = m->counter_offset;  // scanner.rs:86
    ss->cut = false;
    ss->npend = 0;
    memset(ss->pend, 0, sizeof ss->pend);
    memset(&ss->stats, 0, sizeof ss->stats);
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { delete ss; return fail; }
    ss->num_sms = prop.multiProcessorCount;
    bool ok = cuda_ok(cudaMalloc(&ss->d_counters, 8 * sizeof(unsigned long long)), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&ss->d_final, sizeof(FinalState)), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&ss->d_ctl, sizeof(PieceCtl) * kMaxPieces), "cudaMalloc") &&
              cuda_ok(cudaHostAlloc((void**)&ss->h_ctl, sizeof(HostCtl), cudaHostAllocPortable | cudaHostAllocMapped), "cudaHostAlloc");
    for (int i = 0; ok && i < 6; ++i) ok = cuda_ok(cudaEventCreate(&ss->ev[i]), "cudaEventCreate");
    ok = ok && cuda_ok(cudaEventCreateWithFlags(&ss->ev_in, cudaEventDisableTiming), "cudaEventCreate") &&
         cuda_ok(cudaEventCreateWithFlags(&ss->ev_done, cudaEventDisableTiming), "cudaEventCreate");
    // the exact stage and the copies outrank the streaming prefilter: their small kernels must not queue behind it
    int prio_lo = 0, prio_hi = 0;
    if (ok) ok = cuda_ok(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi), "cudaDeviceGetStreamPriorityRange");
    ok = ok && cuda_ok(cudaStreamCreateWithPriority(&ss->side, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate") &&
         cuda_ok(cudaStreamCreateWithPriority(&ss->sA, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate") &&
         cuda_ok(cudaStreamCreateWithPriority(&ss->sB, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate") &&
         cuda_ok(cudaStreamCreateWithPriority(&ss->sC, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate") &&
         cuda_ok(cudaEventCreateWithFlags(&ss->ev_setup, cudaEventDisableTiming), "cudaEventCreate");
    ok = ok && cuda_ok(cudaStreamCreateWithFlags(&ss->own, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; ok && i < sx_scanner_state::kLanes; ++i)
        ok = cuda_ok(cudaStreamCreateWithPriority(&ss->sBk[i], cudaStreamNonBlocking, prio_hi), "cudaStreamCreate") &&
             cuda_ok(cudaStreamCreateWithPriority(&ss->sidek[i], cudaStreamNonBlocking, prio_hi), "cudaStreamCreate");
    if (!ok) { sx_scanner_state_free(ss); return fail; }
    return ss;
}

void sx_scanner_state_free(sx_scanner_state* ss) {
    if (!ss) return;
    if (ss->worker) {  // pending asynchronous scans finish first
        {
            std::lock_guard<std::mutex> lk(ss->worker->mu);
            ss->worker->stop = true;
        }
        ss->worker->cv.notify_all();
        if (ss->worker->th.joinable()) ss->worker->th.join();
        delete ss->worker;
        ss->worker = nullptr;
    }
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ss->device);
    cudaFree(ss->d_in); cudaFree(ss->d_recs); cudaFree(ss->d_text); cudaFree(ss->d_blocks); cudaFree(ss->d_ccount); cudaFree(ss->d_coff); cudaFree(ss->d_list);
    cudaFree(ss->d_counters); cudaFree(ss->d_final); cudaFree(ss->d_ctl);
    cudaFree(ss->d_entries); cudaFree(ss->d_btot); cudaFree(ss->d_tables); cudaFree(ss->d_queue); cudaFree(ss->d_bpos);
    cudaFree(ss->d_clist); cudaFree(ss->d_findings); cudaFree(ss->d_pagebase);
    cudaFreeHost(ss->h_recs); cudaFreeHost(ss->h_text); cudaFreeHost(ss->h_blocks); cudaFreeHost(ss->h_ctl);
    for (auto e : ss->ev) if (e) cudaEventDestroy(e);
    if (ss->ev_in) cudaEventDestroy(ss->ev_in);
    if (ss->ev_done) cudaEventDestroy(ss->ev_done);
    for (auto& pe : ss->pev) {
        for (cudaEvent_t e : {pe.p0, pe.p1, pe.b0, pe.b1, pe.sc}) if (e) cudaEventDestroy(e);
        for (auto e : pe.sev) if (e) cudaEventDestroy(e);
        for (auto e : pe.sord) if (e) cudaEventDestroy(e);
        for (auto e : pe.gp) if (e) cudaEventDestroy(e);
    }
    for (cudaStream_t q : {ss->side, ss->sA, ss->sB, ss->sC, ss->own}) if (q) cudaStreamDestroy(q);
    for (cudaStream_t q : ss->sBk) if (q) cudaStreamDestroy(q);
    for (cudaStream_t q : ss->sidek) if (q) cudaStreamDestroy(q);
    if (ss->ev_setup) cudaEventDestroy(ss->ev_setup);
    delete ss;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
}

void sx_scanner_state_reset(sx_scanner_state* ss) {
    ss->consumed = ss->m.counter_offset;
    ss->cut = false;
    ss->leftover.clear();
    ss->npend = 0;
    memset(ss->pend, 0, sizeof ss->pend);
}
uint64_t sx_scanner_state_consumed_bytes(const sx_scanner_state* ss) { return ss->consumed; }
int sx_scanner_state_maybe_cut(const sx_scanner_state* ss) { return ss->cut ? 1 : 0; }
size_t sx_scanner_state_leftover(const sx_scanner_state* ss, const uint8_t** p) {
    *p = ss->leftover.data();
    return ss->leftover.size();
}
void sx_scanner_state_last_stats(const sx_scanner_state* ss, sx_scan_stats* out) { *out = ss->stats; }
void sx_scanner_state_set_prefilter(sx_scanner_state* ss, int enabled) { ss->use_prefilter = enabled ? 1 : 0; }
void sx_scanner_state_set_tma(sx_scanner_state* ss, int enabled) { ss->use_tma = enabled ? 1 : 0; }
void sx_scanner_state_set_sparse(sx_scanner_state* ss, int mode) { ss->use_sparse = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
void sx_scanner_state_set_direct_output(sx_scanner_state* ss, int enabled) { ss->use_direct = enabled ? 1 : 0; }
size_t sx_scanner_state_last_window_list(const sx_scanner_state* ss, uint32_t* out, size_t cap) {
    if (!ss->stats.prefilter_used) return 0;
    const size_t n = (size_t)ss->stats.windows_listed;
    const size_t k = n < cap ? n : cap;
    if (k && out) {
        cudaSetDevice(ss->device);
        size_t done = 0;
        if (ss->last_list_compact) {  // sparse pipeline: the pieces' compact lists
            for (const auto& pl : ss->last_piece_lists) {
                const size_t c = std::min<size_t>(pl.second, k - done);
                if (c && cudaMemcpy(out + done, ss->d_clist + pl.first, c * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
                done += c;
            }
            return n;
        }
        std::vector<uint32_t> off(ss->last_ncta + 1);
        if (cudaMemcpy(off.data(), ss->d_coff, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
        for (uint32_t b = 0; b < ss->last_ncta && done < k; ++b) {
            const size_t c = std::min<size_t>(off[b + 1] - off[b], k - done);
            if (c && cudaMemcpy(out + done, ss->last_list_base + (size_t)b * ss->last_region_stride, c * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
            done += c;
        }
    }
    return n;
}
void sx_scanner_state_set_pieces(sx_scanner_state* ss, int pieces) { ss->pieces = pieces < 0 ? 0 : (pieces > kMaxPieces ? kMaxPieces : pieces); }

size_t sx_fc_len(const sx_finding_collection* fc) { return fc->wire ? fc->n : fc->v.size(); }
const sx_finding* sx_fc_get(const sx_finding_collection* fc, size_t i) {
    const_cast<sx_finding_collection*>(fc)->ensure_page(i);
    return &fc->v[i];
}
const sx_finding* sx_fc_data(const sx_finding_collection* fc) {
    const_cast<sx_finding_collection*>(fc)->ensure_all();
    return fc->v.data();
}
uint64_t sx_fc_first_byte_position(const sx_finding_collection* fc) { return fc->first_byte_position; }
int sx_fc_str_buf_overflow(const sx_finding_collection* fc) { return fc->str_buf_overflow; }
void sx_fc_free(sx_finding_collection* fc) { delete fc; }

}  // extern "C"

template <class T>
static bool grow_pinned(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return true;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 256;
    if (!cuda_ok(cudaMallocHost(p, want * sizeof(T)), "cudaMallocHost")) return false;
    *cap = want;
    return true;
}

template <class T>
static bool grow(T** p, size_t* cap, size_t need, bool slack = true) {
    if (need <= *cap) return true;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = need + (slack ? std::min<size_t>(need / 4, (1ull << 30) / sizeof(T)) : 0) + 256;  // slack capped at 1 GiB
    if (!cuda_ok(cudaMalloc(p, want * sizeof(T)), "cudaMalloc")) return false;
    *cap = want;
    return true;
}

static PrefK make_pref_k(const ScanParams& P, const PrefCfg& c) {
    PrefK k;
    memset(&k, 0, sizeof k);
    const uint32_t lowsrc = c.family == PF_UNIT ? c.blkH : c.blkA;
    for (int b = 0; b < 4; ++b) {
        k.ka[b] = ((lowsrc >> b) & 1u) ? 0xFFFFFFFFu : 0u;
        k.kh[b] = ((c.blkH >> (4 + b)) & 1u) ? 0xFFFFFFFFu : 0u;
    }
    k.multi = c.multi ? 0xFFFFFFFFu : 0u;
    if (c.family == PF_UNIT) {
        // window starts are multiples of 16, so byte i of any window sits at (i - align) mod unit of its unit
        k.spread_left = c.hi_pos != 0;  // little endian: the tested byte is the unit's last byte
        for (uint32_t i = 0; i < 128; ++i) {
            const int64_t rel = (int64_t)i - (int64_t)P.align;
            const int64_t u0 = rel >= 0 ? (rel / c.unit) * c.unit : -(((-rel) + c.unit - 1) / c.unit) * (int64_t)c.unit;
            const int64_t t = (int64_t)P.align + u0 + c.hi_pos;  // tested byte of the unit holding byte i
            if (t == (int64_t)i) k.hi_mask[i >> 5] |= 1u << (i & 31);
            if (t < 0 || t >= (int64_t)P.W) k.edge_mask[i >> 5] |= 1u << (i & 31);
        }
    }
    return k;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2D view of the input as rows of 128 bytes; one box = one prefilter tile, 128-byte swizzle.
static bool make_tensor_map(CUtensorMap* tm, const uint8_t* d_in, size_t len, uint32_t tile_bytes) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
        fn = (PFN_encodeTiled)p;
    }
    const cuuint64_t rows = (cuuint64_t)(len >> 7);
    if (rows == 0 || (reinterpret_cast<uintptr_t>(d_in) & 15u)) return false;
    const cuuint64_t gdim[2] = {128, rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {128, tile_bytes >> 7};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_in), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int FAMILY, bool DEF, bool FAST>
static cudaError_t launch_prefilter_t(const ScanParams& P, const PrefCfg& c, const PrefK& k, const PrefOut& o, long long total_windows,
                                      int grid, cudaStream_t st, const CUtensorMap& tm, uint32_t use_tma) {
    // the attribute is per device and per instantiation; set once per device (any ordinal), process wide
    static std::mutex mu;
    static std::vector<char> attr_done;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(mu);
        if ((size_t)dev >= attr_done.size()) attr_done.resize((size_t)dev + 1, 0);
        if (!attr_done[dev]) {
            cudaError_t e = cudaFuncSetAttribute(sx_prefilter_kernel<FAMILY, DEF, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrefSmemBytes);
            if (e != cudaSuccess) return e;
            attr_done[dev] = 1;
        }
    }
    sx_prefilter_kernel<FAMILY, DEF, FAST><<<grid, kPrefThreads, kPrefSmemBytes, st>>>(P, c, k, o, total_windows, tm, use_tma);
    return cudaGetLastError();
}

static cudaError_t launch_prefilter(const ScanParams& P, const PrefCfg& c, const PrefK& k, const PrefOut& o, long long total_windows,
                                    int grid, cudaStream_t st, const CUtensorMap& tm, uint32_t use_tma) {
    const bool defshape = c.family == PF_UTF8 && c.blkA == 0xEu && c.blkH == (1u << 6) && !c.multi;
    const bool fast = P.W == 128 && c.T <= 32 && c.kill_trail == 0 && c.sb_rule == 0;  // the bit-plane path keeps only T - 1 flags of the previous window
#define SX_PREF(F, D) (fast ? launch_prefilter_t<F, D, true>(P, c, k, o, total_windows, grid, st, tm, use_tma) \
                            : launch_prefilter_t<F, D, false>(P, c, k, o, total_windows, grid, st, tm, use_tma))
    switch (c.family) {
    case PF_BYTE: return SX_PREF(PF_BYTE, false);
    case PF_UTF8: return defshape ? SX_PREF(PF_UTF8, true) : SX_PREF(PF_UTF8, false);
    case PF_PAIR: return launch_prefilter_t<PF_PAIR, false, false>(P, c, k, o, total_windows, grid, st, tm, use_tma);
    default: return SX_PREF(PF_UNIT, false);
    }
#undef SX_PREF
}

static size_t utf8_char_count(const std::vector<uint8_t>& s) {
    size_t n = 0;
    for (uint8_t b : s) n += (b & 0xC0) != 0x80;
    return n;
}

// =============================================================================================
// sx_scan_stream / sx_scan_range
// =============================================================================================
struct CallCtx {
    sx_scanner_state* ss;
    sx_finding_collection* fc;
    int input_file_id;
    const uint8_t* d_in;
    size_t len;
    cudaStream_t st;
    ScanParams P;
    PrefCfg pc;
    long long total_windows, w_lo, w_hi;  // the call's window range
    bool in_aligned16, prefix_known, buf_is_device;
    std::chrono::steady_clock::time_point t_begin;
    // results of the device part
    size_t nrec = 0, ntext = 0;
    unsigned long long windows_listed = 0;
    FinalState fin;
    bool direct_out = false;   // findings and text are complete in fc->set
    int wire = 0;              // ... as 8-byte (per-stage pipeline) or 16-byte (block path) wire records
    bool records_in_order = false;  // (legacy download) d_recs is in stream order: one descriptor
    uint8_t tail[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t tail_n = 0;
};
static float ms_since(const std::chrono::steady_clock::time_point& t0) {
    return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static cudaError_t launch_exact_enc(const ScanParams& P, const ScanOut& O, const ExactCfg& X, unsigned grid, cudaStream_t st) {
    switch (P.enc) {
    case 0: return launch_exact_0(P, O, X, grid, st);
    case 1: return launch_exact_1(P, O, X, grid, st);
    case 2: return launch_exact_2(P, O, X, grid, st);
    case 3: return launch_exact_3(P, O, X, grid, st);
    case 4: return launch_exact_4(P, O, X, grid, st);
    case 5: return launch_exact_5(P, O, X, grid, st);
    case 6: return launch_exact_6(P, O, X, grid, st);
    case 7: return launch_exact_7(P, O, X, grid, st);
    case 8: return launch_exact_8(P, O, X, grid, st);
    }
    return cudaErrorInvalidValue;
}
static cudaError_t launch_range_carry_enc(const ScanParams& P, const RangeCarryArgs& A, cudaStream_t st) {
    switch (P.enc) {
    case 0: return launch_range_carry_0(P, A, st);
    case 1: return launch_range_carry_1(P, A, st);
    case 2: return launch_range_carry_2(P, A, st);
    case 3: return launch_range_carry_3(P, A, st);
    case 4: return launch_range_carry_4(P, A, st);
    case 5: return launch_range_carry_5(P, A, st);
    case 6: return launch_range_carry_6(P, A, st);
    case 7: return launch_range_carry_7(P, A, st);
    case 8: return launch_range_carry_8(P, A, st);
    }
    return cudaErrorInvalidValue;
}
static cudaError_t launch_sparse_enc(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B, const SparseLaunchCfg& L,
                                     cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs) {
    switch (P.enc) {
    case 0: return launch_sparse_0(P, O, X, B, L, st, ev, side, evs);
    case 1: return launch_sparse_1(P, O, X, B, L, st, ev, side, evs);
    case 2: return launch_sparse_2(P, O, X, B, L, st, ev, side, evs);
    case 3: return launch_sparse_3(P, O, X, B, L, st, ev, side, evs);
    case 4: return launch_sparse_4(P, O, X, B, L, st, ev, side, evs);
    case 5: return launch_sparse_5(P, O, X, B, L, st, ev, side, evs);
    case 6: return launch_sparse_6(P, O, X, B, L, st, ev, side, evs);
    case 7: return launch_sparse_7(P, O, X, B, L, st, ev, side, evs);
    case 8: return launch_sparse_8(P, O, X, B, L, st, ev, side, evs);
    }
    return cudaErrorNotSupported;
}
// Every decoder can take the per-stage pipeline (those without a mask engine run the byte-wise engine in it);
// SX_SPARSE_ALL=0 restricts it to the mask-engine families again (the others then take the block kernel).
static bool has_sparse_enc(uint32_t enc) {
    static int all = -1;
    if (all < 0) { const char* ev = getenv("SX_SPARSE_ALL"); all = ev ? atoi(ev) : 1; }
    return all || enc == ENC_XUD || enc == ENC_UTF8 || enc == ENC_SB;
}

static bool ensure_piece_events(sx_scanner_state* ss, size_t n) {
    while (ss->pev.size() < n) {
        PieceEvents e;
        bool ok = cuda_ok(cudaEventCreate(&e.p0), "cudaEventCreate") && cuda_ok(cudaEventCreate(&e.p1), "cudaEventCreate") &&
                  cuda_ok(cudaEventCreate(&e.b0), "cudaEventCreate") && cuda_ok(cudaEventCreate(&e.b1), "cudaEventCreate");
        for (int i = 0; ok && i < 7; ++i) ok = cuda_ok(cudaEventCreate(&e.sev[i]), "cudaEventCreate");
        for (int i = 0; ok && i < 2; ++i) ok = cuda_ok(cudaEventCreateWithFlags(&e.sord[i], cudaEventDisableTiming), "cudaEventCreate");
        ok = ok && cuda_ok(cudaEventCreateWithFlags(&e.sc, cudaEventDisableTiming), "cudaEventCreate");
        for (int i = 0; ok && i < kGatherParts; ++i) ok = cuda_ok(cudaEventCreateWithFlags(&e.gp[i], cudaEventDisableTiming), "cudaEventCreate");
        ss->pev.push_back(e);
        if (!ok) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Sparse pipeline over the call's window range, cut into pieces: the prefilter of piece k + 1 streams (stream sA) while
// the exact stage of piece k resolves (sB) and the findings of piece k - 1 travel to the host (copy engine, sC).
// Returns 1 on success, 0 on error (sx_last_error set).
// ---------------------------------------------------------------------------------------------
static int run_sparse(CallCtx& c) {
    const int fail = 0;
    sx_scanner_state* const ss = c.ss;
    sx_finding_collection* const fc = c.fc;
    const ScanParams& P = c.P;
    const long long nwin = c.w_hi - c.w_lo;
    const long long tile_lo = c.w_lo / kPrefTileWin, tile_hi = (c.w_hi + kPrefTileWin - 1) / kPrefTileWin;
    const long long ntiles = tile_hi - tile_lo;
    // ---- piece plan -------------------------------------------------------------------------------
    // ONE prefilter launch streams the whole range (alone on the machine: anything that shares its SMs costs it 25-45 %,
    // profiles/r02_pieces.txt); the exact stage is cut into K pieces = groups of consecutive prefilter CTAs, i.e. window
    // ranges whose first window is force-listed and whose carry-in comes from sx_range_carry_kernel.  The pieces' exact
    // stages run on `lanes` streams: the SM-bound heads of piece k + 1 beside the latency-bound members / walks of piece k
    // and the PCIe-bound gather of piece k - 1.  Output-heavy calls (koi8-r on random bytes: 115 M findings per 4 GiB):
    // more pieces on ONE lane, so that piece k travels to the host (copy engine) while piece k + 1 resolves.
    int K = ss->pieces;
    if (const char* ev = getenv("SX_PIECES")) K = atoi(ev);
    int lanes = 2;
    bool heavy = false;
    if (ss->have_history && ss->last_len) {
        const double scale = (double)(nwin * P.W) / (double)ss->last_len;
        const double pcie_ms = ((double)ss->last_nrec * 8 + (double)ss->last_ntext) * scale / 50e6;  // ~50 GB/s
        heavy = pcie_ms > 8.0;
        if (heavy && K <= 0) K = (int)std::min(8.0, std::ceil(pcie_ms / 8.0) + 1.0);
    }
    if (heavy) lanes = 1;
    // Light output: ONE piece.  Measured (profiles/r02_pieces.txt): exact-stage pieces side by side on 2-4 lanes run in lockstep
    // and share the SMs (K = 2: 1.32 ms, 4: 1.41, 8: 1.67 vs 1.29 for K = 1): the latency-bound stages of one piece slow down
    // 3x beside the SM-bound heads of another, which eats what the overlap would give.
    if (K <= 0) K = 1;
    if (const char* ev = getenv("SX_LANES")) lanes = std::max(1, std::min((int)sx_scanner_state::kLanes, atoi(ev)));
    if (K > kMaxPieces) K = kMaxPieces;
    int pgrid_max = ss->num_sms * 3;
    if (const char* ev = getenv("SX_PREF_GRID")) { const int g = atoi(ev); if (g > 0) pgrid_max = g; }
    pgrid_max = std::min(pgrid_max, kMaxPrefCtas);
    if ((long long)K > ntiles) K = (int)std::max<long long>(1, ntiles);
    if (K > pgrid_max) K = pgrid_max;
    // prefilter CTAs: cpp per piece, tpc tiles each
    int cpp = std::max(1, pgrid_max / K);
    if ((long long)cpp * K > ntiles) cpp = (int)std::max<long long>(1, ntiles / K);
    const long long tpc = (ntiles + (long long)cpp * K - 1) / ((long long)cpp * K);
    const int pgrid_all = (int)((ntiles + tpc - 1) / tpc);  // <= cpp * K
    K = (pgrid_all + cpp - 1) / cpp;
    struct Piece { long long w0, w1, t0, t1; unsigned long long cap, ebase, cbase; int pgrid; long long tpc; };
    std::vector<Piece> pcs(K);
    std::vector<unsigned long long> want_cap(K, 0);  // after an overflow: the counted entries
    std::vector<uint32_t> gparts(K, 1);
    if (!ensure_piece_events(ss, (size_t)K)) return fail;
    {   // same shared-memory carve-out as the prefilter, so that these kernels can share SMs with it (sx_sparse_utf8.cuh)
        static thread_local int done_dev = -1;
        if (done_dev != ss->device) {
            cudaFuncSetAttribute(sx_list_compact_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(sx_sp_tables_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(sx_materialize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            done_dev = ss->device;
        }
    }
    if (!grow(&ss->d_ccount, &ss->ccount_cap, (size_t)kMaxPieces * kMaxPrefCtas)) return fail;
    if (!grow(&ss->d_list, &ss->list_cap, (size_t)(c.total_windows + 2 * kPrefTileWin))) return fail;
    if (!grow(&ss->d_tables, &ss->tables_cap, sizeof(Utf8Tables))) return fail;

    size_t need_recs = (size_t)((double)(nwin * P.W) * ss->rec_per_byte) + 4096;
    size_t need_text = (size_t)((double)(nwin * P.W) * ss->text_per_byte) + 65536;
    size_t need_recs_min = 0, need_text_min = 0;
    HostCtl* const hc = ss->h_ctl;

    for (int attempt = 0;; ++attempt) {
        // ---- capacities -----------------------------------------------------------------------------
        unsigned long long etot = 0, ctot = 0;
        for (int k = 0; k < K; ++k) {
            Piece& p = pcs[k];
            p.t0 = tile_lo + (long long)k * cpp * tpc;
            p.t1 = std::min<long long>(tile_hi, p.t0 + (long long)cpp * tpc);
            p.w0 = std::max<long long>(c.w_lo, p.t0 * kPrefTileWin);
            p.w1 = std::min<long long>(c.w_hi, p.t1 * kPrefTileWin);
            const unsigned long long wins = (unsigned long long)(p.w1 - p.w0);
            unsigned long long cap = wins;
            if (nwin > (256ll << 10)) cap = std::min<unsigned long long>(wins, (unsigned long long)((double)wins * ss->listed_frac * 1.5) + 16384);
            if (want_cap[k]) cap = std::min<unsigned long long>(wins, want_cap[k] + 1024);
            p.cap = cap; p.ebase = etot; p.cbase = ctot;
            etot += cap;
            ctot += (cap + kSpThreads - 1) / kSpThreads + 1;
            p.tpc = tpc;
            p.pgrid = std::min(cpp, pgrid_all - k * cpp);  // prefilter CTAs (list regions) of this piece
        }
        const size_t entry_bytes = sizeof(EntryHot) + sizeof(Carry) + 2 * kBufRecs * sizeof(Record);
        if (!grow(&ss->d_entries, &ss->entries_cap, (size_t)etot * entry_bytes + 256, etot * entry_bytes <= (24ull << 30))) return fail;
        if (!grow(&ss->d_btot, &ss->btot_cap, (size_t)(ctot + 1) * sizeof(ulonglong2))) return fail;
        if (!grow(&ss->d_queue, &ss->queue_cap, (size_t)(2 * etot + 64 * K + 128))) return fail;
        if (!grow(&ss->d_clist, &ss->clist_cap, (size_t)etot + 64)) return fail;
        if (!grow(&ss->d_recs, &ss->rec_cap, need_recs)) return fail;
        if (!grow(&ss->d_text, &ss->text_cap, need_text)) return fail;
        EntryHot* const d_hot = reinterpret_cast<EntryHot*>(ss->d_entries);
        Carry* const d_dnull = reinterpret_cast<Carry*>(ss->d_entries + (size_t)etot * sizeof(EntryHot));
        Record* const d_staged = reinterpret_cast<Record*>(ss->d_entries + (size_t)etot * (sizeof(EntryHot) + sizeof(Carry)));
        Record* const d_xstaged = d_staged + (size_t)etot * kBufRecs;
        // ---- output set -----------------------------------------------------------------------------
        unsigned long long out_cap = ~0ull, text_cap = ss->text_cap;
        if (ss->use_direct) {
            const double scale = ss->have_history && ss->last_len ? (double)(nwin * P.W) / (double)ss->last_len : 1.0;
            const size_t guess_f = ss->have_history ? (size_t)(ss->last_nrec * scale) + ss->last_nrec / 16 + 4096 : ss->host_rec_hint;
            const size_t guess_t = ss->have_history ? (size_t)(ss->last_ntext * scale) + ss->last_ntext / 16 + 65536 : ss->host_text_hint;
            const size_t want_f = std::min(need_recs, std::max(guess_f, need_recs_min));
            const size_t want_t = std::min(need_text, std::max(guess_t, need_text_min));
            if (fc->set.fcap < wire8_bytes(want_f) || fc->set.tcap < want_t) {
                pinned_release(fc->set);
                if (!pinned_acquire(wire8_bytes(want_f), want_t, &fc->set)) { set_err(SX_ERR_CUDA, "cannot pin host memory for the findings"); return fail; }
            }
            out_cap = wire8_cap(fc->set.fcap);
            fc->wcap = (size_t)out_cap;
            if (heavy || getenv("SX_NO_ZERO_COPY")) {
                if (!grow(&ss->d_findings, &ss->findings_cap, (size_t)out_cap)) return fail;
                if (!grow(&ss->d_pagebase, &ss->pagebase_cap, (size_t)out_cap / kWirePage + 2)) return fail;
            }
            text_cap = std::min<unsigned long long>(text_cap, fc->set.tcap);
        }
        // ---- enqueue --------------------------------------------------------------------------------
        CK(cudaEventRecord(ss->ev_in, c.st));
        CK(cudaStreamWaitEvent(ss->sA, ss->ev_in, 0));
        CK(cudaStreamWaitEvent(ss->sB, ss->ev_in, 0));
        CK(cudaStreamWaitEvent(ss->sC, ss->ev_in, 0));
        CK(cudaMemsetAsync(ss->d_ctl, 0, sizeof(PieceCtl) * kMaxPieces, ss->sB));
        CK(cudaMemsetAsync(ss->d_counters, 0, 8 * sizeof(unsigned long long), ss->sB));  // [7]: Big5 / EUC-JP walk bound flag
        memset(hc, 0, sizeof(HostCtl));
        {
            RangeCarryArgs ra;
            memset(&ra, 0, sizeof ra);
            for (int k = 0; k < K; ++k) ra.w_first[k] = pcs[k].w0;
            ra.nranges = (uint32_t)K; ra.in_aligned16 = c.in_aligned16 ? 1u : 0u;
            ra.prefix_known = c.prefix_known ? 1u : 0u; ra.max_back = 1u << 16;
            ra.out = reinterpret_cast<RangeCarryOut*>(reinterpret_cast<uint8_t*>(ss->d_ctl) + offsetof(PieceCtl, rc));
            ra.out_stride = sizeof(PieceCtl);
            CK(launch_range_carry_enc(P, ra, ss->sB));
            sx_sp_tables_kernel<<<8, 256, 0, ss->sB>>>(P, reinterpret_cast<Utf8Tables*>(ss->d_tables));
            CK(cudaGetLastError());
            ss->stats.kernel_launches += 2;
            CK(cudaEventRecord(ss->ev_setup, ss->sB));
            for (int i = 0; i < sx_scanner_state::kLanes; ++i) CK(cudaStreamWaitEvent(ss->sBk[i], ss->ev_setup, 0));
        }
        const PrefK pk = make_pref_k(P, c.pc);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof tmap);
        const uint32_t use_tma = (ss->use_tma && make_tensor_map(&tmap, c.d_in, c.len, kPrefTileWin * P.W)) ? 1u : 0u;
        ss->stats.tma_used = use_tma;
        {
            PrefOut po;
            po.list = ss->d_list; po.cta_count = ss->d_ccount; po.tiles_per_cta = tpc;
            po.tile0 = tile_lo; po.tile_end = tile_hi; po.w_first = c.w_lo; po.w_lo = c.w_lo; po.w_hi = c.w_hi;
            po.piece_ctas = K > 1 ? (uint32_t)cpp : 0u;
            CK(cudaEventRecord(ss->pev[0].p0, ss->sA));
            CK(launch_prefilter(P, c.pc, pk, po, c.total_windows, pgrid_all, ss->sA, tmap, use_tma));
            CK(cudaEventRecord(ss->pev[0].p1, ss->sA));
            ss->stats.kernel_launches++;
        }
        const unsigned chunk_grid_max = (unsigned)ss->num_sms * 6u, queue_grid_max = (unsigned)ss->num_sms * 4u;
        for (int k = 0; k < K; ++k) {
            const Piece& p = pcs[k];
            PieceCtl* const ctl = ss->d_ctl + k;
            cudaStream_t sb = ss->sBk[k % lanes], sd = ss->sidek[k % lanes];
            CK(cudaStreamWaitEvent(sb, ss->pev[0].p1, 0));
            CK(cudaEventRecord(ss->pev[k].b0, sb));
            uint32_t* const clist = ss->d_clist + p.ebase;
            sx_list_compact_kernel<<<p.pgrid, 256, 0, sb>>>(ss->d_ccount + (size_t)k * cpp, (uint32_t)p.pgrid, ss->d_list, p.t0, p.tpc,
                                                               clist, ctl, p.cap);
            CK(cudaGetLastError());
            // Where the gather writes: one piece (little output) -> straight into the collection's pinned, device-mapped set
            // (posted PCIe writes, staged in shared memory so they leave as large transactions: nothing is left to download
            // when the kernel ends); several pieces (output-heavy) -> device staging, moved by the copy engine while the next
            // piece resolves.
            const bool zero_copy = ss->use_direct && !heavy && !getenv("SX_NO_ZERO_COPY");
            ScanOut O{ss->d_recs, ss->rec_cap, text_cap, ss->d_blocks, ss->d_counters, &hc->fin,
                      !ss->use_direct ? nullptr : zero_copy ? reinterpret_cast<uint4*>(fc->set.f) : reinterpret_cast<uint4*>(ss->d_findings),
                      out_cap,
                      !ss->use_direct ? nullptr : zero_copy ? reinterpret_cast<unsigned long long*>(fc->set.f + fc->wcap * 8) : ss->d_pagebase,
                      c.input_file_id, ss->m.mission_id};
            ExactCfg X;
            X.list = clist; X.ne_ptr = &ctl->ne; X.ne_static = 0; X.total_windows = c.total_windows;
            X.in_aligned16 = c.in_aligned16 ? 1u : 0u; X.pre_bytes = c.pc.pre_bytes; X.cta_off = nullptr; X.ncta = 0; X.region_stride = 0;
            X.w_first = p.w0; X.w_end = p.w1; X.k0_ptr = &ctl->rc.k0; X.ne_cap = p.cap;
            SparseBufs B;
            B.H = d_hot + p.ebase; B.dnull = d_dnull + p.ebase; B.staged = d_staged + p.ebase * kBufRecs; B.xstaged = d_xstaged + p.ebase * kBufRecs;
            B.btot = reinterpret_cast<ulonglong2*>(ss->d_btot) + p.cbase;
            B.tables = reinterpret_cast<Utf8Tables*>(ss->d_tables);
            B.queue = ss->d_queue + 2 * p.ebase + 64 * (size_t)k; B.queue2 = B.queue + p.cap + 32;
            B.ctl = ctl; B.prev = k ? ctl - 1 : nullptr; B.summary = &hc->piece[k];
            B.text_out = zero_copy ? fc->set.t : ss->d_text;
            SparseLaunchCfg L;
            // one chunk per CTA (CTAs beyond the entry count, which only the device knows, leave at once): measured faster
            // than a persistent grid for these latency-bound kernels
            L.grid_chunks = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((p.cap + kSpThreads - 1) / kSpThreads, 1u << 30));
            (void)chunk_grid_max;
            L.grid_queue = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((p.cap + kSpThreads - 1) / kSpThreads, queue_grid_max));
            L.rec_cap = ss->rec_cap; L.text_cap = text_cap; L.out_cap = out_cap;
            L.ev_scan_prev = k ? ss->pev[k - 1].sc : nullptr; L.ev_scan_done = ss->pev[k].sc;
            // large pieces: the gather runs in parts, the download of a part beside the gathering of the next
            L.gather_parts = (p.cap >= (64u << 10) && !zero_copy) ? 2u : 1u;  // staging: 2 parts 1.43 ms, 4 parts 1.44, 1 part 1.47 per 4 GiB
            if (const char* evp = getenv("SX_GATHER_PARTS")) L.gather_parts = (uint32_t)std::min(kGatherParts, std::max(1, atoi(evp)));
            L.ev_part = ss->pev[k].gp;
            gparts[k] = L.gather_parts;
            CK(launch_sparse_enc(P, O, X, B, L, sb, ss->pev[k].sev, sd, ss->pev[k].sord));
            CK(cudaEventRecord(ss->pev[k].b1, sb));
            ss->stats.kernel_launches += 1 + kSparseLaunches;
        }
        if (c.tail_n && c.buf_is_device)
            CK(cudaMemcpyAsync(hc->tail + 8 - c.tail_n, c.d_in + c.len - c.tail_n, c.tail_n, cudaMemcpyDeviceToHost, ss->sC));
        ss->stats.host_phase_ms[0] = ms_since(c.t_begin);
        // ---- per piece: wait for its exact stage, copy its findings and text to the pinned set ---------------
        bool overflow = false;
        unsigned long long tot_rec = 0, tot_text = 0, tot_listed = 0;
        for (int k = 0; k < K; ++k) {
            CK(cudaEventSynchronize(ss->pev[k].sc));  // the piece's totals are known (its gather may still be running)
            const PieceSummary s = hc->piece[k];
            tot_listed += s.ne_raw;
            if (s.overflow & 0x100u) { set_err(SX_ERR_UNSUPPORTED, "range start: no carry-independent window found (halo too short?)"); cudaDeviceSynchronize(); return fail; }
            if (s.overflow) overflow = true;
            if (overflow) { CK(cudaEventSynchronize(ss->pev[k].b1)); continue; }
            tot_rec = s.rec_base + s.nrec; tot_text = s.text_base + s.ntext;
            unsigned long long r0 = s.rec_base, t0 = s.text_base;
            for (uint32_t part = 0; part < gparts[k]; ++part) {
                CK(cudaEventSynchronize(ss->pev[k].gp[part]));
                const unsigned long long r1 = part + 1 == gparts[k] ? s.rec_base + s.nrec : s.part_rec_end[part];
                const unsigned long long t1 = part + 1 == gparts[k] ? s.text_base + s.ntext : s.part_text_end[part];
                const bool zc = ss->use_direct && !heavy && !getenv("SX_NO_ZERO_COPY");
                const bool fb = hc->piece[k].text_fallback != 0;
                if (fb && r1 > r0) {
                    const int mgrid = (int)std::min<size_t>((size_t)(r1 - r0 + 255) / 256, (size_t)ss->num_sms * 8);
                    sx_materialize_kernel<<<mgrid, 256, 0, ss->sC>>>(P, ss->d_recs + r0, r1 - r0, ss->d_text, ss->text_cap);
                    CK(cudaGetLastError());
                    ss->stats.kernel_launches++;
                }
                if (ss->use_direct) {
                    if (!zc && r1 > r0) {
                        CK(cudaMemcpyAsync(fc->set.f + r0 * 8, ss->d_findings + r0, (size_t)(r1 - r0) * 8, cudaMemcpyDeviceToHost, ss->sC));
                        const unsigned long long p0 = (r0 + kWirePage - 1) / kWirePage, p1 = (r1 + kWirePage - 1) / kWirePage;  // pages that start here
                        if (p1 > p0) CK(cudaMemcpyAsync(fc->set.f + fc->wcap * 8 + p0 * 8, ss->d_pagebase + p0, (size_t)(p1 - p0) * 8, cudaMemcpyDeviceToHost, ss->sC));
                    }
                    // zero copy: the text is already there, unless some chunk left it to the materialize kernel
                    if ((!zc || fb) && t1 > t0) CK(cudaMemcpyAsync(fc->set.t + t0, ss->d_text + t0, (size_t)(t1 - t0), cudaMemcpyDeviceToHost, ss->sC));
                    ss->stats.d2h_bytes += (r1 - r0) * 8 + (t1 - t0);
                }
                r0 = r1; t0 = t1;
            }
        }
        ss->stats.host_phase_ms[1] = ms_since(c.t_begin);
        CK(cudaStreamSynchronize(ss->sC));
        CK(cudaStreamSynchronize(ss->sA));
        for (int i = 0; i < sx_scanner_state::kLanes; ++i) { CK(cudaStreamSynchronize(ss->sBk[i])); CK(cudaStreamSynchronize(ss->sidek[i])); }
        if (P.mb_fail) {
            unsigned long long flag = 0;
            CK(cudaMemcpy(&flag, ss->d_counters + 7, sizeof flag, cudaMemcpyDeviceToHost));
            if (flag) {
                set_err(SX_ERR_UNSUPPORTED, "Big5 / EUC-JP: more than 256 KiB of lead / trail bytes without a byte that resynchronises the decoder");
                return fail;
            }
        }
        if (!overflow) {
            c.nrec = (size_t)tot_rec; c.ntext = (size_t)tot_text; c.windows_listed = tot_listed;
            break;
        }
        if (attempt >= 3) { set_err(SX_ERR_CUDA, "output buffers still too small after regrowing"); return fail; }
        // what the device counted: entries per piece, records / text of the pieces that got that far
        unsigned long long cr = 0, ctx = 0;
        bool entries_ok = true;
        for (int k = 0; k < K; ++k) {
            const PieceSummary s = hc->piece[k];
            want_cap[k] = s.ne_raw;
            if (s.overflow & 1u) entries_ok = false;
            cr = std::max(cr, s.rec_base + s.nrec); ctx = std::max(ctx, s.text_base + s.ntext);
        }
        if (entries_ok) { need_recs = (size_t)cr + 1024; need_text = (size_t)ctx + 4096; }
        else { need_recs = std::max(need_recs, (size_t)cr * 2 + 1024); need_text = std::max(need_text, (size_t)ctx * 2 + 4096); }
        need_recs_min = need_recs; need_text_min = need_text;
        ss->stats.relaunches++;
    }
    // ---- stats ---------------------------------------------------------------------------------------
    {
        float pre = 0, ex = 0, stage[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < K; ++k) {
            float ms = 0;
            if (k == 0 && cudaEventElapsedTime(&ms, ss->pev[0].p0, ss->pev[0].p1) == cudaSuccess) pre += ms;
            if (cudaEventElapsedTime(&ms, ss->pev[k].b0, ss->pev[k].b1) == cudaSuccess) ex += ms;
            if (cudaEventElapsedTime(&ms, ss->pev[k].b0, ss->pev[k].sev[1]) == cudaSuccess) stage[0] += ms;  // compact
            for (int i = 1; i < 6; ++i)
                if (cudaEventElapsedTime(&ms, ss->pev[k].sev[i], ss->pev[k].sev[i + 1]) == cudaSuccess) stage[i] += ms;
        }
        cudaGetLastError();
        ss->stats.prefilter_kernel_ms = pre;
        ss->stats.exact_kernel_ms = ex;
        for (int i = 0; i < 6; ++i) ss->stats.sparse_stage_ms[i] = stage[i];
        float ms = 0;
        for (int k = 0; k < K; ++k)
            if (cudaEventElapsedTime(&ms, ss->pev[0].p0, ss->pev[k].b1) == cudaSuccess) ss->stats.scan_kernel_ms = std::max(ss->stats.scan_kernel_ms, ms);
        cudaGetLastError();
        ss->stats.pieces = (uint32_t)K;
    }
    ss->stats.sparse_used = 1;
    ss->stats.prefilter_used = 1;
    ss->listed_frac = std::max(1.0 / 4096, (double)c.windows_listed / (double)std::max<long long>(1, nwin));
    ss->last_list_compact = true; ss->last_list_len = (size_t)c.windows_listed;
    ss->last_piece_lists.clear();
    for (int k = 0; k < K; ++k) ss->last_piece_lists.emplace_back((size_t)pcs[k].ebase, (size_t)hc->piece[k].ne_raw);
    c.fin = hc->fin;
    if (c.tail_n && c.buf_is_device) memcpy(c.tail + 8 - c.tail_n, hc->tail + 8 - c.tail_n, c.tail_n);
    c.direct_out = ss->use_direct != 0;
    c.wire = 8;
    c.records_in_order = true;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Block-kernel path (UTF-16 / UTF-32, general missions, prefilter off): one prefilter launch over the range, the
// persistent block kernel over its list, records ordered on the device afterwards.  Everything on the caller's stream.
// ---------------------------------------------------------------------------------------------
static int run_block(CallCtx& c) {
    const int fail = 0;
    sx_scanner_state* const ss = c.ss;
    sx_finding_collection* const fc = c.fc;
    const ScanParams& P = c.P;
    const PrefCfg& pc = c.pc;
    cudaStream_t st = c.st;
    const long long nwin = c.w_hi - c.w_lo;
    const long long tile_lo = c.w_lo / kPrefTileWin, tile_hi = (c.w_hi + kPrefTileWin - 1) / kPrefTileWin;
    const long long ntiles = tile_hi - tile_lo;
    const long long max_blocks = (nwin + kThreads - 1) / kThreads;
    if (!grow(&ss->d_blocks, &ss->blocks_cap, (size_t)max_blocks + 1)) return fail;
    if (pc.enabled) {
        if (!grow(&ss->d_ccount, &ss->ccount_cap, (size_t)kMaxPieces * kMaxPrefCtas)) return fail;
        if (!grow(&ss->d_coff, &ss->coff_cap, (size_t)kMaxPrefCtas + 8)) return fail;
        if (!grow(&ss->d_list, &ss->list_cap, (size_t)(c.total_windows + 2 * kPrefTileWin))) return fail;
    }
    size_t need_recs = (size_t)((double)(nwin * P.W) * ss->rec_per_byte) + 4096;
    size_t need_text = (size_t)((double)(nwin * P.W) * ss->text_per_byte) + 65536;
    unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    FinalState fin;
    HostCtl* const hc = ss->h_ctl;
    for (int attempt = 0;; ++attempt) {
        if (!grow(&ss->d_recs, &ss->rec_cap, need_recs)) return fail;
        if (!grow(&ss->d_text, &ss->text_cap, need_text)) return fail;
        CK(cudaMemsetAsync(ss->d_counters, 0, 8 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(ss->d_final, 0, sizeof(FinalState), st));
        ScanOut O{ss->d_recs, ss->rec_cap, ss->text_cap, ss->d_blocks, ss->d_counters, ss->d_final, nullptr, 0, nullptr,
                  c.input_file_id, ss->m.mission_id};
        ExactCfg X;
        X.total_windows = c.total_windows;
        X.in_aligned16 = c.in_aligned16 ? 1u : 0u;
        X.pre_bytes = pc.pre_bytes;
        X.w_first = c.w_lo; X.w_end = c.w_hi; X.k0_ptr = nullptr; X.ne_cap = 0;
        if (c.w_lo > 0) {
            // the carry into the range's first window
            CK(cudaMemsetAsync(ss->d_ctl, 0, sizeof(PieceCtl), st));
            RangeCarryArgs ra;
            memset(&ra, 0, sizeof ra);
            ra.w_first[0] = c.w_lo;
            ra.nranges = 1; ra.in_aligned16 = X.in_aligned16; ra.prefix_known = c.prefix_known ? 1u : 0u;
            ra.max_back = 1u << 16;
            ra.out = reinterpret_cast<RangeCarryOut*>(reinterpret_cast<uint8_t*>(ss->d_ctl) + offsetof(PieceCtl, rc));
            ra.out_stride = sizeof(PieceCtl);
            CK(launch_range_carry_enc(P, ra, st));
            ss->stats.kernel_launches++;
            X.k0_ptr = &ss->d_ctl->rc.k0;
        }
        CK(cudaEventRecord(ss->ev[0], st));
        if (pc.enabled) {
            const PrefK pk = make_pref_k(P, pc);
            int pgrid = (int)std::min<long long>(ntiles, std::min<long long>(kMaxPrefCtas, (long long)ss->num_sms * 3));
            const long long tiles_per_cta = (ntiles + pgrid - 1) / pgrid;
            pgrid = (int)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
            PrefOut po;
            po.list = ss->d_list; po.cta_count = ss->d_ccount; po.tiles_per_cta = tiles_per_cta;
            po.tile0 = tile_lo; po.tile_end = tile_hi; po.w_first = c.w_lo; po.w_lo = c.w_lo; po.w_hi = c.w_hi;
            po.piece_ctas = 0;  // the block kernel takes the whole list at once
            CUtensorMap tmap;
            memset(&tmap, 0, sizeof tmap);
            const uint32_t use_tma = (ss->use_tma && make_tensor_map(&tmap, c.d_in, c.len, kPrefTileWin * P.W)) ? 1u : 0u;
            ss->stats.tma_used = use_tma;
            CK(launch_prefilter(P, pc, pk, po, c.total_windows, pgrid, st, tmap, use_tma));
            CK(cudaEventRecord(ss->ev[4], st));
            sx_list_offsets_kernel<<<1, 1024, 0, st>>>(ss->d_ccount, ss->d_coff, (uint32_t)pgrid, ss->d_counters);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ss->ev[5], st));
            ss->stats.kernel_launches += 2;
            X.list = ss->d_list + (size_t)tile_lo * kPrefTileWin;
            X.ne_ptr = ss->d_counters + 2;
            X.ne_static = 0;
            X.cta_off = ss->d_coff;
            X.ncta = (uint32_t)pgrid;
            X.region_stride = (unsigned long long)tiles_per_cta * kPrefTileWin;
            ss->last_ncta = (uint32_t)pgrid; ss->last_region_stride = (size_t)X.region_stride;
            ss->last_list_compact = false; ss->last_list_base = X.list;
        } else {
            CK(cudaEventRecord(ss->ev[4], st));
            CK(cudaEventRecord(ss->ev[5], st));
            X.list = nullptr;
            X.ne_ptr = nullptr;
            X.ne_static = nwin;
            X.cta_off = nullptr;
            X.ncta = 0;
            X.region_stride = 0;
        }
        const unsigned xgrid = (unsigned)std::max<long long>(1, std::min<long long>(max_blocks, (long long)ss->num_sms * 4));
        CK(launch_exact_enc(P, O, X, xgrid, st));
        ss->stats.kernel_launches++;
        CK(cudaEventRecord(ss->ev[1], st));
        ss->stats.host_phase_ms[0] = ms_since(c.t_begin);
        if (c.tail_n && c.buf_is_device) CK(cudaMemcpyAsync(c.tail + 8 - c.tail_n, c.d_in + c.len - c.tail_n, c.tail_n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(counters, ss->d_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&fin, ss->d_final, sizeof fin, cudaMemcpyDeviceToHost, st));
        if (c.w_lo > 0) CK(cudaMemcpyAsync(&hc->piece[0].overflow, &ss->d_ctl->rc.fail, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ss->stats.d2h_bytes += sizeof counters + sizeof fin;
        if (c.w_lo > 0 && hc->piece[0].overflow) { set_err(SX_ERR_UNSUPPORTED, "range start: no carry-independent window found (halo too short?)"); return fail; }
        if (counters[7]) {
            set_err(SX_ERR_UNSUPPORTED, "Big5 / EUC-JP: more than 256 KiB of lead / trail bytes without a byte that resynchronises the decoder");
            return fail;
        }
        if (!fin.overflow && counters[0] <= ss->rec_cap && counters[1] <= ss->text_cap) break;
        if (attempt >= 2) { set_err(SX_ERR_CUDA, "output buffers still too small after regrowing"); return fail; }
        need_recs = (size_t)counters[0] + 1024;
        need_text = (size_t)counters[1] + 4096;
        ss->stats.relaunches++;
    }
    if (!pc.enabled) counters[2] = (unsigned long long)nwin;
    ss->stats.host_phase_ms[1] = ms_since(c.t_begin);
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, ss->ev[0], ss->ev[4]);
        ss->stats.prefilter_kernel_ms = pc.enabled ? ms : 0.f;
        cudaEventElapsedTime(&ms, ss->ev[4], ss->ev[5]);
        ss->stats.list_kernels_ms = pc.enabled ? ms : 0.f;
        cudaEventElapsedTime(&ms, ss->ev[5], ss->ev[1]);
        ss->stats.exact_kernel_ms = ms;
        cudaEventElapsedTime(&ms, ss->ev[0], ss->ev[1]);
        ss->stats.scan_kernel_ms = ms;
        ss->stats.prefilter_used = pc.enabled;
        ss->stats.pieces = 1;
    }
    c.nrec = (size_t)counters[0];
    c.ntext = (size_t)counters[1];
    c.windows_listed = counters[2];
    ss->last_list_len = (size_t)counters[2];
    c.fin = fin;
    c.direct_out = false;
    c.records_in_order = false;
    const size_t nrec = c.nrec, ntext = c.ntext;
    if (nrec > 0 && ss->use_direct) {
        // order the records on the device and write them as findings into a pinned set; the text follows (transcoded on
        // the device, one copy to the address the findings already point to)
        const size_t nblocks = (size_t)((counters[2] + kThreads - 1) / kThreads);
        pinned_release(fc->set);
        if (pinned_acquire(nrec * sizeof(WireFinding), ntext, &fc->set) && grow(&ss->d_bpos, &ss->bpos_cap, nblocks + 1)) {
            ScanOut O{ss->d_recs, ss->rec_cap, ss->text_cap, ss->d_blocks, ss->d_counters, ss->d_final, reinterpret_cast<uint4*>(fc->set.f),
                      fc->set.fcap / sizeof(WireFinding), nullptr, c.input_file_id, ss->m.mission_id};
            sx_order_scan_kernel<<<1, 1024, 0, st>>>(ss->d_blocks, ss->d_bpos, nblocks);
            const unsigned wgrid = (unsigned)std::min<size_t>((nblocks + 7) / 8, (size_t)ss->num_sms * 8);
            sx_order_write_kernel<<<wgrid, 256, 0, st>>>(P, O, ss->d_blocks, ss->d_bpos, nblocks, nrec);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&c.fin, ss->d_final, sizeof c.fin, cudaMemcpyDeviceToHost, st));  // first / last record flags
            const int mgrid = (int)std::min<size_t>((nrec + 255) / 256, (size_t)ss->num_sms * 8);
            CK(cudaEventRecord(ss->ev[2], st));
            sx_materialize_kernel<<<mgrid, 256, 0, st>>>(P, ss->d_recs, nrec, ss->d_text, ss->text_cap);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ss->ev[3], st));
            ss->stats.kernel_launches += 3;
            if (ntext) CK(cudaMemcpyAsync(fc->set.t, ss->d_text, ntext, cudaMemcpyDeviceToHost, st));
            ss->stats.d2h_bytes += nrec * sizeof(WireFinding) + ntext;
            CK(cudaStreamSynchronize(st));
            float ms = 0;
            cudaEventElapsedTime(&ms, ss->ev[2], ss->ev[3]);
            ss->stats.materialize_kernel_ms = ms;
            c.direct_out = true;
            c.wire = 16;
        } else {
            cudaGetLastError();
            pinned_release(fc->set);
        }
    }
    return 1;
}

static sx_finding_collection* scan_impl(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                        int buf_is_device, int is_last, size_t lo, size_t hi, int prefix_unknown, void* cuda_stream) {
    sx_finding_collection* const fail = nullptr;
    if (!ss) { set_err(SX_ERR_ARGUMENT, "state is NULL"); return fail; }
    if (len > 0 && !buf) { set_err(SX_ERR_ARGUMENT, "buf is NULL"); return fail; }
    const uint32_t q = ss->m.output_line_char_nb_max;
    const uint32_t W = 2 * q;
    if (slice_len == 0 || slice_len > 0x7FFFFFFFull) { set_err(SX_ERR_ARGUMENT, "slice_len must be in 1..2^31-1"); return fail; }
    if (lo > hi || hi > len || (lo % slice_len) != 0 || (hi != len && (hi % slice_len) != 0)) {
        set_err(SX_ERR_ARGUMENT, "range must satisfy lo <= hi <= len with lo and hi multiples of slice_len (or hi == len)");
        return fail;
    }
    if (prefix_unknown && lo == 0 && len > 0) { set_err(SX_ERR_ARGUMENT, "SX_RANGE_PREFIX_UNKNOWN needs a halo in front of the range (lo > 0)"); return fail; }
    const uint32_t wps = (uint32_t)((slice_len + W - 1) / W);
    memset(&ss->stats, 0, sizeof ss->stats);
    CallCtx c;
    c.t_begin = std::chrono::steady_clock::now();
    sx_finding_collection* fc = new sx_finding_collection();
    fc->first_byte_position = ss->consumed + lo;
    if (len == 0 || lo == hi) return fc;  // finding_collection.rs:124: the window loop does not run, state untouched
    struct Guard { sx_finding_collection* p; ~Guard() { delete p; } } guard{fc};

    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    struct DevRestore { int d; ~DevRestore() { if (d >= 0) cudaSetDevice(d); } } dev_restore{prev_dev};  // the caller's current device is left as it was
    CK(cudaSetDevice(ss->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;

    // ---- input -----------------------------------------------------------------------------------
    const uint8_t* d_in = nullptr;
    if (buf_is_device) d_in = (const uint8_t*)buf;
    else {
        if (!grow(&ss->d_in, &ss->d_in_cap, len + 64)) return fail;
        CK(cudaMemcpyAsync(ss->d_in, buf, len, cudaMemcpyHostToDevice, st));
        ss->stats.h2d_bytes += len;
        d_in = ss->d_in;
    }

    // ---- parameters ------------------------------------------------------------------------------
    ScanParams& P = c.P;
    memset(&P, 0, sizeof P);
    P.in = d_in;
    P.len = (int64_t)len;
    P.slice_len = (uint32_t)slice_len;
    P.W = W; P.q = q; P.n = ss->m.chars_min_nb;
    P.enc = ss->m.encoding_id;
    const uint32_t unit = (P.enc == ENC_UTF16LE || P.enc == ENC_UTF16BE) ? 2 : (P.enc == ENC_UTF32LE || P.enc == ENC_UTF32BE) ? 4 : 1;
    P.align = unit > 1 ? (uint32_t)((unit - (ss->npend % unit)) % unit) : 0;
    P.af_lo = ss->m.af_lo; P.af_hi = ss->m.af_hi; P.ubf = ss->m.ubf;
    P.base_consumed = ss->consumed;
    P.npend = ss->npend;
    P.is_last = is_last ? 1 : 0;
    memcpy(P.pend, ss->pend, 8);
    for (size_t i = 0; i < 8 && i < ss->leftover.size(); ++i) P.carry_text8[i] = ss->leftover[i];
    P.carry_text_len = (uint32_t)ss->leftover.size();
    if (ss->cut) P.k0 = carry_cut();
    else if (!ss->leftover.empty()) {
        // the leftover is re-scanned in front of the next buffer (finding_collection.rs:214-221): what that scan would
        // remember of it -- does it hold the grep char (helper.rs:252-254), its last multi-byte lead byte (helper.rs:221)
        uint8_t fl = CF_HOSTCARRY;
        uint32_t last_lead = 0;
        for (uint8_t b : ss->leftover) {
            if (ss->m.grep_char >= 0 && b == (uint8_t)ss->m.grep_char) fl |= CF_GREP;
            if (b >= 0xC0) last_lead = b;
        }
        P.k0 = Carry{K_L, fl, (uint16_t)utf8_char_count(ss->leftover), (uint32_t)ss->npend, 0,
                     ss->m.require_same_unicode_block ? last_lead : 0u};
    } else P.k0 = carry_none();
    // --grep-char / --same-unicode-block / chars_min_nb > q: the general automaton (sx_core.cuh WinAuto) on every window
    P.grep_char = ss->m.grep_char >= 0 ? (int32_t)ss->m.grep_char : -1;
    P.same_block = ss->m.require_same_unicode_block ? 1u : 0u;
    P.general = (P.grep_char >= 0 || P.same_block || P.n > P.q) ? 1u : 0u;
    memcpy(P.sb_table, ss->m.sb_table, sizeof P.sb_table);
    const uint32_t *host_mb_a = nullptr, *host_mb_b = nullptr;
    if (P.enc == ENC_BIG5 || P.enc == ENC_EUCJP) {
        MbDeviceTables t;
        if (!mb_tables_for(ss->device, &t)) { set_err(SX_ERR_CUDA, "cannot upload the Big5 / EUC-JP index tables"); return fail; }
        P.mb_a = P.enc == ENC_BIG5 ? t.big5 : t.jis0208;
        P.mb_b = P.enc == ENC_BIG5 ? nullptr : t.jis0212;
        host_mb_a = P.enc == ENC_BIG5 ? kSxBig5Index : kSxJis0208Index;
        host_mb_b = P.enc == ENC_BIG5 ? nullptr : kSxJis0212Index;
        P.mb_fail = reinterpret_cast<uint32_t*>(ss->d_counters + 7);
    }

    long long total_windows;
    {
        const long long full = (long long)(len / slice_len);
        const size_t rest = len - (size_t)full * slice_len;
        total_windows = full * wps + (long long)((rest + W - 1) / W);
    }
    if (total_windows > 0x7FFFFFF0LL) { set_err(SX_ERR_UNSUPPORTED, "stream too long for one call"); return fail; }
    const bool in_aligned16 = (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0;
    c.pc = make_pref_cfg(P, in_aligned16, host_mb_a, host_mb_b);
    // General missions: an unlisted window's carry-out must not depend on its carry-in.  --grep-char and
    // --same-unicode-block get there with one more listing rule each (PrefCfg::kill_trail / sb_rule, sx_core.cuh);
    // n > q (whole segments dropped) runs without the prefilter.  DESIGN.md 7.
    if (!ss->use_prefilter || (P.general && !pref_general_ok(P))) c.pc.enabled = 0;
    c.ss = ss; c.fc = fc; c.input_file_id = input_file_id; c.d_in = d_in; c.len = len; c.st = st;
    c.total_windows = total_windows;
    c.w_lo = (long long)(lo / slice_len) * wps;
    c.w_hi = hi == len ? total_windows : (long long)(hi / slice_len) * wps;
    c.in_aligned16 = in_aligned16;
    c.prefix_known = !prefix_unknown;
    c.buf_is_device = buf_is_device != 0;
    const bool tail = hi == len;  // the range reaches the end of the buffer: the ScannerState moves on
    // bytes that stay inside the decoder: the last npend bytes of (old pend ++ buffer)
    c.tail_n = tail ? std::min<size_t>(8, len) : 0;
    if (!buf_is_device && c.tail_n) memcpy(c.tail + 8 - c.tail_n, (const uint8_t*)buf + len - c.tail_n, c.tail_n);

    const bool sparse = c.pc.enabled && ss->use_sparse && has_sparse_enc(P.enc) && !P.general;
    if (!(sparse ? run_sparse(c) : run_block(c))) return fail;

    const size_t nrec = c.nrec, ntext = c.ntext;
    const FinalState& fin = c.fin;
    ss->stats.windows_total = (uint64_t)(c.w_hi - c.w_lo);
    ss->stats.windows_listed = c.windows_listed;
    ss->have_history = true;
    ss->last_nrec = nrec; ss->last_ntext = ntext; ss->last_len = hi - lo;
    ss->rec_per_byte = std::max(1.0 / 4096, 1.3 * (double)nrec / (double)(hi - lo));
    ss->text_per_byte = std::max(1.0 / 256, 1.3 * (double)ntext / (double)(hi - lo));
    ss->stats.n_records = nrec;
    ss->stats.text_bytes = ntext;

    // ---- the collection ----------------------------------------------------------------------------
    std::vector<uint8_t> new_leftover;
    bool have_leftover = false;
    if (c.direct_out) {
        // The findings are already in the collection's pinned set in their final form, their text at the address they
        // point to; the host neither copies nor converts records.
        const auto t_post = std::chrono::steady_clock::now();
        ss->stats.host_phase_ms[2] = std::chrono::duration<float, std::milli>(t_post - c.t_begin).count();
        const bool tail_is_leftover = nrec > 0 && (fin.last_flags & RF_LEFTOVER) != 0;
        const size_t n_out = nrec - (tail_is_leftover ? 1 : 0);
        fc->wire = c.wire;
        fc->n = n_out;
        fc->base = P.base_consumed;
        fc->file_id = (int16_t)input_file_id;
        fc->mission_id = ss->m.mission_id;
        auto wire_text = [&](size_t i, const uint8_t** p, size_t* len) { fc->wire_text(i, p, len); };
        // host text in front of device text: the first record of a run that began in the previous call, the leftover
        auto with_host_text = [&](size_t i, std::vector<uint8_t>& dst) {
            const uint8_t* p; size_t len;
            wire_text(i, &p, &len);
            dst.assign(ss->leftover.begin(), ss->leftover.end());
            dst.insert(dst.end(), p, p + len);
        };
        if (tail_is_leftover) {
            if (fin.last_flags & RF_HOSTCARRY) with_host_text(nrec - 1, new_leftover);
            else {
                const uint8_t* p; size_t len;
                wire_text(nrec - 1, &p, &len);
                new_leftover.assign(p, p + len);
            }
            have_leftover = true;
        }
        if (n_out > 0 && (fin.first_flags & RF_HOSTCARRY)) {
            std::vector<uint8_t> t;
            with_host_text(0, t);
            fc->text.resize(t.size() + 1);
            memcpy(fc->text.data(), t.data(), t.size());
            fc->have_first = true;
        }
        ss->stats.host_post_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_post).count();
    } else {
    // legacy download (test hook sx_scanner_state_set_direct_output(0), or no pinned memory): records and text are copied
    // from the device and converted on the host
    const bool in_order = c.records_in_order;  // one descriptor
    const size_t nblocks = in_order ? 1 : (size_t)((c.windows_listed + kThreads - 1) / kThreads);
    if (!grow_pinned(&ss->h_recs, &ss->h_recs_cap, nrec + 1)) return fail;
    if (!grow_pinned(&ss->h_text, &ss->h_text_cap, ntext + 1)) return fail;
    if (!grow_pinned(&ss->h_blocks, &ss->h_blocks_cap, nblocks + 1)) return fail;
    const Record* recs = ss->h_recs;
    const uint8_t* text = ss->h_text;
    const uint2* tiles = ss->h_blocks;
    if (nrec) {
        const int mgrid = (int)std::min<size_t>((nrec + 255) / 256, (size_t)ss->num_sms * 8);
        CK(cudaEventRecord(ss->ev[2], st));
        sx_materialize_kernel<<<mgrid, 256, 0, st>>>(P, ss->d_recs, nrec, ss->d_text, ss->text_cap);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ss->ev[3], st));
        ss->stats.kernel_launches++;
        CK(cudaMemcpyAsync(ss->h_recs, ss->d_recs, nrec * sizeof(Record), cudaMemcpyDeviceToHost, st));
        if (ntext) CK(cudaMemcpyAsync(ss->h_text, ss->d_text, ntext, cudaMemcpyDeviceToHost, st));
        if (in_order) ss->h_blocks[0] = make_uint2(0u, (unsigned)nrec);
        else CK(cudaMemcpyAsync(ss->h_blocks, ss->d_blocks, nblocks * sizeof(uint2), cudaMemcpyDeviceToHost, st));
        ss->stats.d2h_bytes += nrec * sizeof(Record) + ntext + (in_order ? 0 : nblocks * sizeof(uint2));
    }
    CK(cudaStreamSynchronize(st));
    if (nrec) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ss->ev[2], ss->ev[3]);
        ss->stats.materialize_kernel_ms = ms;
    }

    const auto t_post = std::chrono::steady_clock::now();
    ss->stats.host_phase_ms[2] = std::chrono::duration<float, std::milli>(t_post - c.t_begin).count();
    // ---- build the collection in stream order (blocks own contiguous record ranges) -------------------
    // At most two records carry host text in front of their device text: the very first record (a run that
    // began in the previous call) and the final leftover pseudo record, which is always the last one.
    const size_t extra_text = 2 * (ss->leftover.size() + 8 * (size_t)q + 64);
    fc->text.resize(ntext + extra_text + 1);
    if (nrec) {
        std::vector<size_t> out_off(nblocks + 1);
        size_t acc = 0, last_block = 0;
        for (size_t t = 0; t < nblocks; ++t) { out_off[t] = acc; acc += tiles[t].y; if (tiles[t].y) last_block = t; }
        out_off[nblocks] = acc;
        const uint2 lb = tiles[last_block];
        const bool tail_is_leftover = (recs[(size_t)lb.x + lb.y - 1].flags & RF_LEFTOVER) != 0;
        const size_t n_out = acc - (tail_is_leftover ? 1 : 0);
        fc->v.resize(n_out);
        uint8_t* const tbase = fc->text.data();
        sx_finding* const out = fc->v.data();
        const int16_t fid = (int16_t)input_file_id;
        const uint8_t mid = ss->m.mission_id;
        // [t0, t1): descriptors; with a single descriptor (records already in order) [k0, k1) is the record range
        auto fill = [&](size_t t0, size_t t1, size_t k0 = 0, size_t k1 = ~(size_t)0) {
            for (size_t t = t0; t < t1; ++t) {
                const size_t kb = std::min<size_t>(k0, tiles[t].y), ke = std::min<size_t>(k1, tiles[t].y);
                const Record* r = recs + tiles[t].x + kb;
                size_t o = out_off[t] + kb;
                for (size_t k = kb; k < ke; ++k, ++r, ++o) {
                    if (o >= n_out) break;  // the trailing leftover pseudo record
                    sx_finding f;
                    f.position = r->position;
                    f.precision = (uint8_t)r->precision;
                    f.completes_previous = (r->flags & RF_COMPLETES) ? 1 : 0;
                    f.input_file_id = fid;
                    f.mission_id = mid;
                    f.s = tbase + r->text_off;
                    f.s_len = r->text_len;
                    f.in_start = r->in_start;
                    f.in_len = r->in_len;
                    out[o] = f;
                }
            }
        };
        const unsigned nthreads = acc > 65536 ? std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
        if (nthreads <= 1) {
            if (ntext) memcpy(tbase, text, ntext);
            fill(0, nblocks);
        } else {  // memory bound: split the copy of the text arena and the record conversion over a few threads
            std::vector<std::thread> th;
            for (unsigned k = 0; k < nthreads; ++k)
                th.emplace_back([&, k]() {
                    const size_t a0 = ntext * k / nthreads, a1 = ntext * (k + 1) / nthreads;
                    if (a1 > a0) memcpy(tbase + a0, text + a0, a1 - a0);
                    if (nblocks == 1) fill(0, 1, acc * k / nthreads, acc * (k + 1) / nthreads);
                    else fill(nblocks * k / nthreads, nblocks * (k + 1) / nthreads);
                });
            for (auto& t : th) t.join();
        }
        size_t extra_off = ntext;
        auto with_host_text = [&](const Record& r, const uint8_t** s, size_t* s_len) {
            uint8_t* d = tbase + extra_off;
            memcpy(d, ss->leftover.data(), ss->leftover.size());
            memcpy(d + ss->leftover.size(), tbase + r.text_off, r.text_len);
            *s = d;
            *s_len = ss->leftover.size() + r.text_len;
            extra_off += *s_len;
        };
        // first record in stream order
        size_t fb = 0;
        while (fb < nblocks && tiles[fb].y == 0) ++fb;
        if (n_out > 0 && fb < nblocks && (recs[tiles[fb].x].flags & RF_HOSTCARRY)) {
            const uint8_t* s; size_t sl;
            with_host_text(recs[tiles[fb].x], &s, &sl);
            out[0].s = s;
            out[0].s_len = (uint32_t)sl;
        }
        if (tail_is_leftover) {
            const Record& r = recs[(size_t)lb.x + lb.y - 1];
            const uint8_t* s = tbase + r.text_off; size_t sl = r.text_len;
            if (r.flags & RF_HOSTCARRY) with_host_text(r, &s, &sl);
            new_leftover.assign(s, s + sl);
            have_leftover = true;
        }
    }

    }

    // ---- ScannerState update (finding_collection.rs:330-338): only when the range reaches the end of the buffer ----
    if (tail) {
        ss->cut = fin.carry.kind == K_C;
        if (fin.carry.kind == K_L && fin.carry.k > 0 && have_leftover) ss->leftover.swap(new_leftover);
        else ss->leftover.clear();
        uint8_t all[16];
        memcpy(all, ss->pend, 8);  // old pend occupies all[8-npend..8)
        // concatenation old_pend ++ tail, right aligned: when len < 8 the old pend bytes must follow on directly
        uint8_t cat[16];
        size_t cn = 0;
        for (int i = 8 - ss->npend; i < 8; ++i) cat[cn++] = all[i];
        for (size_t i = 8 - c.tail_n; i < 8; ++i) cat[cn++] = c.tail[i];
        const int np = fin.npend;
        memset(ss->pend, 0, 8);
        if (np > 0 && (size_t)np <= cn) memcpy(ss->pend + 8 - np, cat + cn - np, (size_t)np);
        ss->npend = (np > 0 && (size_t)np <= cn) ? np : 0;
        ss->consumed += len;
    }
    {
        ss->stats.host_total_ms = ms_since(c.t_begin);
        if (!c.direct_out) ss->stats.host_post_ms = ss->stats.host_total_ms - ss->stats.host_phase_ms[2];
        ss->stats.host_phase_ms[3] = ss->stats.host_total_ms;
    }
    guard.p = nullptr;
    return fc;
}

extern "C" sx_finding_collection* sx_scan_stream(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len,
                                                 size_t slice_len, int buf_is_device, int is_last, void* cuda_stream) {
    return scan_impl(ss, input_file_id, buf, len, slice_len, buf_is_device, is_last, 0, len, 0, cuda_stream);
}

extern "C" sx_finding_collection* sx_scan_range(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                                int buf_is_device, int is_last, size_t lo, size_t hi, int flags, void* cuda_stream) {
    return scan_impl(ss, input_file_id, buf, len, slice_len, buf_is_device, is_last, lo, hi, (flags & SX_RANGE_PREFIX_UNKNOWN) ? 1 : 0,
                     cuda_stream);
}

extern "C" sx_finding_collection* sx_finding_collection_from(sx_scanner_state* ss, int input_file_id, const uint8_t* buf,
                                                             size_t len, int is_last) {
    return sx_scan_stream(ss, input_file_id, buf, len, len ? len : 1, 0, is_last, nullptr);
}

extern "C" size_t sx_merge(const sx_finding_collection* const* fcs, size_t n, const sx_finding** out) {
    // finding.rs:92-109 via itertools::kmerge (main.rs:133): position, then mission_id; every
    // collection is position-monotone, so a stable sort of the concatenation is the k-way merge.
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        const sx_finding* f = sx_fc_data(fcs[i]);
        for (size_t j = 0, m = sx_fc_len(fcs[i]); j < m; ++j) out[k++] = f + j;
    }
    std::stable_sort(out, out + k, [](const sx_finding* a, const sx_finding* b) {
        if (a->position != b->position) return a->position < b->position;
        return a->mission_id < b->mission_id;
    });
    return k;
}

// ---------------------------------------------------------------------------------------------
// Asynchronous calls and the streaming driver
// ---------------------------------------------------------------------------------------------
static sx_pending* submit_scan(sx_scanner_state* ss, std::function<sx_finding_collection*()> fn) {
    if (!ss) { set_err(SX_ERR_ARGUMENT, "state is NULL"); return nullptr; }
    if (!ss->worker) {
        ss->worker = new ScanWorker();
        ss->worker->th = std::thread([w = ss->worker] { w->run(); });
    }
    sx_pending* p = new sx_pending();
    {
        std::lock_guard<std::mutex> lk(ss->worker->mu);
        ss->worker->q.emplace_back([p, fn = std::move(fn)] {
            sx_finding_collection* fc = fn();
            std::lock_guard<std::mutex> lk2(p->mu);
            p->fc = fc;
            if (!fc) { p->err = g_err_code; p->msg = g_err_msg; }  // the worker thread's error travels with the handle
            p->done = true;
            p->cv.notify_all();
        });
    }
    ss->worker->cv.notify_one();
    return p;
}

extern "C" sx_pending* sx_scan_stream_async(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                            int buf_is_device, int is_last, void* cuda_stream) {
    return submit_scan(ss, [=] { return scan_impl(ss, input_file_id, buf, len, slice_len, buf_is_device, is_last, 0, len, 0, cuda_stream); });
}
extern "C" sx_pending* sx_scan_range_async(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                           int buf_is_device, int is_last, size_t lo, size_t hi, int flags, void* cuda_stream) {
    return submit_scan(ss, [=] {
        return scan_impl(ss, input_file_id, buf, len, slice_len, buf_is_device, is_last, lo, hi, (flags & SX_RANGE_PREFIX_UNKNOWN) ? 1 : 0, cuda_stream);
    });
}
extern "C" int sx_pending_ready(const sx_pending* p) {
    if (!p) return 1;
    std::lock_guard<std::mutex> lk(const_cast<sx_pending*>(p)->mu);
    return p->done ? 1 : 0;
}
extern "C" sx_finding_collection* sx_fc_wait(sx_pending* p) {
    if (!p) { set_err(SX_ERR_ARGUMENT, "handle is NULL"); return nullptr; }
    sx_finding_collection* fc;
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->cv.wait(lk, [&] { return p->done; });
        fc = p->fc;
        if (!fc) set_err(p->err, p->msg);
    }
    delete p;
    return fc;
}

// Streaming driver: main.rs:143-167 over input.rs:104-168 for the GPU.  One input (file, pipe, anything behind `read`) is
// read in pieces of chunk_bytes (a multiple of 4096, so the slice grid inside the library is the reference's: every
// input starts a new grid, input.rs:104-168), every piece is uploaded once per device and scanned by every state -- the
// states' scans run side by side on their worker threads, the read and the upload of piece k + 1 overlap the scan of
// piece k (two pinned staging buffers, two device buffers per device).  `batch` receives the piece's collections in
// state order (they are the caller's: sx_fc_free); merging them with sx_merge gives the order of main.rs:118-136.
// Carry, pending decoder bytes and byte counters stay in the states, so the caller simply calls this once per input
// with the reference's 1-based file label (or -1).  The "last buffer" flag is never set (input.rs:130-137).
extern "C" int sx_scan_reader(sx_scanner_state* const* states, size_t n_states, int input_file_id, sx_read_fn read, void* read_user,
                              size_t chunk_bytes, sx_batch_fn batch, void* batch_user) {
    const int fail = -1;
    if (!states || n_states == 0 || !read || !batch) { set_err(SX_ERR_ARGUMENT, "states, read and batch must be given"); return fail; }
    if (chunk_bytes == 0 || (chunk_bytes % 4096) != 0) { set_err(SX_ERR_ARGUMENT, "chunk_bytes must be a positive multiple of 4096"); return fail; }
    std::vector<int> devices;
    for (size_t i = 0; i < n_states; ++i) {
        if (!states[i]) { set_err(SX_ERR_ARGUMENT, "state is NULL"); return fail; }
        if (std::find(devices.begin(), devices.end(), states[i]->device) == devices.end()) devices.push_back(states[i]->device);
    }
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    struct Dev { int id; uint8_t* buf[2] = {nullptr, nullptr}; cudaStream_t cs = nullptr; cudaEvent_t up[2] = {nullptr, nullptr}; };
    std::vector<Dev> devs(devices.size());
    uint8_t* host[2] = {nullptr, nullptr};
    int rc = 0;
    auto cleanup = [&] {
        for (auto& d : devs) {
            cudaSetDevice(d.id);
            for (int k = 0; k < 2; ++k) { if (d.buf[k]) cudaFree(d.buf[k]); if (d.up[k]) cudaEventDestroy(d.up[k]); }
            if (d.cs) cudaStreamDestroy(d.cs);
        }
        for (int k = 0; k < 2; ++k) if (host[k]) cudaFreeHost(host[k]);
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
    };
    bool ok = true;
    for (int k = 0; ok && k < 2; ++k) ok = cuda_ok(cudaHostAlloc((void**)&host[k], chunk_bytes, cudaHostAllocPortable), "cudaHostAlloc");
    for (size_t i = 0; ok && i < devs.size(); ++i) {
        devs[i].id = devices[i];
        ok = cuda_ok(cudaSetDevice(devs[i].id), "cudaSetDevice") && cuda_ok(cudaStreamCreateWithFlags(&devs[i].cs, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int k = 0; ok && k < 2; ++k)
            ok = cuda_ok(cudaMalloc(&devs[i].buf[k], chunk_bytes), "cudaMalloc") && cuda_ok(cudaEventCreateWithFlags(&devs[i].up[k], cudaEventDisableTiming), "cudaEventCreate");
    }
    if (!ok) { cleanup(); return fail; }
    auto fill = [&](int k) {  // short reads are retried: only the end of the input ends a piece early
        size_t n = 0;
        while (n < chunk_bytes) {
            const size_t got = read(read_user, host[k] + n, chunk_bytes - n);
            if (got == 0) break;
            n += got;
        }
        return n;
    };
    auto upload = [&](int k, size_t n) {
        for (auto& d : devs) {
            if (!cuda_ok(cudaSetDevice(d.id), "cudaSetDevice") ||
                !cuda_ok(cudaMemcpyAsync(d.buf[k], host[k], n, cudaMemcpyHostToDevice, d.cs), "cudaMemcpyAsync") ||
                !cuda_ok(cudaEventRecord(d.up[k], d.cs), "cudaEventRecord")) return false;
        }
        return true;
    };
    int k = 0;
    size_t n = fill(k);
    if (n && !upload(k, n)) { cleanup(); return fail; }
    std::vector<sx_pending*> pend(n_states);
    std::vector<sx_finding_collection*> fcs(n_states);
    while (n && rc == 0) {
        // piece k is on its way to every device: scans wait for the upload on their own streams
        for (size_t i = 0; i < n_states; ++i) {
            sx_scanner_state* ss = states[i];
            const Dev& d = devs[std::find(devices.begin(), devices.end(), ss->device) - devices.begin()];
            cudaSetDevice(d.id);
            cudaStreamWaitEvent(ss->own, d.up[k], 0);
            pend[i] = sx_scan_stream_async(ss, input_file_id, d.buf[k], n, 4096, 1, 0, ss->own);
        }
        // meanwhile: read and upload the next piece (its buffers were released when the scans of piece k - 1 returned)
        const int k2 = k ^ 1;
        size_t n2 = 0;
        bool up_ok = true;
        if (n == chunk_bytes) {
            for (auto& d : devs) { cudaSetDevice(d.id); cudaEventSynchronize(d.up[k]); }  // host[k2] was read by the copy of piece k - 1: long done; host[k] must be free before the NEXT fill
            n2 = fill(k2);
            if (n2) up_ok = upload(k2, n2);
        }
        bool scan_ok = true;
        for (size_t i = 0; i < n_states; ++i) {
            fcs[i] = pend[i] ? sx_fc_wait(pend[i]) : nullptr;
            if (!fcs[i]) scan_ok = false;
        }
        if (!scan_ok || !up_ok) {
            for (auto* fc : fcs) if (fc) sx_fc_free(fc);
            cleanup();
            return fail;
        }
        rc = batch(batch_user, input_file_id, fcs.data(), n_states);
        k = k2;
        n = n2;
    }
    for (auto& d : devs) { cudaSetDevice(d.id); cudaStreamSynchronize(d.cs); }
    cleanup();
    return rc;
}

static size_t file_read_fn(void* user, uint8_t* dst, size_t cap) { return fread(dst, 1, cap, (FILE*)user); }
extern "C" int sx_scan_file(sx_scanner_state* const* states, size_t n_states, int input_file_id, const char* path, size_t chunk_bytes,
                            sx_batch_fn batch, void* batch_user) {
    FILE* f = path ? fopen(path, "rb") : stdin;
    if (!f) {  // input.rs:78-84: reported, then treated as an empty input
        fprintf(stderr, "Error: can not read file `%s`\n", path);
        return 0;
    }
    const int rc = sx_scan_reader(states, n_states, input_file_id, file_read_fn, f, chunk_bytes, batch, batch_user);
    if (path) fclose(f);
    return rc;
}

extern "C" int sx_fill_random(void* device_buf, size_t len, uint64_t seed, uint64_t stream_offset, int device, void* cuda_stream) {
    const int fail = -1;
    if (sx_device_count() <= 0) { set_err(SX_ERR_NO_DEVICE, "no CUDA device"); return fail; }
    CK(cudaSetDevice(device));
    if (len == 0) return 0;
    const unsigned long long nwords = (len + 7) / 8 + 1;
    const int grid = (int)std::min<unsigned long long>((nwords + 255) / 256, 148ull * 16);
    sx_fill_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>((uint8_t*)device_buf, len, seed, stream_offset);
    CK(cudaGetLastError());
    return 0;
}
