// sx_exact.cuh -- the exact kernel (sx_exact_kernel<Dec>) and the structures it shares with the host
// code in sx_scan.cu.  One instantiation per decoder lives in its own translation unit
// (sx_exact_inst.cu compiled with -DSX_INST=n) so that the library builds in parallel.
#pragma once
#include "../../include/stringsext_b200.h"
#include "sx_mask_utf8.cuh"
#include <cstddef>
#include <cuda_runtime.h>
#include <type_traits>

namespace sx {

constexpr int kMaxPrefCtas = 1024;  // prefilter CTAs (list regions)
constexpr int kThreads = 128;      // exact kernel: one list entry per thread
constexpr int kPrefThreads = 256;  // prefilter kernel: one window per thread, 256 windows per tile

struct FinalState {
    Carry carry;
    int32_t npend;
    uint32_t overflow;
    uint32_t first_flags, last_flags;  // RF_* of the first / last record in stream order (direct host output only)
    uint32_t text_fallback;            // sparse pipeline: some CTA left its text to the device arena (materialize + download)
};

struct ScanOut {
    Record* recs;
    unsigned long long rec_cap;
    unsigned long long text_cap;
    uint2* block_desc;             // per exact-kernel block {first record, record count}
    unsigned long long* counters;  // [0] records, [1] text bytes, [2] list entries, [3] next block of the exact kernel
    FinalState* final_state;
    // direct output: the GPU writes every finding as a 16-byte wire record (WireFinding) -- into a device staging array
    // that the copy engine moves to the collection's pinned memory (sparse pipeline), or straight into that memory
    // (block path); the host expands them to sx_finding on demand, page by page (sx_fc_get / sx_fc_data)
    uint4* host_findings;          // nullptr: off (records are downloaded and converted on the host)
    unsigned long long host_cap;
    // per-stage pipeline: 8-byte wire records (Wire8, written through `host_findings` as unsigned long long) whose text
    // offsets are implicit -- the text arena is gap free in record order -- except for one explicit offset per page of
    // kWirePage records, written here
    unsigned long long* page_base;
    int32_t file_id;
    uint32_t mission_id;
};
constexpr uint32_t kWirePage = 4096;

// One finding on the wire (finding.rs:51-74 minus what the host knows: mission, file, text base address):
//   x, y: position relative to the call's first byte (40 bits) | text length (22 bits) << 40 | precision (2 bits) << 62
//   z, w: text offset in the collection's text arena (40 bits) | completes-previous flag << 40
struct WireFinding { unsigned long long a, b; };
static_assert(sizeof(WireFinding) == 16, "wire record");
__host__ __device__ __forceinline__ WireFinding wire_pack(unsigned long long pos_rel, uint32_t text_len, uint32_t precision,
                                                          unsigned long long text_off, bool completes) {
    WireFinding w;
    w.a = (pos_rel & 0xFFFFFFFFFFull) | ((unsigned long long)(text_len & 0x3FFFFFu) << 40) | ((unsigned long long)(precision & 3u) << 62);
    w.b = (text_off & 0xFFFFFFFFFFull) | ((unsigned long long)(completes ? 1u : 0u) << 40);
    return w;
}
// The per-stage pipeline's wire record, 8 bytes: position relative to the call's first byte (40 bits) | text length
// (21 bits) << 40 | precision (2 bits) << 61 | completes-previous flag << 63.  The text of record i starts where the text
// of record i - 1 ends; page_base[i / kWirePage] holds the offset of every kWirePage-th record.
__host__ __device__ __forceinline__ unsigned long long wire8_pack(unsigned long long pos_rel, uint32_t text_len, uint32_t precision,
                                                                  bool completes) {
    return (pos_rel & 0xFFFFFFFFFFull) | ((unsigned long long)(text_len & 0x1FFFFFu) << 40) | ((unsigned long long)(precision & 3u) << 61) |
           ((unsigned long long)(completes ? 1u : 0u) << 63);
}
__device__ __forceinline__ unsigned long long wire8_of(const ScanParams& P, const Record& r) {
    return wire8_pack(r.position - P.base_consumed, r.text_len, r.precision, (r.flags & RF_COMPLETES) != 0);
}
__device__ __forceinline__ void write_host_finding(const ScanParams& P, uint4* dst, const Record& r) {
    const WireFinding w = wire_pack(r.position - P.base_consumed, r.text_len, r.precision, r.text_off, (r.flags & RF_COMPLETES) != 0);
    *dst = make_uint4((uint32_t)w.a, (uint32_t)(w.a >> 32), (uint32_t)w.b, (uint32_t)(w.b >> 32));
}

// Work list of the exact kernel: the windows the prefilter kept, in stream order
// (list == nullptr: every window, entry e is window e).
struct ExactCfg {
    const uint32_t* list;
    const unsigned long long* ne_ptr;  // device: number of list entries (list != nullptr)
    long long ne_static;               // list == nullptr
    long long total_windows;
    uint32_t in_aligned16;
    uint32_t pre_bytes;                // pre-roll length for entries whose predecessor window is not listed
    const uint32_t* cta_off;           // list != nullptr: entry offsets of the prefilter CTAs' list regions (ncta + 1);
    uint32_t ncta;                     // ncta == 0: `list` is compact (entry e is list[e])
    unsigned long long region_stride;  // list region of prefilter CTA b starts at b * region_stride
    // Window range of this pass, [w_first, w_end) of the stream's windows (a whole call: 0 .. total_windows).  The first
    // window of a range is always listed; its carry-in is *k0_ptr (device; written by sx_range_carry_kernel, which walks
    // back to the nearest window whose carry-out does not depend on its carry-in) or P.k0 when k0_ptr == nullptr.  A
    // range that ends before the stream does leaves what its last carry prints in the next window to the next range
    // (whose first window is listed by construction) and produces no final leftover.
    long long w_first, w_end;
    const Carry* k0_ptr;
    unsigned long long ne_cap;         // capacity of the per-entry arrays (pipelined pieces: ne beyond it is an overflow)
};
__device__ __forceinline__ Carry range_carry_in(const ScanParams& P, const ExactCfg& X) { return X.k0_ptr ? *X.k0_ptr : P.k0; }
__device__ __forceinline__ bool range_is_tail(const ExactCfg& X) { return X.w_end >= X.total_windows; }

__device__ __forceinline__ uint32_t swz(uint32_t r) { return r ^ (((r >> 7) & 7u) << 4); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Byte source of the exact kernel: the listed windows are sparse, so they are read straight from
// global memory (L2) with 16-byte vector loads per lane.
struct GlobalTile {
    GlobalSrc g;
    int64_t len;
    bool aligned16;
    const Utf8Tables* tab;
    uint32_t tab_smem;  // 32-bit shared-memory address of tab->tt
    __device__ __forceinline__ const Utf8Tables* tables() const { return tab; }
    __device__ __forceinline__ uint32_t lut(uint32_t i) const {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tab_smem + i));
        return v;
    }
    __device__ __forceinline__ uint32_t cls(uint32_t b) const {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tab_smem + 2048u + b));
        return v;
    }
    __device__ __forceinline__ uint8_t get(int64_t off) const { return g.get(off); }
    __device__ __forceinline__ uint4 load_chunk(int64_t r16, int64_t ws, int64_t we) const {
        if (r16 >= we) return make_uint4(0, 0, 0, 0);
        if (aligned16 && r16 >= 0 && r16 + 16 <= len) return __ldg(reinterpret_cast<const uint4*>(g.in + r16));
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16; ++i) {
            const int64_t o = r16 + i;
            if (o >= ws && o < we) w[i >> 2] |= (uint32_t)g.get(o) << ((i & 3) * 8);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    // Software-pipelined: three 16-byte chunks are in flight while one is being decoded.
    template <class F>
    __device__ __forceinline__ void for_each_byte(int64_t ws, int64_t we, F&& f) const {
        int64_t r16 = ws & ~(int64_t)15;
        uint4 c0 = load_chunk(r16, ws, we);
        uint4 c1 = load_chunk(r16 + 16, ws, we);
        uint4 c2 = load_chunk(r16 + 32, ws, we);
        int64_t pos = ws;
        while (pos < we) {
            const uint4 c3 = load_chunk(r16 + 48, ws, we);
            uint32_t w0 = c0.x, w1 = c0.y, w2 = c0.z, w3 = c0.w;
            const uint32_t i0 = (uint32_t)(pos - r16);
            const uint32_t i1 = (we - r16) < 16 ? (uint32_t)(we - r16) : 16u;
            // deliberately NOT unrolled: one copy of the automaton body (instruction cache)
#pragma unroll 1
            for (uint32_t i = 0; i < i1; ++i) {
                if (i >= i0) f(w0 & 0xFFu, r16 + i);
                w0 = __funnelshift_r(w0, w1, 8);
                w1 = __funnelshift_r(w1, w2, 8);
                w2 = __funnelshift_r(w2, w3, 8);
                w3 >>= 8;
            }
            r16 += 16;
            pos = r16;
            c0 = c1; c1 = c2; c2 = c3;
        }
    }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}

// Exclusive block scan of two 32-bit values (thread order); also returns the block totals.
__device__ __forceinline__ void block_excl_scan2(uint32_t a, uint32_t b, uint32_t* wa, uint32_t* wb, uint32_t& ea,
                                                 uint32_t& eb, uint32_t& ta, uint32_t& tb) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ia = warp_incl_scan(a), ib = warp_incl_scan(b);
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    uint32_t oa = 0, ob = 0;
    ta = 0; tb = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const uint32_t xa = wa[w], xb = wb[w];
        if ((uint32_t)w < warp) { oa += xa; ob += xb; }
        ta += xa; tb += xb;
    }
    ea = oa + ia - a;
    eb = ob + ib - b;
    __syncthreads();
}

struct ExactSmem {
    uint32_t win[kThreads];       // window index of every entry of the block
    uint32_t win_prev, win_next;  // windows of the entries just outside the block (kNoWin: none)
    uint8_t next_adj[kThreads];
    uint8_t adj[kThreads];
    uint8_t declined[kThreads];   // the mask engine declined the entry under its real carry
    WinDesc desc[kThreads];
    Carry kin[kThreads + 1];
    Carry kout[kThreads];
    uint8_t in_known[kThreads + 1];
    uint8_t out_done[kThreads];
    uint32_t warp_a[8], warp_b[8];
    unsigned long long bases[2];
    int32_t last_npend;
    unsigned long long next_block;
    Utf8Tables tables;
    uint32_t cta_off[kMaxPrefCtas + 1];
    Record staged[kThreads][kBufRecs];   // MODE_BUFFER staging of the entry's own records
    Record xstaged[kThreads][kBufRecs];  // ... and of the records its carry-out prints in an unlisted successor
    uint16_t cnt_r[kThreads];            // records / text bytes of the entry under its real carry
    uint32_t cnt_t[kThreads];
    uint8_t have_cnt[kThreads];
    uint16_t xcnt_r[kThreads];           // the same for the extension window
    uint32_t xcnt_t[kThreads];
    uint32_t queue[2 * kThreads];        // compacted work items of the divergent stages
};
constexpr uint32_t kNoWin = 0xFFFFFFFEu;

// entry index -> window index: binary search of the owning prefilter CTA's region (offsets staged in smem)
__device__ __forceinline__ long long list_window(const ExactCfg& X, const uint32_t* soff, long long e) {
    if (!X.list) return X.w_first + e;
    if (X.ncta == 0) return (long long)X.list[e];
    uint32_t lo = 0, hi = X.ncta;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((long long)soff[mid] <= e) lo = mid; else hi = mid;
    }
    return (long long)X.list[(unsigned long long)lo * X.region_stride + (unsigned long long)(e - soff[lo])];
}

// Second passes (replays under the real carry, extension windows) concern a minority of the entries; run by the
// entries' own lanes they would cost every warp a whole window pass for a handful of active lanes.  Instead the
// lanes queue them and the block's threads take the queued items densely.  Lane i contributes item `ia` (if pa)
// and item `ib` (if pb); returns the number of items (uniform).  Contains block barriers.
__device__ __forceinline__ uint32_t block_enqueue(ExactSmem& S, bool pa, uint32_t ia, bool pb, uint32_t ib) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ba = __ballot_sync(0xffffffffu, pa), bb = __ballot_sync(0xffffffffu, pb);
    if (lane == 0) S.warp_a[warp] = __popc(ba) + __popc(bb);
    __syncthreads();
    uint32_t off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const uint32_t c = S.warp_a[w];
        if ((uint32_t)w < warp) off += c;
        total += c;
    }
    const uint32_t lt = (1u << lane) - 1u;
    if (pa) S.queue[off + __popc(ba & lt)] = ia;
    if (pb) S.queue[off + __popc(ba) + __popc(bb & lt)] = ib;
    __syncthreads();
    return total;
}

// One pass over list entries [e0, e0 + nblk): summary, carry resolution and (full) emission.
// Returns the carry out of the last entry.
template <class Dec>
__device__ Carry block_pass(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const Geometry& geo, ExactSmem& S,
                            long long NE, long long e0, uint32_t nblk, bool full, Carry carry_in, long long block_id) {
    const uint32_t i = threadIdx.x;
    const GlobalSrc g{P.in, P.pend};
    const GlobalTile ts{g, P.len, X.in_aligned16 != 0, P.enc == ENC_UTF8 ? &S.tables : nullptr,
                        (uint32_t)__cvta_generic_to_shared(&S.tables.tt[0])};
    const bool active = i < nblk;
    long long w = -1;
    bool adj = false, next_adj = false;
    WinGeom wg;
    WinDesc d;
    d.type = WT_CONST; d.nrec = 0; d.ntext = 0; d.a = 0; d.t_out = 0; d.pad = 0; d.null_out = carry_none();

    constexpr bool kIsUtf8 = std::is_same<Dec, DecUtf8>::value;
    const bool use_mask = kIsUtf8 && !P.general;
    const int run_mode = full ? MODE_BUFFER : MODE_STATE;

    if (active) {
        const long long e = e0 + i;
        w = list_window(X, S.cta_off, e);
        S.win[i] = (uint32_t)w;
        if (i == 0) S.win_prev = e0 > 0 ? (uint32_t)list_window(X, S.cta_off, e0 - 1) : kNoWin;
        if (i == nblk - 1) S.win_next = e + 1 < NE ? (uint32_t)list_window(X, S.cta_off, e + 1) : kNoWin;
        S.have_cnt[i] = 0;
        S.out_done[i] = 0;
        S.in_known[i] = 0;
        S.declined[i] = 0;
        S.xcnt_r[i] = 0;
        S.xcnt_t[i] = 0;
    }
    __syncthreads();

    // An entry resolved under its real carry-in: counts, staged records, carry-out; propagates like a constant window.
    auto store_resolved = [&](uint32_t j, const Carry& kin_j, const WinResult& r) {
        S.kin[j] = kin_j;
        S.in_known[j] = 1;
        S.cnt_r[j] = (uint16_t)(r.nrec > 0xFFFFu ? 0xFFFFu : r.nrec);
        S.cnt_t[j] = r.ntext;
        S.have_cnt[j] = full ? 1 : 0;
        S.kout[j] = r.out;
        S.out_done[j] = 1;
        WinDesc dj;
        dj.type = WT_CONST; dj.pad = 0; dj.a = 0; dj.t_out = 0; dj.nrec = 0; dj.ntext = 0; dj.null_out = r.out;
        S.desc[j] = dj;
        if (j == nblk - 1) S.last_npend = r.npend_out;
        if (S.next_adj[j] && j + 1 < nblk) { S.kin[j + 1] = r.out; S.in_known[j + 1] = 1; }
    };

    // ---- stage A: isolated entries and heads of runs know their carry-in (pre-roll of the unlisted predecessor):
    //      one pass of the mask engine under the real carry gives carry-out, counts and staged records ------------
    if (active) {
        adj = (i > 0 ? S.win[i - 1] : S.win_prev) + 1u == (uint32_t)w;
        next_adj = (i + 1 < nblk ? S.win[i + 1] : S.win_next) == (uint32_t)w + 1u;
        S.next_adj[i] = next_adj ? 1 : 0;
        S.adj[i] = adj ? 1 : 0;
        geo.window(w, wg);
    }
    __syncthreads();
    if constexpr (kIsUtf8) {
        if (active && use_mask && !adj) {
            Carry kin0 = range_carry_in(P, X);
            bool ok = true;
            if (w != X.w_first) {
                const WinGeom rg = preroll_geom(geo, w, X.pre_bytes);
                WinResult rr;
                ok = utf8_mask_window(P, ts, rg, carry_none(), MODE_STATE, nullptr, 0, rr);
                kin0 = rr.out;
            }
            if (ok) {
                WinResult r;
                if (utf8_mask_window(P, ts, wg, kin0, run_mode, &S.staged[i][0], 0, r)) store_resolved(i, kin0, r);
            }
        }
        if (active && adj && i == 0) { S.kin[0] = carry_in; S.in_known[0] = 1; }
        // Members of short runs: as soon as the predecessor is resolved, the mask engine under the real carry
        // (queued, so the few members of a block share warps).  Long runs (text) skip this: their members are
        // summarised under the null carry below, all at once.
        const uint32_t n_adj = (uint32_t)__syncthreads_count(active && adj);
        if (use_mask && n_adj <= 48) {
            for (int round = 0; round < 3; ++round) {
                const bool rdy = active && adj && !S.out_done[i] && !S.declined[i] && S.in_known[i] && S.kin[i].kind != K_UNKNOWN;
                const uint32_t nq = block_enqueue(S, rdy, i, false, 0);
                if (nq == 0) break;
                if (i < nq) {
                    const uint32_t j = S.queue[i];
                    WinGeom wj;
                    geo.window((long long)S.win[j], wj);
                    const Carry kj = S.kin[j];
                    WinResult r;
                    if (utf8_mask_window(P, ts, wj, kj, run_mode, &S.staged[j][0], 0, r)) store_resolved(j, kj, r);
                    else S.declined[j] = 1;
                }
                __syncthreads();
            }
        }
    } else {
        if (active && adj && i == 0) { S.kin[0] = carry_in; S.in_known[0] = 1; }
        __syncthreads();
    }

    // ---- stage A': everything still open takes the byte-wise engine (queued): heads under their real carry,
    //      members of runs summarised under the null carry ------------------------------------------------------------
    const bool need_cls = active && !S.out_done[i];
    {
        const uint32_t nq = block_enqueue(S, need_cls, i, false, 0);
        if (i < nq) {
            const uint32_t j = S.queue[i];
            const long long wj_idx = (long long)S.win[j];
            const bool adj_j = S.adj[j] != 0;
            WinGeom wj;
            geo.window(wj_idx, wj);
            Carry kin0 = carry_none();
            if (!adj_j) {
                if (wj_idx == X.w_first) kin0 = range_carry_in(P, X);
                else {
                    const WinGeom rg = preroll_geom(geo, wj_idx, X.pre_bytes);
                    WinResult rr;
                    WindowEngine<Dec>::run(P, ts, g, rg, carry_none(), MODE_STATE, nullptr, 0, rr, nullptr);
                    kin0 = rr.out;
                }
            }
            WinResult r;
            WinDesc dsum;
            const int mode1 = adj_j ? MODE_COUNT : run_mode;
            WindowEngine<Dec>::run(P, ts, g, wj, kin0, mode1, &S.staged[j][0], 0, r, adj_j ? &dsum : nullptr);
            if (P.general && adj_j && dsum.type == WT_DEP)  // constant iff the flag handed to segment 2 does not matter
                dsum.type = classify_general(wj, r, [&](const WinGeom& g2) {
                    WinResult r2;
                    WindowEngine<Dec>::run(P, ts, g, g2, carry_none(), MODE_STATE, nullptr, 0, r2, nullptr);
                    return r2.out;
                });
            if (j == nblk - 1) S.last_npend = r.npend_out;
            if (!adj_j) {
                S.kin[j] = kin0;
                S.in_known[j] = 1;
                S.cnt_r[j] = (uint16_t)(r.nrec > 0xFFFFu ? 0xFFFFu : r.nrec);
                S.cnt_t[j] = r.ntext;
                S.have_cnt[j] = 1;
                dsum.type = WT_CONST;  // resolved: propagates like a constant window
                dsum.pad = 0; dsum.a = 0; dsum.t_out = 0; dsum.nrec = 0; dsum.ntext = 0;
                dsum.null_out = r.out;
            }
            S.desc[j] = dsum;
        }
        __syncthreads();
    }
    if (active) d = S.desc[i];
    if (need_cls && d.type == WT_CONST) {
        S.kout[i] = d.null_out;
        S.out_done[i] = 1;
        if (next_adj && i + 1 < nblk) { S.kin[i + 1] = d.null_out; S.in_known[i + 1] = 1; }
    }
    __syncthreads();

    // ---- stage B: resolve the carries along runs of adjacent windows (compacted replays) ---------------
    for (;;) {
        if (P.general) {
            // General missions: most windows are WT_GUARD (carry-out = the one seen under the null carry unless the
            // carried leftover fills the first run up); chains of them resolve without a pass, so one lane walks the
            // block in order.  With an unknown carry-in (warm-up of a block that starts inside a run) a guard window
            // behind a guard / constant one is still decided (guard_known_behind), which is what ends the warm-up.
            if (threadIdx.x == 0) {
                for (uint32_t j = 0; j < nblk; ++j) {
                    if (S.out_done[j] || !S.in_known[j]) continue;
                    const WinDesc dj = S.desc[j];
                    if (dj.type != WT_GUARD) continue;
                    const Carry kj = S.kin[j];
                    Carry out;
                    if (kj.kind == K_UNKNOWN) {
                        if (j > 0 && S.adj[j] && guard_known_behind(P, dj, S.desc[j - 1])) out = dj.null_out;
                        else out = kj;
                    } else if (guard_benign(P, dj, kj)) out = dj.null_out;
                    else continue;  // a leftover that fills the first run up: replay below
                    S.kout[j] = out;
                    S.out_done[j] = 1;
                    if (S.next_adj[j] && j + 1 < nblk) { S.kin[j + 1] = out; S.in_known[j + 1] = 1; }
                }
            }
            __syncthreads();
        }
        const bool rdy = active && !S.out_done[i] && S.in_known[i];
        const uint32_t nq = block_enqueue(S, rdy, i, false, 0);
        if (nq == 0) break;
        if (i < nq) {
            const uint32_t j = S.queue[i];
            const Carry kin = S.kin[j];
            const WinDesc dj = S.desc[j];
            WinGeom wj;
            geo.window((long long)S.win[j], wj);
            Carry out;
            if (kin.kind == K_UNKNOWN) out = kin;
            else if (dj.type == WT_CASEB) out = eval_caseb(P, dj, kin, (uint32_t)(wj.we - wj.ws));
            else {
                // replay under the real carry; in a full pass this also yields the counts and staged records
                WinResult r;
                WindowEngine<Dec>::run(P, ts, g, wj, kin, full ? MODE_BUFFER : MODE_STATE, &S.staged[j][0], 0, r, nullptr);
                out = r.out;
                if (full) {
                    S.cnt_r[j] = (uint16_t)(r.nrec > 0xFFFFu ? 0xFFFFu : r.nrec);
                    S.cnt_t[j] = r.ntext;
                    S.have_cnt[j] = 1;
                }
            }
            S.kout[j] = out;
            S.out_done[j] = 1;
            if (S.next_adj[j] && j + 1 < nblk) { S.kin[j + 1] = out; S.in_known[j + 1] = 1; }
        }
        __syncthreads();
    }
    const Carry carry_out = S.kout[nblk - 1];
    if (!full) {
        __syncthreads();
        return carry_out;
    }

    // ---- stage C: count, reserve, write --------------------------------------------------------------
    uint32_t cr = 0, ct = 0, xr = 0, xt = 0;
    bool need_emit = false, ext = false;
    Carry kin = carry_none(), kout = carry_none();
    if (active) {
        kin = S.kin[i];
        kout = S.kout[i];
        // member of a run whose carry-out came from its descriptor: one pass under the real carry, only if the
        // window can print at all (records are staged, so usually no write pass follows)
        if (!S.have_cnt[i]) need_emit = needs_emit(P, d, kin);
        // a "cut" carry (or a leftover already long enough to print) out of a listed window reaches an
        // unlisted successor: that window may print a continuation / the leftover
        ext = carry_needs_extension(P, kout) && !next_adj && (w + 1) < X.w_end;
    }
    {
        const uint32_t nq = block_enqueue(S, need_emit, i, ext, i | 0x10000u);
        for (uint32_t t = i; t < nq; t += kThreads) {
            const uint32_t item = S.queue[t], j = item & 0xFFFFu;
            const bool isx = (item >> 16) != 0;
            WinGeom wj;
            if (isx) wj = ext_geom(geo, (long long)S.win[j] + 1, X.pre_bytes);
            else geo.window((long long)S.win[j], wj);
            const Carry kj = isx ? S.kout[j] : S.kin[j];
            Record* const dst = isx ? &S.xstaged[j][0] : &S.staged[j][0];
            WinResult r;
            bool done = false;
            if constexpr (kIsUtf8) done = use_mask && utf8_mask_window(P, ts, wj, kj, MODE_BUFFER, dst, 0, r);
            if (!done) WindowEngine<Dec>::run(P, ts, g, wj, kj, MODE_BUFFER, dst, 0, r, nullptr);
            const uint16_t nr = (uint16_t)(r.nrec > 0xFFFFu ? 0xFFFFu : r.nrec);
            if (isx) { S.xcnt_r[j] = nr; S.xcnt_t[j] = r.ntext; }
            else { S.cnt_r[j] = nr; S.cnt_t[j] = r.ntext; S.have_cnt[j] = 1; }
        }
        __syncthreads();
    }
    WinGeom xg = wg;
    if (active) {
        if (S.have_cnt[i]) {
            cr = S.cnt_r[i]; ct = S.cnt_t[i];
            if (cr == 0xFFFFu) {  // more records than the 16-bit counter holds: count on the lane
                WinResult r;
                WindowEngine<Dec>::run(P, ts, g, wg, kin, MODE_COUNT, nullptr, 0, r, nullptr);
                cr = r.nrec; ct = r.ntext;
            }
        }
        if (ext) {
            xg = ext_geom(geo, w + 1, X.pre_bytes);
            xr = S.xcnt_r[i]; xt = S.xcnt_t[i];
            if (xr == 0xFFFFu) {
                WinResult r;
                WindowEngine<Dec>::run(P, ts, g, xg, kout, MODE_COUNT, nullptr, 0, r, nullptr);
                xr = r.nrec; xt = r.ntext;
            }
        }
    }
    const bool is_final = active && (e0 + i == NE - 1);
    const bool extra = is_final && range_is_tail(X) && kout.kind == K_L && kout.k > 0;  // the scanner's final leftover as a pseudo record
    uint32_t sum_r = cr + xr + (extra ? 1u : 0u), sum_t = ct + xt + (extra ? kout.out_bytes : 0u);
    uint32_t er, et, tr, tt;
    block_excl_scan2(sum_r, sum_t, S.warp_a, S.warp_b, er, et, tr, tt);
    if (i == 0) {
        unsigned long long br = 0, bt = 0;
        if (tr) br = atomicAdd(&O.counters[0], (unsigned long long)tr);
        if (tt) bt = atomicAdd(&O.counters[1], (unsigned long long)tt);
        S.bases[0] = br;
        S.bases[1] = bt;
        O.block_desc[block_id] = make_uint2((uint32_t)br, tr);
        if (br + tr > O.rec_cap || bt + tt > O.text_cap) O.final_state->overflow = 1;
    }
    __syncthreads();
    const unsigned long long br = S.bases[0], bt = S.bases[1];
    if ((br + tr <= O.rec_cap) && (bt + tt <= O.text_cap)) {
        uint32_t ro = er, to = et;
        if (cr) {
            if (cr <= kBufRecs) {  // staged by the counting pass: patch the text offsets and copy
                for (uint32_t k = 0; k < cr; ++k) {
                    Record r = S.staged[i][k];
                    r.text_off += bt + to;
                    O.recs[br + ro + k] = r;
                }
            } else {
                WinResult r;
                WindowEngine<Dec>::run(P, ts, g, wg, kin, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
            }
        }
        ro += cr; to += ct;
        if (xr) {
            if (xr <= kBufRecs) {
                for (uint32_t k = 0; k < xr; ++k) {
                    Record r = S.xstaged[i][k];
                    r.text_off += bt + to;
                    O.recs[br + ro + k] = r;
                }
            } else {
                WinResult r;
                WindowEngine<Dec>::run(P, ts, g, xg, kout, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
            }
        }
        ro += xr; to += xt;
        if (extra) {
            Record r;
            r.position = 0;
            r.in_start = P.len - (int64_t)kout.in_bytes;
            r.in_len = kout.in_bytes - (uint32_t)S.last_npend;
            r.text_len = kout.out_bytes;
            r.text_off = bt + to;
            r.flags = RF_LEFTOVER | ((kout.flags & CF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u) |
                      ((kout.flags & CF_HALF) ? (uint32_t)RF_HALFSTART : 0u);
            r.precision = 0;
            O.recs[br + ro] = r;
        }
    }
    if (is_final) { O.final_state->carry = kout; O.final_state->npend = S.last_npend; }
    __syncthreads();
    return carry_out;
}

template <class Dec>
__global__ void __launch_bounds__(kThreads, 4)
sx_exact_kernel(const __grid_constant__ ScanParams P, const ScanOut O, const ExactCfg X) {
    __shared__ ExactSmem S;
    Geometry geo;
    geo.init(P);
    const long long NE = X.list ? (long long)*X.ne_ptr : X.ne_static;
    if ((long long)blockIdx.x * kThreads >= NE) return;
    if (P.enc == ENC_UTF8)
        for (uint32_t k = threadIdx.x; k < 2048; k += kThreads) utf8_tables_fill(P, S.tables, k);
    if (X.list)
        for (uint32_t k = threadIdx.x; k <= X.ncta; k += kThreads) S.cta_off[k] = X.cta_off[k];
    __syncthreads();
    // Persistent CTAs: the number of list entries is only known on the device, so the grid is a few CTAs per SM
    // and every CTA claims 128-entry blocks from a device counter (block `b` owns entries [b * 128, b * 128 + 128)).
    for (;;) {
        if (threadIdx.x == 0) S.next_block = atomicAdd(&O.counters[3], 1ull);
        __syncthreads();
        const long long b = (long long)S.next_block;
        __syncthreads();
        if (b * kThreads >= NE) break;
        const long long e0 = b * kThreads;
        const uint32_t nblk = (uint32_t)((NE - e0) < (long long)kThreads ? (NE - e0) : (long long)kThreads);
        Carry c = carry_none();
        const bool first_adj = e0 > 0 && list_window(X, S.cta_off, e0 - 1) == list_window(X, S.cta_off, e0) - 1;
        if (first_adj) {
            // Warm-up: the block starts inside a run of adjacent windows.  Replay preceding entries in
            // state-only mode; a non-adjacent entry or a constant window makes the carry known.
            long long back = 8;
            for (;;) {
                long long es = e0 - back;
                if (es < 0) es = 0;
                c = carry_unknown();
                for (long long e = es; e < e0; e += kThreads) {
                    const uint32_t n = (uint32_t)((e0 - e) < (long long)kThreads ? (e0 - e) : (long long)kThreads);
                    c = block_pass<Dec>(P, O, X, geo, S, NE, e, n, false, c, 0);
                }
                if (c.kind != K_UNKNOWN || es == 0) break;
                back *= 4;
            }
        }
        block_pass<Dec>(P, O, X, geo, S, NE, e0, nblk, true, c, b);
    }
}


// Carry into the first window of a range that does not start at the stream start (pieces of a pipelined call, ranges
// of a stream sharded over several GPUs): walk back from the window before it to the nearest window whose carry-out
// does not depend on its carry-in -- the rule the exact stage already relies on (mask engine: WinResult.cut1 == 0;
// byte-wise engine: WinDesc WT_CONST) -- then replay forward under the real carries.  On binary input the window right
// before the range almost always qualifies; on text about 85 % of the windows do.  One warp per range, lane 0 works.
//   out[r]: the carry; fail[r] != 0: no such window within `max_back` windows, or the walk reached window 0 while the
//   carry at the stream start is not known (prefix_known == 0)
struct RangeCarryOut { Carry k0; uint32_t fail; uint32_t back; };
constexpr int kMaxRanges = 32;
struct RangeCarryArgs {
    long long w_first[kMaxRanges];
    uint32_t nranges, in_aligned16, prefix_known, max_back;
    RangeCarryOut* out;        // device; element r at byte offset r * out_stride
    size_t out_stride;
};
template <class Dec>
__global__ void __launch_bounds__(32)
sx_range_carry_kernel(const __grid_constant__ ScanParams P, const __grid_constant__ RangeCarryArgs A) {
    const uint32_t nranges = A.nranges, in_aligned16 = A.in_aligned16, prefix_known = A.prefix_known, max_back = A.max_back;
    RangeCarryOut* const out = A.out;
    const size_t out_stride = A.out_stride;
    __shared__ Utf8Tables T;
    constexpr bool kMask = MaskFamily<Dec>::kHas;
    if (kMask && blockIdx.x < A.nranges && A.w_first[blockIdx.x] > 0)  // a range at the stream start just takes P.k0
        for (uint32_t k = threadIdx.x; k < 2048; k += 32) mask_tables_fill(P, T, k);
    __syncthreads();
    if (threadIdx.x != 0 || blockIdx.x >= nranges) return;
    RangeCarryOut* const o = reinterpret_cast<RangeCarryOut*>(reinterpret_cast<uint8_t*>(out) + (size_t)blockIdx.x * out_stride);
    const long long wb = A.w_first[blockIdx.x];
    Geometry geo;
    geo.init(P);
    const GlobalSrc g{P.in, P.pend};
    const GlobalTile ts{g, P.len, in_aligned16 != 0, kMask ? &T : nullptr, (uint32_t)__cvta_generic_to_shared(&T.tt[0])};
    Carry c = P.k0;
    uint32_t fail = 0, back = 0;
    long long start = 0;
    if (wb > 0) {
        long long j = wb - 1;
        for (;;) {
            if (j < 0) { fail = prefix_known ? 0u : 1u; c = P.k0; start = 0; break; }
            if (back >= max_back) { fail = 1; break; }
            WinGeom wg;
            geo.window(j, wg);
            WinResult rr;
            bool indep;
            Carry co;
            bool done = false;
            if constexpr (kMask) {
                if (!P.general && mask_window<MaskFamily<Dec>::kSByte>(P, ts, wg, carry_none(), MODE_STATE, nullptr, 0, rr)) {
                    indep = rr.cut1 == 0; co = rr.out; done = true;
                }
            }
            if (!done) {
                WinDesc d;
                WindowEngine<Dec>::run(P, ts, g, wg, carry_none(), MODE_COUNT, nullptr, 0, rr, &d);
                if (P.general && d.type == WT_DEP)
                    d.type = classify_general(wg, rr, [&](const WinGeom& g2) {
                        WinResult r2;
                        WindowEngine<Dec>::run(P, ts, g, g2, carry_none(), MODE_STATE, nullptr, 0, r2, nullptr);
                        return r2.out;
                    });
                indep = d.type == WT_CONST; co = d.null_out;
            }
            ++back;
            if (indep) { c = co; start = j + 1; break; }
            --j;
        }
        if (!fail) {
            for (long long w = start; w < wb; ++w) {
                WinGeom wg;
                geo.window(w, wg);
                WinResult rr;
                bool done = false;
                if constexpr (kMask) done = !P.general && mask_window<MaskFamily<Dec>::kSByte>(P, ts, wg, c, MODE_STATE, nullptr, 0, rr);
                if (!done) WindowEngine<Dec>::run(P, ts, g, wg, c, MODE_STATE, nullptr, 0, rr, nullptr);
                c = rr.out;
            }
        }
    }
    o->k0 = c;
    o->fail = fail;
    o->back = back;
}

// Launchers, one set per decoder, each in its own translation unit (sx_exact_inst.cu compiled with -DSX_INST=<ENC_*>).
struct SparseBufs;
struct SparseLaunchCfg;
#define SX_DECLARE_INST(N)                                                                                                          \
    cudaError_t launch_exact_##N(const ScanParams& P, const ScanOut& O, const ExactCfg& X, unsigned grid, cudaStream_t st);         \
    cudaError_t launch_range_carry_##N(const ScanParams& P, const RangeCarryArgs& A, cudaStream_t st);                              \
    cudaError_t launch_sparse_##N(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B,                    \
                                  const SparseLaunchCfg& L, cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs); \
    bool has_sparse_##N();
SX_DECLARE_INST(0) SX_DECLARE_INST(1) SX_DECLARE_INST(2) SX_DECLARE_INST(3) SX_DECLARE_INST(4) SX_DECLARE_INST(5) SX_DECLARE_INST(6)
SX_DECLARE_INST(7) SX_DECLARE_INST(8)
#undef SX_DECLARE_INST
constexpr int kNumInst = 9;

}  // namespace sx
