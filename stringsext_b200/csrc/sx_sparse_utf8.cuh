// sx_sparse_utf8.cuh -- the exact stage for UTF-8 missions over a SPARSE window list (binary input: the
// prefilter keeps a few per cent of the windows, most of them isolated).
//
// sx_exact_kernel (sx_exact.cuh) resolves the carries of a block of entries with block barriers between its
// stages; on a sparse list almost every stage has a handful of busy lanes, so the kernel is bound by the latency
// of single warps (ncu: barrier stalls, profiles/).  Here every stage is its own data-parallel kernel -- one
// thread per list entry, no barrier on the data path -- and the per-entry results live in a device array:
//
//   sx_sp_heads_kernel   entries whose predecessor window is not listed: carry-in from the pre-roll, then ONE pass
//                        of the mask engine (sx_mask_utf8.cuh) -> carry-out, counts, first records
//   sx_sp_chains_kernel  runs of adjacent windows, walked in order by the thread of their first member under the
//                        real carries (mask engine, byte-wise engine when it declines); also heads the mask
//                        engine declined
//   sx_sp_ext_kernel     extension windows (a "cut" / long leftover carry reaching an unlisted successor),
//                        per-entry totals -> per-CTA totals
//   sx_sp_scan_kernel    exclusive scan of the per-CTA totals
//   sx_sp_gather_kernel  records and text offsets in stream order (exactly the order FindingCollection::from
//                        pushes them, finding_collection.rs:255-285), final carry for the ScannerState
//
// The host picks this path when the list is sparse (sx_scan.cu); dense lists (text) keep sx_exact_kernel, whose
// transfer-function classification resolves long runs of adjacent windows in parallel.
#pragma once
#include "sx_exact.cuh"
#include <algorithm>

namespace sx {

enum : uint8_t { ES_PENDING = 0, ES_DONE = 1, ES_DECLINED = 2 };
struct EntryState {
    Carry kin, kout;
    uint32_t cnt_r, cnt_t;    // records / text bytes of the window under its real carry
    uint32_t xcnt_r, xcnt_t;  // ... of the extension window
    uint8_t status;
    int8_t npend;             // bytes inside the decoder at the window end
    uint16_t pad0;
    uint32_t pad1;
    Record staged[kBufRecs];
    Record xstaged[kBufRecs];
};

struct SparseBufs {
    EntryState* E;
    ulonglong2* btot;          // per-CTA {records, text bytes}: totals, then exclusive prefix
    Utf8Tables* tables;        // filled by sx_sp_tables_kernel
    uint32_t* queue;           // work items of sx_sp_chains_kernel: first members of runs, declined lone heads (| kQueueHead)
    unsigned long long* qcount;
    long long NE;
};

constexpr int kSpThreads = 128;
constexpr uint32_t kQueueHead = 0x80000000u;

__global__ void __launch_bounds__(256) sx_sp_tables_kernel(const __grid_constant__ ScanParams P, Utf8Tables* T) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 2048) utf8_tables_fill(P, *T, k);
}

struct SpCtx {
    Geometry geo;
    GlobalSrc g;
    GlobalTile ts;
};
__device__ __forceinline__ void sp_setup(const ScanParams& P, const ExactCfg& X, const SparseBufs& B, Utf8Tables& S, SpCtx& c) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(B.tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S);
    for (uint32_t k = threadIdx.x; k < sizeof(Utf8Tables) / 4; k += blockDim.x) dst[k] = src[k];
    __syncthreads();
    c.geo.init(P);
    c.g = GlobalSrc{P.in, P.pend};
    c.ts = GlobalTile{c.g, P.len, X.in_aligned16 != 0, &S, (uint32_t)__cvta_generic_to_shared(&S.tt[0])};
}
__device__ __forceinline__ void sp_store(EntryState* es, const Carry& kin, const WinResult& r) {
    es->kin = kin;
    es->kout = r.out;
    es->cnt_r = r.nrec;
    es->cnt_t = r.ntext;
    es->npend = (int8_t)r.npend_out;
    es->status = ES_DONE;
}
// a head (predecessor window not listed) with the byte-wise engine: pre-roll, then one pass under the real carry
__device__ __noinline__ void sp_head_bytewise(const ScanParams& P, const ExactCfg& X, const SpCtx& c, long long w, EntryState* es) {
    WinGeom wg;
    c.geo.window(w, wg);
    Carry kin0 = P.k0;
    if (w != 0) {
        const WinGeom rg = preroll_geom(c.geo, w, X.pre_bytes);
        WinResult rr;
        WindowEngine<DecUtf8>::run(P, c.ts, c.g, rg, carry_none(), MODE_STATE, nullptr, 0, rr, nullptr);
        kin0 = rr.out;
    }
    WinResult r;
    WindowEngine<DecUtf8>::run(P, c.ts, c.g, wg, kin0, MODE_BUFFER, es->staged, 0, r, nullptr);
    sp_store(es, kin0, r);
}

// warp-aggregated append to the chain queue
__device__ __forceinline__ void sp_push(const SparseBufs& B, bool pred, uint32_t item) {
    const uint32_t m = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    const uint32_t lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(B.qcount, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    B.queue[base + __popc(m & ((1u << lane) - 1u))] = item;
}

__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_heads_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    const long long e = (long long)blockIdx.x * kSpThreads + threadIdx.x;
    const bool active = e < B.NE;
    long long w = 0, wp = -2;
    bool adj = false;
    EntryState* es = nullptr;
    if (active) {
        w = list_window(X, X.cta_off, e);
        wp = e > 0 ? list_window(X, X.cta_off, e - 1) : -2;
        adj = wp == w - 1;
        es = &B.E[e];
        es->xcnt_r = 0;
        es->xcnt_t = 0;
        es->status = ES_PENDING;
    }
    // the first member of a run of adjacent windows walks the run in sx_sp_chains_kernel
    bool chain_start = false;
    if (active && adj) chain_start = !(e > 1 && list_window(X, X.cta_off, e - 2) == wp - 1);
    sp_push(B, chain_start, (uint32_t)e);
    bool declined = false;
    if (active && !adj) {
        WinGeom wg;
        c.geo.window(w, wg);
        Carry kin0 = P.k0;
        bool ok = true;
        if (w != 0) {
            const WinGeom rg = preroll_geom(c.geo, w, X.pre_bytes);
            WinResult rr;
            ok = utf8_mask_window(P, c.ts, rg, carry_none(), MODE_STATE, nullptr, 0, rr);
            kin0 = rr.out;
        }
        WinResult r;
        if (ok) ok = utf8_mask_window(P, c.ts, wg, kin0, MODE_BUFFER, es->staged, 0, r);
        if (ok) sp_store(es, kin0, r);
        else {
            es->status = ES_DECLINED;
            // with a run behind it the run's walker resolves it, otherwise it is queued on its own
            declined = !(e + 1 < B.NE && list_window(X, X.cta_off, e + 1) == w + 1);
        }
    }
    sp_push(B, declined, (uint32_t)e | kQueueHead);
}

// Persistent over the queue: runs of adjacent windows walked under the real carries, declined heads.
__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_chains_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    const unsigned long long nq = *B.qcount;
    for (unsigned long long t = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x; t < nq;
         t += (unsigned long long)gridDim.x * kSpThreads) {
        const uint32_t item = B.queue[t];
        const long long e = (long long)(item & ~kQueueHead);
        const long long w = list_window(X, X.cta_off, e);
        if (item & kQueueHead) {
            sp_head_bytewise(P, X, c, w, &B.E[e]);
            continue;
        }
        if (B.E[e - 1].status != ES_DONE) sp_head_bytewise(P, X, c, w - 1, &B.E[e - 1]);
        Carry kin = B.E[e - 1].kout;
        long long m = e, wm = w;
        for (;;) {
            EntryState* const es = &B.E[m];
            WinGeom wg;
            c.geo.window(wm, wg);
            WinResult r;
            if (!utf8_mask_window(P, c.ts, wg, kin, MODE_BUFFER, es->staged, 0, r))
                WindowEngine<DecUtf8>::run(P, c.ts, c.g, wg, kin, MODE_BUFFER, es->staged, 0, r, nullptr);
            sp_store(es, kin, r);
            kin = r.out;
            ++m;
            if (m >= B.NE) break;
            const long long wn = list_window(X, X.cta_off, m);
            if (wn != wm + 1) break;
            wm = wn;
        }
    }
}

__device__ __forceinline__ void sp_block_sum2(unsigned long long& a, unsigned long long& b, unsigned long long* sa, unsigned long long* sb) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, d);
        b += __shfl_down_sync(0xffffffffu, b, d);
    }
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    a = 0; b = 0;
    for (uint32_t k = 0; k < blockDim.x / 32; ++k) { a += sa[k]; b += sb[k]; }
    __syncthreads();
}

__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_ext_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    __shared__ unsigned long long sa[kSpThreads / 32], sb[kSpThreads / 32];
    SpCtx c;
    sp_setup(P, X, B, T, c);
    const long long e = (long long)blockIdx.x * kSpThreads + threadIdx.x;
    unsigned long long sum_r = 0, sum_t = 0;
    if (e < B.NE) {
        EntryState* const es = &B.E[e];
        const Carry kout = es->kout;
        const long long w = list_window(X, X.cta_off, e);
        const bool next_adj = e + 1 < B.NE && list_window(X, X.cta_off, e + 1) == w + 1;
        uint32_t xr = 0, xt = 0;
        // a "cut" carry (or a leftover already long enough to print) out of a listed window reaches an unlisted
        // successor: that window may print a continuation / the leftover
        if (carry_needs_extension(P, kout) && !next_adj && (w + 1) < X.total_windows) {
            const WinGeom xg = ext_geom(c.geo, w + 1, X.pre_bytes);
            WinResult r;
            if (!utf8_mask_window(P, c.ts, xg, kout, MODE_BUFFER, es->xstaged, 0, r))
                WindowEngine<DecUtf8>::run(P, c.ts, c.g, xg, kout, MODE_BUFFER, es->xstaged, 0, r, nullptr);
            xr = r.nrec; xt = r.ntext;
            es->xcnt_r = xr;
            es->xcnt_t = xt;
        }
        const bool extra = (e == B.NE - 1) && kout.kind == K_L && kout.k > 0;  // the scanner's final leftover (pseudo record)
        sum_r = (unsigned long long)es->cnt_r + xr + (extra ? 1u : 0u);
        sum_t = (unsigned long long)es->cnt_t + xt + (extra ? kout.out_bytes : 0u);
    }
    sp_block_sum2(sum_r, sum_t, sa, sb);
    if (threadIdx.x == 0) B.btot[blockIdx.x] = make_ulonglong2(sum_r, sum_t);
}

// exclusive scan of the per-CTA totals, in place; grand totals -> counters[0], counters[1]
__global__ void __launch_bounds__(1024) sx_sp_scan_kernel(ulonglong2* btot, uint32_t nb, unsigned long long* counters) {
    __shared__ unsigned long long wa[32], wb[32];
    __shared__ unsigned long long carry_a, carry_b;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t k = base + tid;
        const ulonglong2 v = k < nb ? btot[k] : make_ulonglong2(0, 0);
        unsigned long long a = v.x, b = v.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long ta = __shfl_up_sync(0xffffffffu, a, d), tb = __shfl_up_sync(0xffffffffu, b, d);
            if (lane >= (uint32_t)d) { a += ta; b += tb; }
        }
        if (lane == 31) { wa[warp] = a; wb[warp] = b; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long xa = wa[lane], xb = wb[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long ta = __shfl_up_sync(0xffffffffu, xa, d), tb = __shfl_up_sync(0xffffffffu, xb, d);
                if (lane >= (uint32_t)d) { xa += ta; xb += tb; }
            }
            wa[lane] = xa; wb[lane] = xb;  // inclusive over warps
        }
        __syncthreads();
        const unsigned long long oa = carry_a + (warp ? wa[warp - 1] : 0ull), ob = carry_b + (warp ? wb[warp - 1] : 0ull);
        if (k < nb) btot[k] = make_ulonglong2(oa + a - v.x, ob + b - v.y);
        __syncthreads();
        if (tid == 1023) { carry_a = oa + a; carry_b = ob + b; }
        __syncthreads();
    }
    if (tid == 0) { counters[0] = carry_a; counters[1] = carry_b; }
}

__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_gather_kernel(const __grid_constant__ ScanParams P, const ScanOut O, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    __shared__ uint32_t wa[8], wb[8];
    SpCtx c;
    sp_setup(P, X, B, T, c);
    const long long e = (long long)blockIdx.x * kSpThreads + threadIdx.x;
    const bool active = e < B.NE;
    uint32_t cr = 0, ct = 0, xr = 0, xt = 0;
    bool extra = false;
    Carry kin = carry_none(), kout = carry_none();
    const EntryState* es = nullptr;
    if (active) {
        es = &B.E[e];
        cr = es->cnt_r; ct = es->cnt_t; xr = es->xcnt_r; xt = es->xcnt_t;
        kin = es->kin; kout = es->kout;
        extra = (e == B.NE - 1) && kout.kind == K_L && kout.k > 0;
    }
    const uint32_t sum_r = cr + xr + (extra ? 1u : 0u), sum_t = ct + xt + (extra ? kout.out_bytes : 0u);
    uint32_t er, et, tr, tt;
    {   // exclusive block scan (thread order)
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t ia = warp_incl_scan(sum_r), ib = warp_incl_scan(sum_t);
        if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
        __syncthreads();
        uint32_t oa = 0, ob = 0;
        tr = 0; tt = 0;
#pragma unroll
        for (int k = 0; k < kSpThreads / 32; ++k) {
            if ((uint32_t)k < warp) { oa += wa[k]; ob += wb[k]; }
            tr += wa[k]; tt += wb[k];
        }
        er = oa + ia - sum_r;
        et = ob + ib - sum_t;
    }
    const ulonglong2 base = B.btot[blockIdx.x];
    const unsigned long long br = base.x, bt = base.y;
    const bool fits = (br + tr <= O.rec_cap) && (bt + tt <= O.text_cap);
    if (!fits && threadIdx.x == 0) O.final_state->overflow = 1;
    if (active && fits) {
        uint32_t ro = er, to = et;
        if (cr) {
            if (cr <= kBufRecs) {
                for (uint32_t k = 0; k < cr; ++k) {
                    Record r = es->staged[k];
                    r.text_off += bt + to;
                    O.recs[br + ro + k] = r;
                }
            } else {
                WinGeom wg;
                c.geo.window(list_window(X, X.cta_off, e), wg);
                WinResult r;
                WindowEngine<DecUtf8>::run(P, c.ts, c.g, wg, kin, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
            }
        }
        ro += cr; to += ct;
        if (xr) {
            if (xr <= kBufRecs) {
                for (uint32_t k = 0; k < xr; ++k) {
                    Record r = es->xstaged[k];
                    r.text_off += bt + to;
                    O.recs[br + ro + k] = r;
                }
            } else {
                const WinGeom xg = ext_geom(c.geo, list_window(X, X.cta_off, e) + 1, X.pre_bytes);
                WinResult r;
                WindowEngine<DecUtf8>::run(P, c.ts, c.g, xg, kout, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
            }
        }
        ro += xr; to += xt;
        if (extra) {
            Record r;
            r.position = 0;
            r.in_start = P.len - (int64_t)kout.in_bytes;
            r.in_len = kout.in_bytes - (uint32_t)es->npend;
            r.text_len = kout.out_bytes;
            r.text_off = bt + to;
            r.flags = RF_LEFTOVER | ((kout.flags & CF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u);
            r.precision = 0;
            O.recs[br + ro] = r;
        }
    }
    if (active && e == B.NE - 1) { O.final_state->carry = kout; O.final_state->npend = es->npend; }
}

inline cudaError_t launch_sparse_utf8_impl(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B, int num_sms,
                                           cudaStream_t st) {
    const unsigned nb = (unsigned)((B.NE + kSpThreads - 1) / kSpThreads);
    sx_sp_tables_kernel<<<8, 256, 0, st>>>(P, B.tables);
    sx_sp_heads_kernel<<<nb, kSpThreads, 0, st>>>(P, X, B);
    sx_sp_chains_kernel<<<std::min<unsigned>(nb, (unsigned)num_sms * 4u), kSpThreads, 0, st>>>(P, X, B);
    sx_sp_ext_kernel<<<nb, kSpThreads, 0, st>>>(P, X, B);
    sx_sp_scan_kernel<<<1, 1024, 0, st>>>(B.btot, nb, O.counters);
    sx_sp_gather_kernel<<<nb, kSpThreads, 0, st>>>(P, O, X, B);
    return cudaGetLastError();
}
constexpr uint32_t kSparseLaunches = 6;

}  // namespace sx
