// sx_sparse_utf8.cuh -- the exact stage of plain missions as a pipeline of data-parallel kernels, one thread per listed
// window ("entry"), no barrier on the data path.  UTF-8 and the single-byte family run the bit-parallel window engine
// (sx_mask_utf8.cuh) in it; UTF-16/32, Big5 and EUC-JP "decline" everywhere and run the convergent byte-wise engine
// (sx_fast_generic.cuh) through the fallbacks every stage has.
//
// sx_exact_kernel (sx_exact.cuh) resolves the carries of a block of entries with block barriers between its stages;
// on the sparse lists of binary input almost every stage has a handful of busy lanes, so the kernel is bound by the
// latency of single warps (ncu: barrier stalls, profiles/).  Here every stage is its own kernel and the per-entry
// results live in device arrays (EntryHot + staged records):
//
//   sx_list_compact_kernel  the prefilter CTAs' list regions -> one contiguous list of the piece; entry count
//   sx_sp_heads_kernel   queues the members of runs of adjacent windows (entries whose predecessor window is listed);
//                        entries whose predecessor window is not listed: carry-in from the 32 bytes in front of the
//                        window (the pre-roll folded into the mask frame), ONE pass of the mask engine
//                        (sx_mask_utf8.cuh) -> carry-out, counts, first records
//   sx_sp_declined_kernel heads the mask engine declined (byte-wise engine): a few, on a side stream beside the members,
//                        for the mask-engine decoders; ALL heads, in stream order in front of the members, for the others
//   sx_sp_members_kernel members, in parallel: carry-in = carry-out of the entry before, taken from a resolved head or
//                        recomputed from that window alone when it does not depend on ITS carry-in (WinResult.cut1 /
//                        WinDesc.type == WT_CONST)
//   sx_sp_fix_kernel     the rest, few: members behind a carry-dependent window (walked in order; a window that is one
//                        short run is passed in closed form, eval_caseb, and resolved afterwards by sx_sp_late_kernel)
//   sx_sp_ext_kernel     extension windows (a "cut" / long leftover carry reaching an unlisted successor),
//                        per-entry totals -> per-chunk totals
//   sx_sp_scan_kernel    exclusive scan of the per-chunk totals on top of the totals of the pieces before this one
//   sx_sp_gather_kernel  records in stream order (exactly the order FindingCollection::from pushes them,
//                        finding_collection.rs:255-285) as 8-byte wire records + their UTF-8 text, back to back, one
//                        explicit text offset per page of 4096 records: straight into the collection's pinned,
//                        device-mapped memory (little output) or into device staging that the copy engine moves part
//                        by part (output-heavy calls); final carry for the ScannerState
//
// A call may be cut into PIECES (window ranges, sx_scan.cu) behind ONE prefilter launch: each piece runs this pipeline
// on its own slices of the arrays below; the carry into a piece's first window comes from sx_range_carry_kernel
// (sx_exact.cuh), so pieces do not wait for each other except for the record offsets.
//
// Text-like input (every window listed, several findings per window) takes the same path: most windows' carry-out is
// independent of their carry-in, so the members resolve in parallel too.  The block kernel remains for general missions
// and for scans without the prefilter (sx_scan.cu).
#pragma once
#include "sx_exact.cuh"
#include <algorithm>

namespace sx {

enum : uint8_t { ES_PENDING = 0, ES_DONE = 1, ES_DECLINED = 2 /* head */, ES_DEPENDENT = 3 /* member of a run */ };
// Per-entry state every stage reads and writes: one 64-byte line.  The records a window stages (kBufRecs of its own,
// kBufRecs of its extension window) and the closed-form descriptor's null carry live in separate arrays that are only
// touched when there is something to store -- most listed windows of binary input print nothing.
struct EntryHot {
    Carry kin, kout;
    uint32_t cnt_r, cnt_t;    // records / text bytes of the window under its real carry
    uint32_t xcnt_r, xcnt_t;  // ... of the extension window
    uint8_t status;
    int8_t npend;             // bytes inside the decoder at the window end
    uint8_t resolved;         // kin / kout / counts / staged records are final
    uint8_t kin_known;        // sx_sp_fix_kernel: kin is final, the window itself is resolved by sx_sp_late_kernel
    // closed-form transfer function of the window (written by the thread of the NEXT entry, sx_sp_members_kernel):
    // one run of d_a (< q) passing chars with d_t text bytes covering the window, eval_caseb(); null carry in dnull[]
    uint8_t d_caseb, pad0;
    uint16_t d_a, d_t, pad1;
    uint32_t pad2;
};
static_assert(sizeof(EntryHot) == 64, "EntryHot is one 64-byte line");

// Device-resident control block of one piece.
struct PieceCtl {
    unsigned long long ne;        // list entries of the piece (0 when they do not fit the entry arrays: overflow)
    unsigned long long ne_raw;    // windows the prefilter kept
    unsigned long long qcount, qcount2, qcount2_snap;
    unsigned long long nrec, ntext;          // findings / text bytes of this piece
    unsigned long long rec_base, text_base;  // ... of all pieces before it
    uint32_t overflow;            // 1: entries > capacity, 2: records / text beyond the output buffers
    uint32_t text_fallback;       // some CTA left its text to sx_materialize_kernel
    RangeCarryOut rc;             // carry into the piece's first window (sx_range_carry_kernel); rc.fail: none found
};
// What the host reads per piece (pinned, device-mapped; written by sx_sp_scan_kernel / sx_sp_gather_kernel).
constexpr int kGatherParts = 4;
struct PieceSummary {
    unsigned long long nrec, ntext, rec_base, text_base, ne_raw;
    uint32_t overflow, text_fallback;
    // the gather kernel is launched in parts (chunk ranges) so that the download of a part overlaps the gathering of the
    // next one: cumulative record / text offsets at the END of every part
    unsigned long long part_rec_end[kGatherParts], part_text_end[kGatherParts];
};

struct SparseBufs {
    EntryHot* H;
    Carry* dnull;              // per entry: null carry of its closed-form descriptor (d_caseb)
    Record* staged;            // per entry kBufRecs records
    Record* xstaged;           // per entry kBufRecs records of the extension window
    ulonglong2* btot;          // per chunk of kSpThreads entries {records, text bytes}: totals, then exclusive prefix
    Utf8Tables* tables;        // filled by sx_sp_tables_kernel
    uint32_t* queue;           // members of runs of adjacent windows (sx_sp_members_kernel)
    uint32_t* queue2;          // what is left for sx_sp_fix_kernel: declined heads, members with a carry-dependent predecessor
    PieceCtl* ctl;             // this piece
    const PieceCtl* prev;      // the piece before it (nullptr: first piece of the call)
    PieceSummary* summary;     // host-visible
    uint8_t* text_out;         // device text arena (staging of the finding text)
};

constexpr int kSpThreads = 128;
constexpr uint32_t kQueueHead = 0x80000000u;

static __global__ void __launch_bounds__(256) sx_sp_tables_kernel(const __grid_constant__ ScanParams P, Utf8Tables* T) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 2048) mask_tables_fill(P, *T, k);
}

// The list regions of the prefilter CTAs of one piece -> a contiguous list; entry count and queue counters of the piece.
static __global__ void __launch_bounds__(256)
sx_list_compact_kernel(const uint32_t* __restrict__ cta_count, uint32_t ncta, const uint32_t* __restrict__ list, long long tile0,
                       long long tiles_per_cta, uint32_t* __restrict__ out, PieceCtl* ctl, unsigned long long cap) {
    __shared__ unsigned long long s_pre[8], s_tot[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    unsigned long long pre = 0, tot = 0;
    for (uint32_t k = tid; k < ncta; k += 256) {
        const uint32_t c = cta_count[k];
        tot += c;
        if (k < b) pre += c;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        pre += __shfl_down_sync(0xffffffffu, pre, d);
        tot += __shfl_down_sync(0xffffffffu, tot, d);
    }
    if (lane == 0) { s_pre[warp] = pre; s_tot[warp] = tot; }
    __syncthreads();
    pre = 0; tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { pre += s_pre[k]; tot += s_tot[k]; }
    if (b == 0 && tid == 0) {
        ctl->ne_raw = tot;
        ctl->ne = tot <= cap ? tot : 0ull;
        if (tot > cap) ctl->overflow |= 1u;
        ctl->qcount = 0; ctl->qcount2 = 0; ctl->qcount2_snap = 0;
    }
    if (tot > cap) return;
    const uint32_t n = cta_count[b];
    const uint32_t* src = list + (size_t)(tile0 + (long long)b * tiles_per_cta) * kPrefTileWin;
    for (uint32_t i = tid; i < n; i += 256) out[pre + i] = src[i];
}

// The bit-parallel engine where the decoder has one (UTF-8, single-byte family); every other decoder "declines", i.e. all
// its windows take the byte-wise engine through the fallbacks every stage has anyway -- still one thread per entry and
// no block barrier, which is what makes this pipeline faster than the block kernel on long lists.
template <class Dec, class TileSrc>
__device__ __forceinline__ bool sp_mask_window(const ScanParams& P, const TileSrc& ts, const WinGeom& geo, const Carry& kin, int mode,
                                               Record* wr, uint64_t text_off, WinResult& res) {
    if constexpr (MaskFamily<Dec>::kHas) return mask_window<MaskFamily<Dec>::kSByte>(P, ts, geo, kin, mode, wr, text_off, res);
    else return false;
}
template <class Dec, class TileSrc>
__device__ __forceinline__ bool sp_mask_head(const ScanParams& P, const TileSrc& ts, const WinGeom& geo, uint32_t pre_bytes, int mode,
                                             Record* wr, uint64_t text_off, WinResult& res) {
    if constexpr (MaskFamily<Dec>::kHas) return mask_head<MaskFamily<Dec>::kSByte>(P, ts, geo, pre_bytes, mode, wr, text_off, res);
    else return false;
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
    // decoders without tables (UTF-16 / UTF-32 / Big5 / EUC-JP) get none: WindowEngine<Dec> then takes the generic automaton
    c.ts = GlobalTile{c.g, P.len, X.in_aligned16 != 0, (P.enc == ENC_UTF8 || P.enc == ENC_XUD || P.enc == ENC_SB) ? &S : nullptr,
                      (uint32_t)__cvta_generic_to_shared(&S.tt[0])};
}
__device__ __forceinline__ Record* sp_staged(const SparseBufs& B, long long e) { return B.staged + (size_t)e * kBufRecs; }
__device__ __forceinline__ Record* sp_xstaged(const SparseBufs& B, long long e) { return B.xstaged + (size_t)e * kBufRecs; }
__device__ __forceinline__ void sp_store(EntryHot* es, const Carry& kin, const WinResult& r) {
    es->kin = kin;
    es->kout = r.out;
    es->cnt_r = r.nrec;
    es->cnt_t = r.ntext;
    es->npend = (int8_t)r.npend_out;
    es->resolved = 1;
    if (es->status == ES_PENDING) es->status = ES_DONE;  // sx_sp_fix_kernel leaves DECLINED / DEPENDENT as they are
}
// a head (predecessor window not listed) with the byte-wise engine: pre-roll, then one pass under the real carry
template <class Dec>
__device__ __noinline__ void sp_head_bytewise(const ScanParams& P, const ExactCfg& X, const SparseBufs& B, const SpCtx& c, long long w,
                                              long long e) {
    WinGeom wg;
    c.geo.window(w, wg);
    Carry kin0 = range_carry_in(P, X);
    if (w != X.w_first) {
        const WinGeom rg = preroll_geom(c.geo, w, X.pre_bytes);
        WinResult rr;
        WindowEngine<Dec>::run(P, c.ts, c.g, rg, carry_none(), MODE_STATE, nullptr, 0, rr, nullptr);
        kin0 = rr.out;
    }
    WinResult r;
    WindowEngine<Dec>::run(P, c.ts, c.g, wg, kin0, MODE_BUFFER, sp_staged(B, e), 0, r, nullptr);
    sp_store(&B.H[e], kin0, r);
}

// warp-aggregated append to a work queue
__device__ __forceinline__ void sp_push(uint32_t* queue, unsigned long long* qcount, bool pred, uint32_t item) {
    const uint32_t m = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    const uint32_t lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(qcount, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    queue[base + __popc(m & ((1u << lane) - 1u))] = item;
}

// Entry roles: a HEAD has no listed predecessor window (its carry-in comes from the pre-roll); a MEMBER continues a
// run of adjacent windows (its carry-in is the carry-out of the entry before it).
// a head with the mask engine (pre-roll + one pass under the real carry); false when the engine declines
template <class Dec>
__device__ __forceinline__ bool sp_head_mask(const ScanParams& P, const ExactCfg& X, const SparseBufs& B, const SpCtx& c, long long w,
                                             long long e) {
    WinGeom wg;
    c.geo.window(w, wg);
    WinResult r;
    EntryHot* const es = &B.H[e];
    // one pass: the pre-roll region is the 32 bytes in front of the window (same class planes, same decoder algebra)
    if (w != X.w_first && sp_mask_head<Dec>(P, c.ts, wg, X.pre_bytes, MODE_BUFFER, sp_staged(B, e), 0, r)) {
        sp_store(es, r.in, r);
        return true;
    }
    Carry kin0 = range_carry_in(P, X);
    if (w != X.w_first) {
        const WinGeom rg = preroll_geom(c.geo, w, X.pre_bytes);
        WinResult rr;
        if (!sp_mask_window<Dec>(P, c.ts, rg, carry_none(), MODE_STATE, nullptr, 0, rr)) return false;
        kin0 = rr.out;
    }
    if (!sp_mask_window<Dec>(P, c.ts, wg, kin0, MODE_BUFFER, sp_staged(B, e), 0, r)) return false;
    sp_store(es, kin0, r);
    return true;
}
// a member under its real carry: mask engine, byte-wise engine when it declines
template <class Dec>
__device__ __forceinline__ Carry sp_member(const ScanParams& P, const SparseBufs& B, const SpCtx& c, long long w, const Carry& kin,
                                           long long e) {
    WinGeom wg;
    c.geo.window(w, wg);
    WinResult r;
    if (!sp_mask_window<Dec>(P, c.ts, wg, kin, MODE_BUFFER, sp_staged(B, e), 0, r))
        WindowEngine<Dec>::run(P, c.ts, c.g, wg, kin, MODE_BUFFER, sp_staged(B, e), 0, r, nullptr);
    sp_store(&B.H[e], kin, r);
    return r.out;
}

// The kernels are persistent over the piece's entries (the entry count only exists on the device): chunk = kSpThreads
// consecutive entries, CTA b takes chunks b, b + gridDim.x, ...
// 6 CTAs of 128 threads per SM (80 registers): measured best for this latency-bound kernel (4: 0.35 ms, 6: 0.29 ms, 8: 0.39 ms
// before the one-pass heads)
template <class Dec, int MINB>
__global__ void __launch_bounds__(kSpThreads, MINB)
sx_sp_heads_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    const long long NE = (long long)B.ctl->ne;
    if ((long long)blockIdx.x * kSpThreads >= NE) return;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    for (long long e0 = (long long)blockIdx.x * kSpThreads; e0 < NE; e0 += (long long)gridDim.x * kSpThreads) {
        const long long e = e0 + threadIdx.x;
        bool declined = false, member = false;
        if (e < NE) {
            const long long w = list_window(X, X.cta_off, e);
            EntryHot* const es = &B.H[e];
            es->xcnt_r = 0;
            es->xcnt_t = 0;
            es->status = ES_PENDING;
            es->resolved = 0;
            es->kin_known = 0;
            es->d_caseb = 0;
            member = e > 0 && list_window(X, X.cta_off, e - 1) == w - 1;
            if (!member && !sp_head_mask<Dec>(P, X, B, c, w, e)) { es->status = ES_DECLINED; declined = true; }
        }
        sp_push(B.queue, &B.ctl->qcount, member, (uint32_t)e);  // resolved by sx_sp_members_kernel
        sp_push(B.queue2, &B.ctl->qcount2, declined, (uint32_t)e);
    }
}
// queue2 holds the declined heads now (the members kernel appends its dependent members behind them): snapshot its length
static __global__ void sx_sp_snapshot_kernel(PieceCtl* ctl) { ctl->qcount2_snap = ctl->qcount2; }

// Members of runs, all at once: the carry-in of a member is the carry-out of the entry before it -- known when that
// entry is a resolved head, and computable from the window alone when its carry-out does not depend on its own
// carry-in (mask engine under the null carry, WinResult.cut1 == 0).  What remains is walked in order by
// sx_sp_fix_kernel.  Persistent over the queue.
template <class Dec, int MINB>
__global__ void __launch_bounds__(kSpThreads, MINB)
sx_sp_members_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    const unsigned long long nq = B.ctl->qcount;
    if ((unsigned long long)blockIdx.x * kSpThreads >= nq) return;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    const unsigned long long stride = (unsigned long long)gridDim.x * kSpThreads;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * kSpThreads; t0 < nq; t0 += stride) {
        const unsigned long long t = t0 + threadIdx.x;
        bool dependent = false;
        long long e = 0;
        if (t < nq) {
            e = (long long)B.queue[t];
            const long long w = list_window(X, X.cta_off, e);
            const bool pred_is_member = e > 1 && list_window(X, X.cta_off, e - 2) == w - 2;
            Carry kin = carry_none();
            bool known = false;
            if (!pred_is_member) {
                // byte-wise decoders: sx_sp_declined_kernel has run (stream order), the head keeps its ES_DECLINED status
                known = B.H[e - 1].status == ES_DONE || (!MaskFamily<Dec>::kHas && B.H[e - 1].resolved);
                if (known) kin = B.H[e - 1].kout;
            } else {
                WinGeom pg;
                c.geo.window(w - 1, pg);
                WinResult rr;
                if (sp_mask_window<Dec>(P, c.ts, pg, carry_none(), MODE_STATE, nullptr, 0, rr)) {
                    known = rr.cut1 == 0;
                    kin = rr.out;
                    if (rr.caseb) {  // the walk of sx_sp_fix_kernel gets through that window without a pass
                        EntryHot* const pe = &B.H[e - 1];
                        pe->d_a = rr.a; pe->d_t = rr.t_out; B.dnull[e - 1] = rr.out; pe->d_caseb = 1;
                    }
                } else {
                    // the mask engine declines (e.g. a window crowded with findings): the byte-wise engine's
                    // transfer-function summary decides, so that such windows do not chain up in sx_sp_fix_kernel
                    WinDesc d;
                    WindowEngine<Dec>::run(P, c.ts, c.g, pg, carry_none(), MODE_COUNT, nullptr, 0, rr, &d);
                    known = d.type == WT_CONST;
                    kin = d.null_out;
                    if (d.type == WT_CASEB) {
                        EntryHot* const pe = &B.H[e - 1];
                        pe->d_a = d.a; pe->d_t = d.t_out; B.dnull[e - 1] = d.null_out; pe->d_caseb = 1;
                    }
                }
            }
            if (known) sp_member<Dec>(P, B, c, w, kin, e);
            else { B.H[e].status = ES_DEPENDENT; dependent = true; }
        }
        sp_push(B.queue2, &B.ctl->qcount2, dependent, (uint32_t)e);
    }
}

// Heads the mask engine declined (rare: 0.04 % of the entries on binary input), byte-wise.  The byte-wise engine takes
// ~0.1 ms for one window on a lone warp, so this kernel runs on a side stream beside sx_sp_members_kernel instead of
// sitting in front of the walks of sx_sp_fix_kernel.  Persistent over the first qcount2_snap items of queue2 (the
// declined heads).
template <class Dec, int MINB>
__global__ void __launch_bounds__(kSpThreads, MINB)
sx_sp_declined_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    const unsigned long long nq = B.ctl->qcount2_snap;
    if ((unsigned long long)blockIdx.x * kSpThreads >= nq) return;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    for (unsigned long long t = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x; t < nq;
         t += (unsigned long long)gridDim.x * kSpThreads) {
        const long long e = (long long)B.queue2[t];
        sp_head_bytewise<Dec>(P, X, B, c, list_window(X, X.cta_off, e), e);
    }
}

// What the parallel stages left: heads the mask engine declined (byte-wise engine) and members whose carry-in needs
// the entry before them resolved first -- walked in stream order from the first member whose predecessor is resolved.
// Entry statuses are frozen here (decisions only read what the earlier kernels wrote).  Persistent over queue2.
template <class Dec>
__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_fix_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    const unsigned long long nq = B.ctl->qcount2;
    if ((unsigned long long)blockIdx.x * kSpThreads >= nq) return;
    const long long NE = (long long)B.ctl->ne;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    for (unsigned long long t = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x; t < nq;
         t += (unsigned long long)gridDim.x * kSpThreads) {
        const long long e = (long long)B.queue2[t];
        const long long w = list_window(X, X.cta_off, e);
        if (B.H[e].status == ES_DECLINED) continue;  // a head: resolved by sx_sp_declined_kernel
        const uint8_t ps = B.H[e - 1].status;
        if (ps == ES_DEPENDENT) continue;  // the walk that started further left comes through here
        Carry kin = B.H[e - 1].kout;
        long long m = e, wm = w;
        for (;;) {
            EntryHot* const es = &B.H[m];
            if (es->d_caseb) {
                // one short run covering the window: its carry-out follows in closed form, the window itself (its
                // findings under this carry-in) is left to sx_sp_late_kernel, in parallel with the others
                es->kin = kin;
                es->kin_known = 1;
                WinDesc d;
                d.type = WT_CASEB; d.pad = 0; d.a = es->d_a; d.t_out = es->d_t; d.nrec = 0; d.ntext = 0; d.null_out = B.dnull[m];
                WinGeom wg;
                c.geo.window(wm, wg);
                kin = eval_caseb(P, d, kin, (uint32_t)(wg.we - wg.ws));
            } else kin = sp_member<Dec>(P, B, c, wm, kin, m);
            ++m;
            if (m >= NE || B.H[m].status != ES_DEPENDENT) break;
            const long long wn = list_window(X, X.cta_off, m);
            if (wn != wm + 1) break;
            wm = wn;
        }
    }
}

// members whose carry-in sx_sp_fix_kernel settled in closed form: resolved here, all at once
template <class Dec>
__global__ void __launch_bounds__(kSpThreads, 4)
sx_sp_late_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    const unsigned long long nq = B.ctl->qcount2;
    if ((unsigned long long)blockIdx.x * kSpThreads >= nq) return;
    SpCtx c;
    sp_setup(P, X, B, T, c);
    for (unsigned long long t = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x; t < nq;
         t += (unsigned long long)gridDim.x * kSpThreads) {
        const long long e = (long long)B.queue2[t];
        EntryHot* const es = &B.H[e];
        if (es->kin_known && !es->resolved) sp_member<Dec>(P, B, c, list_window(X, X.cta_off, e), es->kin, e);
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

// is the window behind entry e (window w) listed?  The first window of the NEXT range is, by construction.
__device__ __forceinline__ bool sp_next_adjacent(const ExactCfg& X, long long NE, long long e, long long w) {
    if (e + 1 < NE) return list_window(X, X.cta_off, e + 1) == w + 1;
    return w + 1 == X.w_end && !range_is_tail(X);
}

template <class Dec, int MINB>
__global__ void __launch_bounds__(kSpThreads, MINB)
sx_sp_ext_kernel(const __grid_constant__ ScanParams P, const ExactCfg X, const SparseBufs B) {
    __shared__ Utf8Tables T;
    __shared__ unsigned long long sa[kSpThreads / 32], sb[kSpThreads / 32];
    const long long NE = (long long)B.ctl->ne;
    if ((long long)blockIdx.x * kSpThreads >= NE) return;
    SpCtx c;
    bool ready = false;  // the tables are only needed by the few chunks that hold an extension window
    for (long long e0 = (long long)blockIdx.x * kSpThreads; e0 < NE; e0 += (long long)gridDim.x * kSpThreads) {
        const long long e = e0 + threadIdx.x;
        unsigned long long sum_r = 0, sum_t = 0;
        bool need_x = false;
        long long w = 0;
        Carry kout = carry_none();
        EntryHot* es = nullptr;
        if (e < NE) {
            es = &B.H[e];
            w = list_window(X, X.cta_off, e);
            kout = es->kout;
            // a "cut" carry (or a leftover already long enough to print) out of a listed window reaches an unlisted
            // successor: that window may print a continuation / the leftover
            need_x = carry_needs_extension(P, kout) && (w + 1) < X.w_end && !sp_next_adjacent(X, NE, e, w);
        }
        if (__syncthreads_or(need_x) && !ready) { sp_setup(P, X, B, T, c); ready = true; }
        if (e < NE) {
            uint32_t xr = 0, xt = 0;
            if (need_x) {
                const WinGeom xg = ext_geom(c.geo, w + 1, X.pre_bytes);
                WinResult r;
                if (!sp_mask_window<Dec>(P, c.ts, xg, kout, MODE_BUFFER, sp_xstaged(B, e), 0, r))
                    WindowEngine<Dec>::run(P, c.ts, c.g, xg, kout, MODE_BUFFER, sp_xstaged(B, e), 0, r, nullptr);
                xr = r.nrec; xt = r.ntext;
                es->xcnt_r = xr;
                es->xcnt_t = xt;
            }
            // the scanner's final leftover (pseudo record)
            const bool extra = (e == NE - 1) && range_is_tail(X) && kout.kind == K_L && kout.k > 0;
            sum_r = (unsigned long long)es->cnt_r + xr + (extra ? 1u : 0u);
            sum_t = (unsigned long long)es->cnt_t + xt + (extra ? kout.out_bytes : 0u);
        }
        sp_block_sum2(sum_r, sum_t, sa, sb);
        if (threadIdx.x == 0) B.btot[e0 / kSpThreads] = make_ulonglong2(sum_r, sum_t);
    }
}

// exclusive scan of the per-chunk totals, in place, starting from the totals of the pieces before this one; the piece's
// totals -> its control block and the host-visible summary
static __global__ void __launch_bounds__(1024) sx_sp_scan_kernel(const SparseBufs B, unsigned long long rec_cap, unsigned long long text_cap,
                                                                 unsigned long long out_cap, uint32_t nparts) {
    __shared__ unsigned long long wa[32], wb[32];
    __shared__ unsigned long long carry_a, carry_b;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ulonglong2* const btot = B.btot;
    const unsigned long long ne = B.ctl->ne;
    const uint32_t nb = (uint32_t)((ne + kSpThreads - 1) / kSpThreads);
    const unsigned long long base_a = B.prev ? B.prev->rec_base + B.prev->nrec : 0ull;
    const unsigned long long base_b = B.prev ? B.prev->text_base + B.prev->ntext : 0ull;
    if (tid == 0) { carry_a = base_a; carry_b = base_b; }
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
    if (tid == 0) {
        PieceCtl* const ctl = B.ctl;
        ctl->rec_base = base_a; ctl->text_base = base_b;
        ctl->nrec = carry_a - base_a; ctl->ntext = carry_b - base_b;
        if (B.prev && (B.prev->overflow != 0)) ctl->overflow |= 4u;  // an earlier piece already failed
        if (carry_a > rec_cap || carry_a > out_cap || carry_b > text_cap) ctl->overflow |= 2u;
        if (ctl->rc.fail) ctl->overflow |= 0x100u;
        PieceSummary s;
        s.nrec = ctl->nrec; s.ntext = ctl->ntext; s.rec_base = base_a; s.text_base = base_b; s.ne_raw = ctl->ne_raw;
        s.overflow = ctl->overflow; s.text_fallback = 0;
        for (int i = 0; i < kGatherParts; ++i) {
            const uint32_t cb = (uint32_t)(((unsigned long long)nb * (unsigned)(i + 1)) / (unsigned)nparts);
            const bool last = i + 1 >= (int)nparts || cb >= nb;
            s.part_rec_end[i] = last ? carry_a : btot[cb].x;
            s.part_text_end[i] = last ? carry_b : btot[cb].y;
        }
        *B.summary = s;
    }
}

template <class Dec, int MINB>
__global__ void __launch_bounds__(kSpThreads, MINB)
sx_sp_gather_kernel(const __grid_constant__ ScanParams P, const ScanOut O, const ExactCfg X, const SparseBufs B, uint32_t part, uint32_t nparts) {
    __shared__ Utf8Tables T;
    __shared__ uint32_t wa[8], wb[8];
    // the chunk's findings are contiguous in the output, so they are assembled in shared memory and leave as fully
    // coalesced 16-byte stores
    constexpr uint32_t kStage = 1024;
    __shared__ unsigned long long sbuf[kStage];
    // ... and so is their text (UTF-8 -> UTF-8: the bytes of the input range)
    constexpr uint32_t kTextStage = 4096;
    __shared__ __align__(16) uint8_t tbuf[kTextStage];
    const long long NE = (long long)B.ctl->ne;
    // this launch: chunks [c_lo, c_hi) of the piece
    const unsigned long long nb = (unsigned long long)((NE + kSpThreads - 1) / kSpThreads);
    const long long c_lo = (long long)((nb * part) / nparts), c_hi = part + 1 >= nparts ? (long long)nb : (long long)((nb * (part + 1)) / nparts);
    if (c_lo + (long long)blockIdx.x >= c_hi) return;
    if (B.ctl->overflow & 2u) return;  // the outputs do not fit: the host reruns with the counted sizes
    SpCtx c;
    sp_setup(P, X, B, T, c);
    for (long long e0 = (c_lo + (long long)blockIdx.x) * kSpThreads; e0 < c_hi * kSpThreads; e0 += (long long)gridDim.x * kSpThreads) {
    const long long e = e0 + threadIdx.x;
    const bool active = e < NE;
    uint32_t cr = 0, ct = 0, xr = 0, xt = 0;
    bool extra = false;
    Carry kin = carry_none(), kout = carry_none();
    const EntryHot* es = nullptr;
    if (active) {
        es = &B.H[e];
        cr = es->cnt_r; ct = es->cnt_t; xr = es->xcnt_r; xt = es->xcnt_t;
        kin = es->kin; kout = es->kout;
        extra = (e == NE - 1) && range_is_tail(X) && kout.kind == K_L && kout.k > 0;
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
    const ulonglong2 base = B.btot[e0 / kSpThreads];
    const unsigned long long br = base.x, bt = base.y;
    const bool host_out = O.host_findings != nullptr;
    const bool staged_out = host_out && tr <= kStage;
    // chunks with more findings than the staging buffer holds (text-like input) write their records to device memory first
    // and convert them kRound at a time below; their text goes through sx_materialize_kernel
    const bool round_out = host_out && !staged_out;
    const bool text_staged = tt <= kTextStage;
    if (!text_staged && threadIdx.x == 0) { B.ctl->text_fallback = 1; B.summary->text_fallback = 1; }
    if (active) {
        uint32_t ro = er, to = et;
        // the first / last record of the stream: their flags travel in the final state (host-carried text, leftover)
        auto put = [&](unsigned long long idx, const Record& r) {
            O.recs[idx] = r;
            if (staged_out) sbuf[idx - br] = wire8_of(P, r);
            if (host_out && (idx % kWirePage) == 0) O.page_base[idx / kWirePage] = r.text_off;
            if (idx == 0) O.final_state->first_flags = r.flags;
            if (text_staged) transcode_record(P, c.g, r, tbuf + (r.text_off - bt));
        };
        if (cr) {
            if (cr <= kBufRecs) {
                const Record* st = sp_staged(B, e);
                for (uint32_t k = 0; k < cr; ++k) {
                    Record r = st[k];
                    r.text_off += bt + to;
                    put(br + ro + k, r);
                }
            } else {
                // more findings than the entry's staging slots: written straight to their place by one more pass
                WinGeom wg;
                c.geo.window(list_window(X, X.cta_off, e), wg);
                WinResult r;
                if (!sp_mask_window<Dec>(P, c.ts, wg, kin, MODE_WRITE, O.recs + br + ro, bt + to, r))
                    WindowEngine<Dec>::run(P, c.ts, c.g, wg, kin, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
                for (uint32_t k = 0; k < cr; ++k) put(br + ro + k, O.recs[br + ro + k]);
            }
        }
        ro += cr; to += ct;
        if (xr) {
            if (xr <= kBufRecs) {
                const Record* st = sp_xstaged(B, e);
                for (uint32_t k = 0; k < xr; ++k) {
                    Record r = st[k];
                    r.text_off += bt + to;
                    put(br + ro + k, r);
                }
            } else {
                const WinGeom xg = ext_geom(c.geo, list_window(X, X.cta_off, e) + 1, X.pre_bytes);
                WinResult r;
                if (!sp_mask_window<Dec>(P, c.ts, xg, kout, MODE_WRITE, O.recs + br + ro, bt + to, r))
                    WindowEngine<Dec>::run(P, c.ts, c.g, xg, kout, MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
                for (uint32_t k = 0; k < xr; ++k) put(br + ro + k, O.recs[br + ro + k]);
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
            r.flags = RF_LEFTOVER | ((kout.flags & CF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u) |
                      ((kout.flags & CF_HALF) ? (uint32_t)RF_HALFSTART : 0u);
            r.precision = 0;
            put(br + ro, r);
            O.final_state->last_flags = r.flags;
        }
    }
    if (active && e == NE - 1 && range_is_tail(X)) { O.final_state->carry = kout; O.final_state->npend = es->npend; }
    __syncthreads();
    if (staged_out) {
        unsigned long long* const dst = reinterpret_cast<unsigned long long*>(O.host_findings) + br;
        for (uint32_t k = threadIdx.x; k < tr; k += kSpThreads) dst[k] = sbuf[k];
    }
    if (round_out) {
        unsigned long long* const dst = reinterpret_cast<unsigned long long*>(O.host_findings) + br;
        for (uint32_t k = threadIdx.x; k < tr; k += kSpThreads) dst[k] = wire8_of(P, O.recs[br + k]);
    }
    if (text_staged && tt) {
        // [bt, bt + tt) of the text arena: bytes up to the first 16-byte boundary, aligned 16-byte body, tail bytes
        uint8_t* const dst = B.text_out + bt;
        const uint32_t head = (uint32_t)((16u - (uint32_t)(reinterpret_cast<unsigned long long>(dst) & 15u)) & 15u);
        const uint32_t h = head < tt ? head : tt;
        const uint32_t body = (tt - h) >> 4;
        for (uint32_t k = threadIdx.x; k < h; k += kSpThreads) dst[k] = tbuf[k];
        for (uint32_t k = threadIdx.x; k < body; k += kSpThreads) {
            const uint8_t* sp = tbuf + h + 16u * k;  // shared memory is byte addressable: assemble the 16 bytes
            uint4 v;
            v.x = sp[0] | (sp[1] << 8) | (sp[2] << 16) | ((uint32_t)sp[3] << 24);
            v.y = sp[4] | (sp[5] << 8) | (sp[6] << 16) | ((uint32_t)sp[7] << 24);
            v.z = sp[8] | (sp[9] << 8) | (sp[10] << 16) | ((uint32_t)sp[11] << 24);
            v.w = sp[12] | (sp[13] << 8) | (sp[14] << 16) | ((uint32_t)sp[15] << 24);
            *reinterpret_cast<uint4*>(dst + h + 16u * k) = v;
        }
        for (uint32_t k = h + 16u * body + threadIdx.x; k < tt; k += kSpThreads) dst[k] = tbuf[k];
    }
    __syncthreads();  // sbuf / tbuf are reused by the next chunk
    }
}

// One piece through the pipeline.  st: the stream of the exact stage; side: declined heads beside the members.
// ev (optional, 7 timing events) is recorded between the stages; evs[0..1] order the side stream.
struct SparseLaunchCfg {
    unsigned grid_chunks;   // persistent grids of the per-chunk kernels (heads, ext, gather)
    unsigned grid_queue;    // ... of the queue kernels (members, declined, fix, late)
    unsigned long long rec_cap, text_cap, out_cap;
    cudaEvent_t ev_scan_prev = nullptr;  // the scan kernel of the piece before this one (another stream)
    cudaEvent_t ev_scan_done = nullptr;
    uint32_t gather_parts = 1;           // <= kGatherParts
    cudaEvent_t* ev_part = nullptr;      // recorded after every part of the gather
};
template <class Dec>
inline cudaError_t launch_sparse_impl(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B, const SparseLaunchCfg& L,
                                      cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs) {
    {
        // Same shared-memory carve-out as the prefilter (which needs the maximum): kernels that prefer different L1 /
        // shared splits cannot share an SM, and these have to run beside the prefilter CTAs of the next piece.
        static thread_local int done_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (done_dev != dev) {
            static int co_env = -2;
            if (co_env == -2) { const char* cv = getenv("SX_CARVEOUT"); co_env = cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared; }
            const int co = co_env;
            cudaFuncSetAttribute(sx_sp_heads_kernel<Dec, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_members_kernel<Dec, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_declined_kernel<Dec, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_fix_kernel<Dec>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_late_kernel<Dec>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_ext_kernel<Dec, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_gather_kernel<Dec, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_sp_snapshot_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            cudaFuncSetAttribute(sx_list_compact_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co);
            done_dev = dev;
        }
    }
    if (ev) cudaEventRecord(ev[1], st);
    {
        static int minb = -1;
        // CTAs per SM of the heads kernel, measured (profiles/r02_tuning.txt): 4: 0.214 ms, 5: 0.201, 6: 0.237, 8: 0.231
        if (minb < 0) { const char* hv = getenv("SX_HEADS_MINB"); minb = hv ? atoi(hv) : 5; }
        if (minb == 4) sx_sp_heads_kernel<Dec, 4><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
        else if (minb == 5) sx_sp_heads_kernel<Dec, 5><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
        else if (minb == 8) sx_sp_heads_kernel<Dec, 8><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
        else sx_sp_heads_kernel<Dec, 6><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
    }
    if (ev) cudaEventRecord(ev[2], st);
    sx_sp_snapshot_kernel<<<1, 1, 0, st>>>(B.ctl);
    // Mask-engine decoders: the few declined heads run on the side stream beside the members.  Byte-wise decoders: every
    // head is "declined", so they run first, in stream order, and the members find their heads resolved.
    constexpr bool kSide = MaskFamily<Dec>::kHas;
    cudaStream_t dst = kSide ? side : st;
    if (kSide) {
        cudaEventRecord(evs[0], st);
        cudaStreamWaitEvent(side, evs[0], 0);
    }
    static int mminb = -1, dminb = -1;
    // CTAs per SM, measured on 4 GiB ranges (profiles/r02_tuning.txt): koi8-r (mask engine) members 4 / 6 / 8: 56.5 / 49.2 / 46.8 ms;
    // euc-jp (byte-wise engine) members 14.8 / 14.3 / 21.3 ms, declined heads 35.1 / 30.2 / 28.1 ms
    if (mminb < 0) {
        const char* e1 = getenv("SX_MEMBERS_MINB");
        mminb = e1 ? atoi(e1) : (MaskFamily<Dec>::kHas ? 8 : 6);
        const char* e2 = getenv("SX_DECLINED_MINB");
        dminb = e2 ? atoi(e2) : 8;
    }
    const unsigned gq = L.grid_queue;
    if (dminb == 8) sx_sp_declined_kernel<Dec, 8><<<gq * 2, kSpThreads, 0, dst>>>(P, X, B);
    else if (dminb == 6) sx_sp_declined_kernel<Dec, 6><<<gq * 3 / 2, kSpThreads, 0, dst>>>(P, X, B);
    else sx_sp_declined_kernel<Dec, 4><<<gq, kSpThreads, 0, dst>>>(P, X, B);
    if (kSide) cudaEventRecord(evs[1], side);
    if (mminb == 8) sx_sp_members_kernel<Dec, 8><<<gq * 2, kSpThreads, 0, st>>>(P, X, B);
    else if (mminb == 6) sx_sp_members_kernel<Dec, 6><<<gq * 3 / 2, kSpThreads, 0, st>>>(P, X, B);
    else sx_sp_members_kernel<Dec, 4><<<gq, kSpThreads, 0, st>>>(P, X, B);
    if (ev) cudaEventRecord(ev[3], st);
    if (kSide) cudaStreamWaitEvent(st, evs[1], 0);
    sx_sp_fix_kernel<Dec><<<L.grid_queue, kSpThreads, 0, st>>>(P, X, B);
    sx_sp_late_kernel<Dec><<<L.grid_queue, kSpThreads, 0, st>>>(P, X, B);
    if (ev) cudaEventRecord(ev[4], st);
    static int xminb = -1, gminb = -1;
    // CTAs per SM, measured (profiles/r02_tuning.txt): ext 4: 0.044 ms, 6: 0.037, 8: 0.037; gather 4: 0.184 ms, 6: 0.163, 8: 0.155
    if (xminb < 0) { const char* e1 = getenv("SX_EXT_MINB"); xminb = e1 ? atoi(e1) : 6; const char* e2 = getenv("SX_GATHER_MINB"); gminb = e2 ? atoi(e2) : 8; }
    if (xminb == 8) sx_sp_ext_kernel<Dec, 8><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
    else if (xminb == 6) sx_sp_ext_kernel<Dec, 6><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
    else sx_sp_ext_kernel<Dec, 4><<<L.grid_chunks, kSpThreads, 0, st>>>(P, X, B);
    if (ev) cudaEventRecord(ev[5], st);
    if (L.ev_scan_prev) cudaStreamWaitEvent(st, L.ev_scan_prev, 0);
    sx_sp_scan_kernel<<<1, 1024, 0, st>>>(B, L.rec_cap, L.text_cap, L.out_cap, L.gather_parts);
    if (L.ev_scan_done) cudaEventRecord(L.ev_scan_done, st);
    for (uint32_t part = 0; part < L.gather_parts; ++part) {
        const unsigned gg = L.grid_chunks / L.gather_parts + 1;
        if (gminb == 8) sx_sp_gather_kernel<Dec, 8><<<gg, kSpThreads, 0, st>>>(P, O, X, B, part, L.gather_parts);
        else if (gminb == 6) sx_sp_gather_kernel<Dec, 6><<<gg, kSpThreads, 0, st>>>(P, O, X, B, part, L.gather_parts);
        else sx_sp_gather_kernel<Dec, 4><<<gg, kSpThreads, 0, st>>>(P, O, X, B, part, L.gather_parts);
        if (L.ev_part) cudaEventRecord(L.ev_part[part], st);
    }
    if (ev) cudaEventRecord(ev[6], st);
    return cudaGetLastError();
}
constexpr uint32_t kSparseLaunches = 9;  // heads, snapshot, declined, members, fix, late, ext, scan, gather

}  // namespace sx
