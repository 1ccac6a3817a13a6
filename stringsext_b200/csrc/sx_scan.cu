// sx_scan.cu -- sm_100a kernels of the scanner hot path + the C ABI (include/stringsext_b200.h).
//
// sx_scan_kernel<Dec>: persistent CTAs, each owning a contiguous range of tiles of the input
// stream (a tile = 8 reference slices of 4096 B at the default geometry, input.rs:22).  Per tile:
//   stage 0  the tile is staged from HBM into shared memory in a 128-byte-swizzled layout (lane l
//            reading 16-byte chunk c of window l is then bank-conflict free),
//   stage A  one lane per window: decoder + SplitStr automaton under the null carry
//            -> transfer-function descriptor (WinDesc, sx_core.cuh),
//   stage B  carries are resolved across the tile (constant / accumulate / replay),
//   stage C  windows that emit are counted, a block scan + one atomicAdd per tile reserves record
//            and text space, and the records are written in stream order inside the tile.
// sx_materialize_kernel: transcodes each record's input range to UTF-8 text.
//
// There is no CPU scanning path in this file: without a CUDA device every call fails.
#include "../../include/stringsext_b200.h"
#include "sx_core.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace sx {

constexpr int kThreads = 256;
constexpr int kTileBytes = 32768;
constexpr int kMaxWin = 512;  // windows per tile
constexpr int kMaxWpt = kMaxWin / kThreads;

struct FinalState {
    Carry carry;
    int32_t npend;
    uint32_t overflow;
};

struct ScanOut {
    Record* recs;
    unsigned long long rec_cap;
    unsigned long long text_cap;
    uint2* tile_desc;              // per tile {first record, record count}
    unsigned long long* counters;  // [0] records, [1] text bytes
    FinalState* final_state;
};

// A tile is a run of `nwin_tile` consecutive windows (<= kTileBytes of contiguous input; exactly
// 8 slices of 4096 B at the default geometry).
struct TileCfg {
    long long ntiles;
    long long total_windows;
    uint32_t nwin_tile, wpt;
};

struct SmemLayout {
    uint8_t* data;               // kTileBytes
    WinDesc* desc;               // kMaxWin
    Carry* kin;                  // kMaxWin + 1
    uint8_t* done;               // kMaxWin + 1
    uint32_t* warp_a;            // 8
    uint32_t* warp_b;            // 8
    unsigned long long* bases;   // 2
    int32_t* misc;               // [0] npend at the end of the tile's last window
};
constexpr size_t kSmemBytes =
    kTileBytes + sizeof(WinDesc) * kMaxWin + sizeof(Carry) * (kMaxWin + 4) + (kMaxWin + 16) + 64 + 64 + 16 + 16;

__device__ __forceinline__ SmemLayout carve(uint8_t* p) {
    SmemLayout S;
    S.data = p; p += kTileBytes;
    S.desc = reinterpret_cast<WinDesc*>(p); p += sizeof(WinDesc) * kMaxWin;
    S.kin = reinterpret_cast<Carry*>(p); p += sizeof(Carry) * (kMaxWin + 4);
    S.warp_a = reinterpret_cast<uint32_t*>(p); p += 32;
    S.warp_b = reinterpret_cast<uint32_t*>(p); p += 32;
    S.bases = reinterpret_cast<unsigned long long*>(p); p += 16;
    S.misc = reinterpret_cast<int32_t*>(p); p += 16;
    S.done = p;
    return S;
}

__device__ __forceinline__ uint32_t swz(uint32_t r) { return r ^ (((r >> 7) & 7u) << 4); }

struct SmemTile {
    const uint8_t* sm;
    int64_t lo, hi;
    GlobalSrc g;
    __device__ __forceinline__ uint8_t get(int64_t off) const {
        if (off >= lo && off < hi) return sm[swz((uint32_t)(off - lo))];
        return g.get(off);
    }
    template <class F>
    __device__ __forceinline__ void for_each_byte(int64_t ws, int64_t we, F&& f) const {
        uint32_t r = (uint32_t)(ws - lo);
        const uint32_t rend = (uint32_t)(we - lo);
        while (r < rend) {
            const uint32_t r16 = r & ~15u;
            const uint4 v = *reinterpret_cast<const uint4*>(sm + swz(r16));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t i0 = r - r16;
            const uint32_t i1 = (rend - r16) < 16u ? (rend - r16) : 16u;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((uint32_t)i >= i0 && (uint32_t)i < i1) f((w[i >> 2] >> ((i & 3) * 8)) & 0xFFu, lo + (int64_t)(r16 + i));
            }
            r = r16 + 16;
        }
    }
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Tile loader: coalesced 16-byte streaming loads, swizzled 16-byte shared stores.
__device__ __forceinline__ void load_tile(uint8_t* sm, const ScanParams& P, int64_t lo, int64_t hi) {
    const uint32_t nbytes = (uint32_t)(hi - lo);
    const uint32_t nchunks = (nbytes + 15u) >> 4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(P.in) + (uintptr_t)lo) & 15u) == 0;
    for (uint32_t c = threadIdx.x; c < nchunks; c += kThreads) {
        uint4 v;
        if (aligned && (c + 1) * 16u <= nbytes) {
            v = ldg_stream(reinterpret_cast<const uint4*>(P.in + lo) + c);
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < 16; ++i) {
                const uint32_t o = c * 16u + i;
                if (o < nbytes) w[i >> 2] |= (uint32_t)P.in[lo + o] << ((i & 3) * 8);
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(sm + swz(c * 16u)) = v;
    }
}

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

template <class Dec>
__device__ Carry tile_pass(const ScanParams& P, const ScanOut& O, const TileCfg& cfg, const Geometry& geo,
                           const SmemLayout& S, long long tile, bool full, Carry carry_in) {
    const uint32_t tid = threadIdx.x;
    const GlobalSrc g{P.in, P.pend};
    const int64_t w0 = (int64_t)tile * cfg.nwin_tile;
    const uint32_t nvalid = (uint32_t)((cfg.total_windows - w0) < (long long)cfg.nwin_tile ? (cfg.total_windows - w0) : (long long)cfg.nwin_tile);
    int64_t lo, hi;
    {
        WinGeom wa, wb;
        geo.window(w0, wa);
        geo.window(w0 + nvalid - 1, wb);
        lo = wa.ws;
        hi = wb.we;
    }

    load_tile(S.data, P, lo, hi);
    __syncthreads();

    const SmemTile ts{S.data, lo, hi, g};
    const bool last_tile = (tile == cfg.ntiles - 1);

    // ---- stage A: per-window summary under the null carry -------------------------------------
    for (uint32_t k = 0; k < cfg.wpt; ++k) {
        const uint32_t i = tid * cfg.wpt + k;
        if (i < nvalid) {
            WinGeom wg;
            geo.window(w0 + i, wg);
            WinResult r;
            WinDesc d;
            scan_window<Dec>(P, ts, g, wg, carry_none(), MODE_COUNT, nullptr, 0, r, &d);
            S.desc[i] = d;
            if (d.type == WT_CONST) { S.kin[i + 1] = d.null_out; S.done[i + 1] = 1; }
            else S.done[i + 1] = 0;
            if (i == nvalid - 1) S.misc[0] = r.npend_out;
        }
    }
    if (tid == 0) { S.kin[0] = carry_in; S.done[0] = 1; }
    __syncthreads();

    // ---- stage B: resolve the carries ----------------------------------------------------------
    for (;;) {
        uint32_t rdy = 0;
        for (uint32_t k = 0; k < cfg.wpt; ++k) {
            const uint32_t i = tid * cfg.wpt + k;
            if (i < nvalid && !S.done[i + 1] && S.done[i]) rdy |= 1u << k;
        }
        __syncthreads();
        for (uint32_t k = 0; k < cfg.wpt; ++k) {
            if (!((rdy >> k) & 1u)) continue;
            const uint32_t i = tid * cfg.wpt + k;
            const Carry kin = S.kin[i];
            const WinDesc d = S.desc[i];
            WinGeom wg;
            geo.window(w0 + i, wg);
            Carry out;
            if (kin.kind == K_UNKNOWN) out = kin;
            else if (d.type == WT_CASEB) out = eval_caseb(P, d, kin, (uint32_t)(wg.we - wg.ws));
            else {
                WinResult r;
                scan_window<Dec>(P, ts, g, wg, kin, MODE_STATE, nullptr, 0, r, nullptr);
                out = r.out;
            }
            S.kin[i + 1] = out;
            S.done[i + 1] = 1;
        }
        if (!__syncthreads_or(rdy != 0)) break;
    }
    const Carry carry_out = S.kin[nvalid];
    if (!full) {
        __syncthreads();
        return carry_out;
    }

    // ---- stage C: count, reserve, write --------------------------------------------------------
    uint32_t cr[kMaxWpt], ct[kMaxWpt];
    uint32_t emit_mask = 0;
    uint32_t sum_r = 0, sum_t = 0;
    for (uint32_t k = 0; k < cfg.wpt; ++k) {
        cr[k] = 0; ct[k] = 0;
        const uint32_t i = tid * cfg.wpt + k;
        if (i >= nvalid) continue;
        const Carry kin = S.kin[i];
        const WinDesc d = S.desc[i];
        if (needs_emit(P, d, kin)) {
            emit_mask |= 1u << k;
            if (carry_is_null(kin) && d.nrec != 0xFFFFu) { cr[k] = d.nrec; ct[k] = d.ntext; }
            else {
                WinGeom wg;
                geo.window(w0 + i, wg);
                WinResult r;
                scan_window<Dec>(P, ts, g, wg, kin, MODE_COUNT, nullptr, 0, r, nullptr);
                cr[k] = r.nrec; ct[k] = r.ntext;
            }
        }
        sum_r += cr[k]; sum_t += ct[k];
    }
    // the scanner's final leftover travels as one extra pseudo record at the very end of the stream
    const bool owns_last = last_tile && nvalid > 0 && ((nvalid - 1) / cfg.wpt == tid);
    const bool extra = owns_last && carry_out.kind == K_L && carry_out.k > 0;
    if (extra) { sum_r += 1; sum_t += carry_out.out_bytes; }

    uint32_t er, et, tr, tt;
    block_excl_scan2(sum_r, sum_t, S.warp_a, S.warp_b, er, et, tr, tt);
    if (tid == 0) {
        unsigned long long br = 0, bt = 0;
        if (tr) br = atomicAdd(&O.counters[0], (unsigned long long)tr);
        if (tt) bt = atomicAdd(&O.counters[1], (unsigned long long)tt);
        S.bases[0] = br;
        S.bases[1] = bt;
        O.tile_desc[tile] = make_uint2((uint32_t)br, tr);
        if (br + tr > O.rec_cap || bt + tt > O.text_cap) O.final_state->overflow = 1;
        if (last_tile) { O.final_state->carry = carry_out; O.final_state->npend = S.misc[0]; }
    }
    __syncthreads();
    const unsigned long long br = S.bases[0], bt = S.bases[1];
    const bool fits = (br + tr <= O.rec_cap) && (bt + tt <= O.text_cap);
    if (fits) {
        uint32_t ro = er, to = et;
        for (uint32_t k = 0; k < cfg.wpt; ++k) {
            if ((emit_mask >> k) & 1u) {
                const uint32_t i = tid * cfg.wpt + k;
                WinGeom wg;
                geo.window(w0 + i, wg);
                WinResult r;
                scan_window<Dec>(P, ts, g, wg, S.kin[i], MODE_WRITE, O.recs + br + ro, bt + to, r, nullptr);
            }
            ro += cr[k]; to += ct[k];
        }
        if (extra) {
            Record r;
            r.position = 0;
            r.in_start = P.len - (int64_t)carry_out.in_bytes;
            r.in_len = carry_out.in_bytes - (uint32_t)S.misc[0];
            r.text_len = carry_out.out_bytes;
            r.text_off = bt + to;
            r.flags = RF_LEFTOVER | ((carry_out.flags & CF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u);
            r.precision = 0;
            O.recs[br + ro] = r;
        }
    }
    __syncthreads();
    return carry_out;
}

template <class Dec>
__global__ void __launch_bounds__(kThreads, 2)
sx_scan_kernel(const __grid_constant__ ScanParams P, const ScanOut O, const TileCfg cfg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemLayout S = carve(smem_raw);
    Geometry geo;
    geo.init(P);
    const long long per = (cfg.ntiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per;
    const long long t1 = (t0 + per) < cfg.ntiles ? (t0 + per) : cfg.ntiles;
    if (t0 >= t1) return;
    Carry c;
    if (t0 == 0) c = P.k0;
    else {
        // Warm-up: find the carry at the range start by replaying preceding tiles in state-only mode
        // from an unknown carry; almost always one tile suffices (it ends in a constant window).
        long long back = 1;
        for (;;) {
            long long ts_ = t0 - back;
            if (ts_ <= 0) { ts_ = 0; c = P.k0; }
            else c = carry_unknown();
            for (long long t = ts_; t < t0; ++t) c = tile_pass<Dec>(P, O, cfg, geo, S, t, false, c);
            if (c.kind != K_UNKNOWN) break;
            back *= 2;
        }
    }
    for (long long t = t0; t < t1; ++t) c = tile_pass<Dec>(P, O, cfg, geo, S, t, true, c);
}

__global__ void __launch_bounds__(256)
sx_materialize_kernel(const __grid_constant__ ScanParams P, const Record* __restrict__ recs, unsigned long long n,
                      uint8_t* __restrict__ text, unsigned long long text_cap) {
    const GlobalSrc g{P.in, P.pend};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const Record r = recs[i];
        if (r.text_off + r.text_len <= text_cap) transcode_range(P, g, r.in_start, r.in_len, text + r.text_off);
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
    uint2* d_tile = nullptr; size_t tile_cap = 0;
    unsigned long long* d_counters = nullptr;
    FinalState* d_final = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int num_sms = 0;
    double rec_per_byte = 1.0 / 1024, text_per_byte = 1.0 / 64;
    sx_scan_stats stats;
};

struct sx_finding_collection {
    std::vector<sx_finding> v;
    std::vector<uint8_t> text;
    uint64_t first_byte_position = 0;
    int str_buf_overflow = 0;
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
    if (m->grep_char >= 0) { set_err(SX_ERR_UNSUPPORTED, "grep_char is not implemented by the CUDA scanner yet"); return fail; }
    if (m->require_same_unicode_block) { set_err(SX_ERR_UNSUPPORTED, "require_same_unicode_block is not implemented by the CUDA scanner yet"); return fail; }
    if (m->output_line_char_nb_max < 6 || m->output_line_char_nb_max > 8192) { set_err(SX_ERR_UNSUPPORTED, "output_line_char_nb_max must be in 6..8192"); return fail; }
    if (m->chars_min_nb == 0 || m->chars_min_nb > m->output_line_char_nb_max) { set_err(SX_ERR_UNSUPPORTED, "chars_min_nb must be in 1..output_line_char_nb_max"); return fail; }
    if (m->encoding_id > SX_ENC_UTF_32BE) { set_err(SX_ERR_ARGUMENT, "unknown encoding_id"); return fail; }
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
    bool ok = cuda_ok(cudaMalloc(&ss->d_counters, 2 * sizeof(unsigned long long)), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&ss->d_final, sizeof(FinalState)), "cudaMalloc");
    for (int i = 0; ok && i < 4; ++i) ok = cuda_ok(cudaEventCreate(&ss->ev[i]), "cudaEventCreate");
    if (!ok) { sx_scanner_state_free(ss); return fail; }
    return ss;
}

void sx_scanner_state_free(sx_scanner_state* ss) {
    if (!ss) return;
    cudaSetDevice(ss->device);
    cudaFree(ss->d_in); cudaFree(ss->d_recs); cudaFree(ss->d_text); cudaFree(ss->d_tile);
    cudaFree(ss->d_counters); cudaFree(ss->d_final);
    for (auto e : ss->ev) if (e) cudaEventDestroy(e);
    delete ss;
}

uint64_t sx_scanner_state_consumed_bytes(const sx_scanner_state* ss) { return ss->consumed; }
int sx_scanner_state_maybe_cut(const sx_scanner_state* ss) { return ss->cut ? 1 : 0; }
size_t sx_scanner_state_leftover(const sx_scanner_state* ss, const uint8_t** p) {
    *p = ss->leftover.data();
    return ss->leftover.size();
}
void sx_scanner_state_last_stats(const sx_scanner_state* ss, sx_scan_stats* out) { *out = ss->stats; }

size_t sx_fc_len(const sx_finding_collection* fc) { return fc->v.size(); }
const sx_finding* sx_fc_get(const sx_finding_collection* fc, size_t i) { return &fc->v[i]; }
const sx_finding* sx_fc_data(const sx_finding_collection* fc) { return fc->v.data(); }
uint64_t sx_fc_first_byte_position(const sx_finding_collection* fc) { return fc->first_byte_position; }
int sx_fc_str_buf_overflow(const sx_finding_collection* fc) { return fc->str_buf_overflow; }
void sx_fc_free(sx_finding_collection* fc) { delete fc; }

}  // extern "C"

template <class T>
static bool grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return true;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 256;
    if (!cuda_ok(cudaMalloc(p, want * sizeof(T)), "cudaMalloc")) return false;
    *cap = want;
    return true;
}

template <class Dec>
static cudaError_t launch_scan(const ScanParams& P, const ScanOut& O, const TileCfg& cfg, int grid, cudaStream_t st) {
    static thread_local bool attr_done[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sx_scan_kernel<Dec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    sx_scan_kernel<Dec><<<grid, kThreads, kSmemBytes, st>>>(P, O, cfg);
    return cudaGetLastError();
}

static cudaError_t launch_scan_enc(const ScanParams& P, const ScanOut& O, const TileCfg& cfg, int grid, cudaStream_t st) {
    switch (P.enc) {
    case ENC_XUD: return launch_scan<DecXud>(P, O, cfg, grid, st);
    case ENC_UTF8: return launch_scan<DecUtf8>(P, O, cfg, grid, st);
    case ENC_UTF16LE: return launch_scan<DecUtf16<false>>(P, O, cfg, grid, st);
    case ENC_UTF16BE: return launch_scan<DecUtf16<true>>(P, O, cfg, grid, st);
    case ENC_SB: return launch_scan<DecSb>(P, O, cfg, grid, st);
    case ENC_UTF32LE: return launch_scan<DecUtf32<false>>(P, O, cfg, grid, st);
    case ENC_UTF32BE: return launch_scan<DecUtf32<true>>(P, O, cfg, grid, st);
    }
    return cudaErrorInvalidValue;
}

static size_t utf8_char_count(const std::vector<uint8_t>& s) {
    size_t n = 0;
    for (uint8_t b : s) n += (b & 0xC0) != 0x80;
    return n;
}

extern "C" sx_finding_collection* sx_scan_stream(sx_scanner_state* ss, int input_file_id, const void* buf, size_t len,
                                                 size_t slice_len, int buf_is_device, int is_last, void* cuda_stream) {
    sx_finding_collection* const fail = nullptr;
    if (!ss) { set_err(SX_ERR_ARGUMENT, "state is NULL"); return fail; }
    if (len > 0 && !buf) { set_err(SX_ERR_ARGUMENT, "buf is NULL"); return fail; }
    const uint32_t q = ss->m.output_line_char_nb_max;
    const uint32_t W = 2 * q;
    if (slice_len == 0 || slice_len > 0x7FFFFFFFull) { set_err(SX_ERR_ARGUMENT, "slice_len must be in 1..2^31-1"); return fail; }
    const uint32_t wps = (uint32_t)((slice_len + W - 1) / W);
    memset(&ss->stats, 0, sizeof ss->stats);
    sx_finding_collection* fc = new sx_finding_collection();
    fc->first_byte_position = ss->consumed;
    if (len == 0) return fc;  // finding_collection.rs:124: the window loop does not run, state untouched
    struct Guard { sx_finding_collection* p; ~Guard() { delete p; } } guard{fc};

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
    ScanParams P;
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
    else if (!ss->leftover.empty()) P.k0 = Carry{K_L, CF_HOSTCARRY, (uint16_t)utf8_char_count(ss->leftover), (uint32_t)ss->npend, 0};
    else P.k0 = carry_none();
    memcpy(P.sb_table, ss->m.sb_table, sizeof P.sb_table);

    TileCfg cfg;
    {
        const long long full = (long long)(len / slice_len);
        const size_t rest = len - (size_t)full * slice_len;
        cfg.total_windows = full * wps + (long long)((rest + W - 1) / W);
    }
    cfg.nwin_tile = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)kMaxWin, (uint32_t)kTileBytes / W));
    cfg.wpt = (cfg.nwin_tile + kThreads - 1) / kThreads;
    cfg.ntiles = (cfg.total_windows + cfg.nwin_tile - 1) / cfg.nwin_tile;
    if (cfg.ntiles > 0xFFFFFFFFLL) { set_err(SX_ERR_UNSUPPORTED, "stream too long for one call"); return fail; }
    const int grid = (int)std::min<long long>(cfg.ntiles, (long long)ss->num_sms * 2);

    if (!grow(&ss->d_tile, &ss->tile_cap, (size_t)cfg.ntiles)) return fail;
    size_t need_recs = (size_t)(len * ss->rec_per_byte) + 4096;
    size_t need_text = (size_t)(len * ss->text_per_byte) + 65536;
    unsigned long long counters[2] = {0, 0};
    FinalState fin;
    for (int attempt = 0;; ++attempt) {
        if (!grow(&ss->d_recs, &ss->rec_cap, need_recs)) return fail;
        if (!grow(&ss->d_text, &ss->text_cap, need_text)) return fail;
        CK(cudaMemsetAsync(ss->d_counters, 0, 2 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(ss->d_final, 0, sizeof(FinalState), st));
        ScanOut O{ss->d_recs, ss->rec_cap, ss->text_cap, ss->d_tile, ss->d_counters, ss->d_final};
        CK(cudaEventRecord(ss->ev[0], st));
        CK(launch_scan_enc(P, O, cfg, grid, st));
        CK(cudaEventRecord(ss->ev[1], st));
        ss->stats.kernel_launches++;
        CK(cudaMemcpyAsync(counters, ss->d_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&fin, ss->d_final, sizeof fin, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ss->stats.d2h_bytes += sizeof counters + sizeof fin;
        if (!fin.overflow && counters[0] <= ss->rec_cap && counters[1] <= ss->text_cap) break;
        if (attempt >= 2) { set_err(SX_ERR_CUDA, "output buffers still too small after regrowing"); return fail; }
        need_recs = (size_t)counters[0] + 1024;
        need_text = (size_t)counters[1] + 4096;
        ss->stats.relaunches++;
    }
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, ss->ev[0], ss->ev[1]);
        ss->stats.scan_kernel_ms = ms;
    }
    const size_t nrec = (size_t)counters[0];
    const size_t ntext = (size_t)counters[1];
    ss->rec_per_byte = std::max(1.0 / 4096, 1.3 * (double)nrec / (double)len);
    ss->text_per_byte = std::max(1.0 / 256, 1.3 * (double)ntext / (double)len);
    ss->stats.n_records = nrec;
    ss->stats.text_bytes = ntext;

    // ---- text + download ---------------------------------------------------------------------------
    std::vector<Record> recs(nrec);
    std::vector<uint8_t> text(ntext);
    std::vector<uint2> tiles((size_t)cfg.ntiles);
    if (nrec) {
        const int mgrid = (int)std::min<size_t>((nrec + 255) / 256, (size_t)ss->num_sms * 8);
        CK(cudaEventRecord(ss->ev[2], st));
        sx_materialize_kernel<<<mgrid, 256, 0, st>>>(P, ss->d_recs, nrec, ss->d_text, ss->text_cap);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ss->ev[3], st));
        ss->stats.kernel_launches++;
        CK(cudaMemcpyAsync(recs.data(), ss->d_recs, nrec * sizeof(Record), cudaMemcpyDeviceToHost, st));
        if (ntext) CK(cudaMemcpyAsync(text.data(), ss->d_text, ntext, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(tiles.data(), ss->d_tile, tiles.size() * sizeof(uint2), cudaMemcpyDeviceToHost, st));
        ss->stats.d2h_bytes += nrec * sizeof(Record) + ntext + tiles.size() * sizeof(uint2);
    }
    // bytes that stay inside the decoder: the last npend bytes of (old pend ++ buffer)
    uint8_t tail[8] = {0};
    const size_t tail_n = std::min<size_t>(8, len);
    if (buf_is_device) CK(cudaMemcpyAsync(tail + 8 - tail_n, d_in + len - tail_n, tail_n, cudaMemcpyDeviceToHost, st));
    else memcpy(tail + 8 - tail_n, (const uint8_t*)buf + len - tail_n, tail_n);
    CK(cudaStreamSynchronize(st));
    if (nrec) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ss->ev[2], ss->ev[3]);
        ss->stats.materialize_kernel_ms = ms;
    }

    // ---- build the collection in stream order (tiles are contiguous record blocks) -------------------
    size_t extra_text = 0;
    for (const Record& r : recs) if (r.flags & RF_HOSTCARRY) extra_text += ss->leftover.size() + r.text_len;
    fc->text.resize(ntext + extra_text + 1);
    if (ntext) memcpy(fc->text.data(), text.data(), ntext);
    size_t extra_off = ntext;
    fc->v.reserve(nrec);
    std::vector<uint8_t> new_leftover;
    bool have_leftover = false;
    for (size_t t = 0; t < tiles.size(); ++t) {
        const size_t b = tiles[t].x, c = tiles[t].y;
        for (size_t i = b; i < b + c && i < nrec; ++i) {
            const Record& r = recs[i];
            const uint8_t* s = fc->text.data() + r.text_off;
            size_t s_len = r.text_len;
            if (r.flags & RF_HOSTCARRY) {  // prepend the text the previous call left in the ScannerState
                uint8_t* d = fc->text.data() + extra_off;
                memcpy(d, ss->leftover.data(), ss->leftover.size());
                memcpy(d + ss->leftover.size(), s, r.text_len);
                s = d;
                s_len = ss->leftover.size() + r.text_len;
                extra_off += s_len;
            }
            if (r.flags & RF_LEFTOVER) {
                new_leftover.assign(s, s + s_len);
                have_leftover = true;
                continue;
            }
            sx_finding f;
            f.position = r.position;
            f.precision = (uint8_t)r.precision;
            f.completes_previous = (r.flags & RF_COMPLETES) ? 1 : 0;
            f.input_file_id = (int16_t)input_file_id;
            f.mission_id = ss->m.mission_id;
            f.s = s;
            f.s_len = (uint32_t)s_len;
            f.in_start = r.in_start;
            f.in_len = r.in_len;
            fc->v.push_back(f);
        }
    }

    // ---- ScannerState update (finding_collection.rs:330-338) -----------------------------------------
    ss->cut = fin.carry.kind == K_C;
    if (fin.carry.kind == K_L && fin.carry.k > 0 && have_leftover) ss->leftover.swap(new_leftover);
    else ss->leftover.clear();
    {
        uint8_t all[16];
        memcpy(all, ss->pend, 8);  // old pend occupies all[8-npend..8)
        memcpy(all + 8, tail, 8);  // tail occupies all[16-tail_n..16)
        // concatenation old_pend ++ tail, right aligned: when len < 8 the old pend bytes must follow on directly
        uint8_t cat[16];
        size_t cn = 0;
        for (int i = 8 - ss->npend; i < 8; ++i) cat[cn++] = all[i];
        for (size_t i = 8 - tail_n; i < 8; ++i) cat[cn++] = tail[i];
        const int np = fin.npend;
        memset(ss->pend, 0, 8);
        if (np > 0 && (size_t)np <= cn) memcpy(ss->pend + 8 - np, cat + cn - np, (size_t)np);
        ss->npend = (np > 0 && (size_t)np <= cn) ? np : 0;
    }
    ss->consumed += len;
    guard.p = nullptr;
    return fc;
}

extern "C" sx_finding_collection* sx_finding_collection_from(sx_scanner_state* ss, int input_file_id, const uint8_t* buf,
                                                             size_t len, int is_last) {
    return sx_scan_stream(ss, input_file_id, buf, len, len ? len : 1, 0, is_last, nullptr);
}

extern "C" size_t sx_merge(const sx_finding_collection* const* fcs, size_t n, const sx_finding** out) {
    // finding.rs:92-109 via itertools::kmerge (main.rs:133): position, then mission_id; every
    // collection is position-monotone, so a stable sort of the concatenation is the k-way merge.
    size_t k = 0;
    for (size_t i = 0; i < n; ++i)
        for (const sx_finding& f : fcs[i]->v) out[k++] = &f;
    std::stable_sort(out, out + k, [](const sx_finding* a, const sx_finding* b) {
        if (a->position != b->position) return a->position < b->position;
        return a->mission_id < b->mission_id;
    });
    return k;
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
