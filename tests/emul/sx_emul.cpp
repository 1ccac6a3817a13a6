// sx_emul.cpp -- TEST-ONLY host harness: runs the product's per-window automaton
// (stringsext_b200/csrc/sx_core.cuh) sequentially on the CPU so that the window decomposition,
// transfer-function classification and emit rules can be differential-tested against the oracle
// without a GPU.  It is NOT part of the product library and nothing in stringsext_b200/ loads it.
#include <stdint.h>
#ifndef __CUDACC__
struct uint4 { uint32_t x, y, z, w; };
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return s ? (uint32_t)((((uint64_t)hi << 32) | lo) >> s) : lo; }
#endif
#include "../../stringsext_b200/csrc/sx_mask_utf8.cuh"
#include "../../stringsext_b200/csrc/sx_mb_tables.inc"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>


using namespace sx;

static Utf8Tables g_tables;
static int g_use_fast = 1;
static int g_use_mask = 1;
static uint64_t g_mask_ok = 0, g_mask_declined = 0;
static uint64_t g_mask_mismatch = 0;
static uint64_t g_head_ok = 0;
static uint64_t g_guard_behind = 0, g_guard_behind_hard = 0;
static uint64_t g_guard_ok = 0, g_guard_killer = 0;  // WT_GUARD windows: carry-in benign / a leftover that fills the first run
static uint64_t g_indep = 0, g_dep = 0;
struct HostTile {
    GlobalSrc g;
    const Utf8Tables* tables() const { return g_use_fast ? &g_tables : nullptr; }
    uint32_t lut(uint32_t i) const { return g_tables.tt[i]; }
    uint32_t cls(uint32_t b) const { return g_tables.cls[b]; }
    bool use_mask() const { return g_use_mask != 0; }
    void mask_result(bool ok) const { if (ok) g_mask_ok++; else g_mask_declined++; }
    void mask_mismatch(int64_t ws, int64_t we, const Carry& kin, int mode) const {
        if (g_mask_mismatch++ < 5)
            fprintf(stderr, "mask engine mismatch: window [%lld,%lld) kin kind=%d k=%d in_bytes=%u flags=%d mode=%d\n", (long long)ws,
                    (long long)we, kin.kind, kin.k, kin.in_bytes, kin.flags, mode);
    }
    uint4 load_chunk(int64_t r16, int64_t ws, int64_t we) const {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16; ++i) {
            const int64_t o = r16 + i;
            if (o >= ws && o < we) w[i >> 2] |= (uint32_t)g.get(o) << ((i & 3) * 8);
        }
        uint4 v; v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
        return v;
    }
    uint8_t get(int64_t off) const { return g.get(off); }
    template <class F> void for_each_byte(int64_t ws, int64_t we, F&& f) const {
        for (int64_t p = ws; p < we; ++p) f((uint32_t)g.get(p), p);
    }
};

// Mirrors the device pipeline: [prefilter -> E list] -> exact processing of the listed windows
// (adjacent entries chain their carries, a non-adjacent entry starts from the null carry, a
// window whose carry-out is "cut" extends into an unlisted successor).
template <class Dec>
static void run(const ScanParams& P, int use_pref, std::vector<Record>& recs, std::vector<uint8_t>& text,
                Carry* final_carry, int32_t* final_npend, uint64_t* stats, std::vector<uint32_t>& list) {
    GlobalSrc g{P.in, P.pend};
    HostTile ts{g};
    Geometry geo;
    geo.init(P);
    const int64_t full = P.len / P.slice_len;
    const int64_t rest = P.len - full * (int64_t)P.slice_len;
    const int64_t total = full * geo.wps + (rest + P.W - 1) / P.W;
    PrefCfg pc = make_pref_cfg(P, true, P.mb_a, P.mb_b);
    // general missions: every window goes through the exact stage (an unlisted window's carry-out would depend on its
    // carry-in: killed segments, stale lead bytes -- sx_scan.cu), except --grep-char alone (PrefCfg::kill_trail)
    if (!use_pref || (P.general && !pref_general_ok(P))) pc.enabled = 0;
    stats[7] = pc.enabled;
    list.clear();
    if (pc.enabled) {
        for (int64_t w = 0; w < total; ++w)
            if (pref_interesting_ref(P, pc, geo, ts, w, total)) list.push_back((uint32_t)w);
    } else {
        for (int64_t w = 0; w < total; ++w) list.push_back((uint32_t)w);
    }
    const size_t ne = list.size();
    Carry kprev = carry_none();
    int32_t last_npend = P.npend;
    auto emit_window = [&](const WinGeom& wg, const Carry& kin, WinResult& rstate) {
        WinResult rc;
        WindowEngine<Dec>::run(P, ts, g, wg, kin, MODE_COUNT, nullptr, 0, rc, nullptr);
        rstate = rc;
        if (rc.nrec == 0) return;
        const size_t base = recs.size();
        recs.resize(base + rc.nrec);
        WinResult rw;
        WindowEngine<Dec>::run(P, ts, g, wg, kin, MODE_WRITE, recs.data() + base, text.size(), rw, nullptr);
        const size_t tb = text.size();
        text.resize(tb + rc.ntext + 8);
        for (size_t k = base; k < recs.size(); ++k) {
            Record& r = recs[k];
            const uint32_t n = transcode_record(P, g, r, text.data() + r.text_off);
            if (n != r.text_len) stats[6]++;
        }
        text.resize(tb + rc.ntext);
    };
    WinDesc prev_d;
    bool prev_valid = false, prev_off_null = false;
    for (size_t e = 0; e < ne; ++e) {
        const int64_t w = list[e];
        WinGeom wg;
        geo.window(w, wg);
        const bool adjacent = e > 0 && (int64_t)list[e - 1] == w - 1;
        const uint32_t pre_bytes = pc.pre_bytes;
        Carry kin = adjacent ? kprev : P.k0;
        if (!adjacent && w != 0) {
            const WinGeom rg = preroll_geom(geo, w, pre_bytes);
            WinResult rr;
            WindowEngine<Dec>::run(P, ts, g, rg, carry_none(), MODE_STATE, nullptr, 0, rr, nullptr);
            kin = rr.out;
        }
        // summary under the null carry + classification self-check
        WinResult r0;
        WinDesc d;
        WindowEngine<Dec>::run(P, ts, g, wg, carry_none(), MODE_COUNT, nullptr, 0, r0, &d);
        if (P.general && d.type == WT_DEP)  // grep_char / same block / n > q: constant iff the flag handed to segment 2 does not matter
            d.type = classify_general(wg, r0, [&](const WinGeom& g2) {
                WinResult r2;
                WindowEngine<Dec>::run(P, ts, g, g2, carry_none(), MODE_STATE, nullptr, 0, r2, nullptr);
                return r2.out;
            });
        WinResult rs;
        WindowEngine<Dec>::run(P, ts, g, wg, kin, MODE_STATE, nullptr, 0, rs, nullptr);
        Carry kout;
        if (d.type == WT_CONST) kout = d.null_out;
        else if (d.type == WT_CASEB) { kout = eval_caseb(P, d, kin, (uint32_t)(wg.we - wg.ws)); stats[2]++; }
        else if (d.type == WT_GUARD && guard_benign(P, d, kin)) { kout = d.null_out; g_guard_ok++; }
        else { kout = rs.out; stats[1]++; if (d.type == WT_GUARD) g_guard_killer++; }
        stats[0]++;
        if (memcmp(&rs.out, &kout, sizeof(Carry)) != 0) { stats[3]++; kout = rs.out; }
        // the block kernel's warm-up rule (carry-in unknown): checked against the replay under the real carry
        if (P.general && adjacent && prev_valid && guard_known_behind(P, d, prev_d)) {
            g_guard_behind++;
            if (prev_off_null) g_guard_behind_hard++;  // the predecessor's carry-out was not its null_out (killed / cut)
            if (memcmp(&rs.out, &d.null_out, sizeof(Carry)) != 0) stats[3]++;
        }
        prev_d = d;
        prev_valid = true;
        prev_off_null = memcmp(&kout, &d.null_out, sizeof(Carry)) != 0;
        WinResult rc;
        emit_window(wg, kin, rc);
        if (!adjacent && w != 0 && MaskFamily<Dec>::kHas && !P.general && g_use_mask && g_use_fast) {
            // the product resolves heads in one pass (carry-in derived from the 32 bytes before the window): same
            // carry-in as the pre-roll, same counts and carry-out as the pass under that carry
            WinResult rh;
            if (mask_head<MaskFamily<Dec>::kSByte>(P, ts, wg, pre_bytes, MODE_COUNT, nullptr, 0, rh)) {
                g_head_ok++;
                const bool same = memcmp(&rh.in, &kin, sizeof(Carry)) == 0 && memcmp(&rh.out, &rc.out, sizeof(Carry)) == 0 &&
                                  rh.nrec == rc.nrec && rh.ntext == rc.ntext && rh.npend_out == rc.npend_out;
                if (!same) {
                    stats[3] += 1000000;
                    if (g_mask_mismatch++ < 5)
                        fprintf(stderr, "one-pass head mismatch: window [%lld,%lld) kin k=%d/%d in=%u/%u out=%u/%u nrec %u/%u\n", (long long)wg.ws,
                                (long long)wg.we, rh.in.k, kin.k, rh.in.in_bytes, kin.in_bytes, rh.in.out_bytes, kin.out_bytes, rh.nrec, rc.nrec);
                }
            }
        }
        if (MaskFamily<Dec>::kHas && !P.general && g_use_mask && g_use_fast) {
            // the sparse pipeline's closed form for windows that are one short run (WinResult.caseb of a pass under the
            // null carry): eval_caseb from the real carry-in must give the real carry-out
            WinResult rn;
            const bool rn_ok = mask_window<MaskFamily<Dec>::kSByte>(P, ts, wg, carry_none(), MODE_STATE, nullptr, 0, rn);
            if (rn_ok && rn.cut1 == 0) {
                // "the carry-out does not depend on the carry-in" (sx_sp_members_kernel takes the null-carry result as the
                // next entry's carry-in): must equal the carry-out under the real carry-in
                g_indep++;
                if (memcmp(&rn.out, &rs.out, sizeof(Carry)) != 0) {
                    stats[3] += 100000000;
                    if (g_mask_mismatch++ < 5)
                        fprintf(stderr, "independence mismatch: window [%lld,%lld) kin kind=%d k=%d: null-carry out kind=%d k=%d vs real kind=%d k=%d\n",
                                (long long)wg.ws, (long long)wg.we, kin.kind, kin.k, rn.out.kind, rn.out.k, rs.out.kind, rs.out.k);
                }
            } else if (rn_ok) g_dep++;
            if (rn_ok && rn.caseb) {
                WinDesc dc;
                dc.type = WT_CASEB; dc.pad = 0; dc.a = rn.a; dc.t_out = rn.t_out; dc.nrec = 0; dc.ntext = 0; dc.null_out = rn.out;
                const Carry ke = eval_caseb(P, dc, kin, (uint32_t)(wg.we - wg.ws));
                if (memcmp(&ke, &rs.out, sizeof(Carry)) != 0) {
                    stats[3] += 10000000;
                    if (g_mask_mismatch++ < 5)
                        fprintf(stderr, "caseb closed form mismatch: window [%lld,%lld) kin kind=%d k=%d -> kind %d/%d k %d/%d in %u/%u out %u/%u\n",
                                (long long)wg.ws, (long long)wg.we, kin.kind, kin.k, ke.kind, rs.out.kind, ke.k, rs.out.k, ke.in_bytes,
                                rs.out.in_bytes, ke.out_bytes, rs.out.out_bytes);
                }
            }
        }
        if (!needs_emit(P, d, kin) && rc.nrec != 0) stats[4]++;
        if (carry_is_null(kin) && d.nrec != 0xFFFF && (rc.nrec != d.nrec || rc.ntext != d.ntext)) stats[5]++;
        last_npend = rs.npend_out;
        kprev = kout;
        // extension: a "cut" carry reaches an unlisted successor
        const bool next_adjacent = e + 1 < ne && (int64_t)list[e + 1] == w + 1;
        if (carry_needs_extension(P, kout) && !next_adjacent && w + 1 < total) {
            WinGeom xf;
            geo.window(w + 1, xf);
            WinResult rf;
            WindowEngine<Dec>::run(P, ts, g, xf, kout, MODE_COUNT, nullptr, 0, rf, nullptr);
            if (carry_needs_extension(P, rf.out)) stats[3] += 1000;  // must never happen
            // the product only reads the head of the window (ext_geom): same records as the full window
            const WinGeom xg = ext_geom(geo, w + 1, pre_bytes);
            WinResult rx;
            emit_window(xg, kout, rx);
            if (rx.nrec != rf.nrec || rx.ntext != rf.ntext) stats[3] += 100000;
        }
    }
    *final_carry = ne ? kprev : P.k0;
    *final_npend = last_npend;
    if (final_carry->kind == K_L && final_carry->k > 0) {
        Record r{};
        r.in_start = P.len - (int64_t)final_carry->in_bytes;
        r.in_len = final_carry->in_bytes - (uint32_t)*final_npend;
        r.text_len = final_carry->out_bytes;
        r.text_off = text.size();
        r.flags = RF_LEFTOVER | ((final_carry->flags & CF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u) |
                  ((final_carry->flags & CF_HALF) ? (uint32_t)RF_HALFSTART : 0u);
        text.resize(text.size() + r.text_len + 8);
        transcode_record(P, g, r, text.data() + r.text_off);
        text.resize(text.size() - 8);
        recs.push_back(r);
    }
}

extern "C" {
struct emul_out {
    Record* recs;
    size_t nrecs;
    uint8_t* text;
    size_t ntext;
    Carry final_carry;
    int32_t final_npend;
    uint64_t stats[8];
    uint32_t* list;
    size_t nlist;
};

// Returns 0 on success.  `params` is a fully populated ScanParams (in = host pointer).
void sx_emul_set_fast(int on) { g_use_fast = on; }
void sx_emul_set_mask(int on) { g_use_mask = on; }
void sx_emul_mask_counts(uint64_t* ok, uint64_t* declined) { *ok = g_mask_ok; *declined = g_mask_declined; }
uint64_t sx_emul_mask_mismatches() { return g_mask_mismatch; }
uint64_t sx_emul_head_ok() { return g_head_ok; }
void sx_emul_indep_counts(uint64_t* indep, uint64_t* dep) { *indep = g_indep; *dep = g_dep; }
void sx_emul_guard_counts(uint64_t* ok, uint64_t* killer) { *ok = g_guard_ok; *killer = g_guard_killer; }
uint64_t sx_emul_guard_behind() { return g_guard_behind; }
uint64_t sx_emul_guard_behind_hard() { return g_guard_behind_hard; }

int sx_emul_scan(const ScanParams* P_in, int use_pref, emul_out* out) {
    ScanParams Pl = *P_in;
    static uint32_t mb_fail;
    mb_fail = 0;
    if (Pl.enc == ENC_BIG5) { Pl.mb_a = kSxBig5Index; Pl.mb_b = nullptr; }
    if (Pl.enc == ENC_EUCJP) { Pl.mb_a = kSxJis0208Index; Pl.mb_b = kSxJis0212Index; }
    Pl.mb_fail = &mb_fail;
    const ScanParams* P = &Pl;
    for (uint32_t i = 0; i < 2048; ++i) mask_tables_fill(*P, g_tables, i);
    std::vector<uint32_t> list;
    std::vector<Record> recs;
    std::vector<uint8_t> text;
    memset(out->stats, 0, sizeof out->stats);
    switch (P->enc) {
    case ENC_XUD: run<DecXud>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_UTF8: run<DecUtf8>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_UTF16LE: run<DecUtf16<false>>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_UTF16BE: run<DecUtf16<true>>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_SB: run<DecSb>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_UTF32LE: run<DecUtf32<false>>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_UTF32BE: run<DecUtf32<true>>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_BIG5: run<DecBig5>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    case ENC_EUCJP: run<DecEucJp>(*P, use_pref, recs, text, &out->final_carry, &out->final_npend, out->stats, list); break;
    default: return 1;
    }
    if (mb_fail) return 2;
    out->nrecs = recs.size();
    out->recs = (Record*)malloc(sizeof(Record) * (recs.size() + 1));
    memcpy(out->recs, recs.data(), sizeof(Record) * recs.size());
    out->ntext = text.size();
    out->text = (uint8_t*)malloc(text.size() + 1);
    memcpy(out->text, text.data(), text.size());
    out->nlist = list.size();
    out->list = (uint32_t*)malloc(sizeof(uint32_t) * (list.size() + 1));
    memcpy(out->list, list.data(), sizeof(uint32_t) * list.size());
    return 0;
}
void sx_emul_free(emul_out* o) { free(o->recs); free(o->text); free(o->list); }
size_t sx_emul_sizeof_params() { return sizeof(ScanParams); }
}
