// sx_fast_utf8.cuh -- convergent (branch-light) window automaton for UTF-8 missions.
//
// Same semantics as scan_window<DecUtf8> in sx_core.cuh (the streaming restatement of
// FindingCollection::from + SplitStr::next, /root/reference/src/finding_collection.rs:84-342,
// /root/reference/src/helper.rs:210-432), reorganised for a GPU lane:
//   * the WHATWG UTF-8 decoder is a table-driven DFA (byte class LUT + 8x12 transition LUT in shared
//     memory), so every lane executes the same instruction stream whatever its bytes are,
//   * positions are 32-bit offsets relative to the window start,
//   * the common transitions (passing char, short run dropped at a breaker, malformed sequence
//     with nothing to print) are a handful of predicated instructions; everything that prints or
//     cuts (yield, q-cut, leftover, probe) is out of line.
// tests/emul cross-checks this engine against the generic one and against the oracle.
#pragma once
#include "sx_core.cuh"

namespace sx {

struct Utf8Tables {
    uint8_t cls[256];    // byte class 0..11
    uint8_t pass[256];   // Utf8Filter verdict for a char whose UTF-8 lead byte is b (mission.rs:333-348)
    uint8_t trans[128];  // [state << 4 | class]: bits 0-2 next state, 3-4 event, 5 pre-malformed, 6 lead latch
};
enum : uint32_t { FE_NONE = 0, FE_ASCII = 1, FE_CHAR = 2, FE_MAL = 3 };
enum : uint32_t { FT_PRE = 0x20, FT_LEAD = 0x40 };

SX_HD uint32_t utf8_class(uint32_t b) {
    if (b < 0x80) return 0;
    if (b < 0x90) return 1;
    if (b < 0xA0) return 2;
    if (b < 0xC0) return 3;
    if (b < 0xC2 || b > 0xF4) return 4;
    if (b < 0xE0) return 5;
    if (b == 0xE0) return 6;
    if (b == 0xED) return 8;
    if (b < 0xF0) return 7;
    if (b == 0xF0) return 9;
    if (b == 0xF4) return 11;
    return 10;
}
// states: 0 neutral, 1 one more continuation (80-BF), 2 two more, 3 after E0 (A0-BF), 4 after ED (80-9F),
//         5 after F0 (90-BF), 6 after F1-F3 (80-BF), 7 after F4 (80-8F)
SX_HD uint32_t utf8_neutral_entry(uint32_t c) {
    switch (c) {
    case 0: return 0 | (FE_ASCII << 3);
    case 1: case 2: case 3: case 4: return 0 | (FE_MAL << 3);
    case 5: return 1 | FT_LEAD;
    case 6: return 3 | FT_LEAD;
    case 7: return 2 | FT_LEAD;
    case 8: return 4 | FT_LEAD;
    case 9: return 5 | FT_LEAD;
    case 10: return 6 | FT_LEAD;
    default: return 7 | FT_LEAD;
    }
}
SX_HD uint32_t utf8_trans_entry(uint32_t s, uint32_t c) {
    if (s == 0) return utf8_neutral_entry(c);
    const bool c1 = c == 1, c2 = c == 2, c3 = c == 3;
    bool ok = false;
    uint32_t ns = 0, ev = FE_NONE;
    switch (s) {
    case 1: ok = c1 || c2 || c3; ns = 0; ev = FE_CHAR; break;
    case 2: ok = c1 || c2 || c3; ns = 1; break;
    case 3: ok = c3; ns = 1; break;
    case 4: ok = c1 || c2; ns = 1; break;
    case 5: ok = c2 || c3; ns = 2; break;
    case 6: ok = c1 || c2 || c3; ns = 2; break;
    default: ok = c1; ns = 2; break;
    }
    if (ok) return ns | (ev << 3);
    return utf8_neutral_entry(c) | FT_PRE;  // the offending byte is not consumed: malformed first, then re-read in the neutral state
}
SX_HD void utf8_tables_fill(const ScanParams& P, Utf8Tables& T, uint32_t i) {  // i in 0..255
    T.cls[i] = (uint8_t)utf8_class(i);
    T.pass[i] = (i < 0x80 || i >= 0xC0) ? (pass_filter(P, i) ? 1 : 0) : 0;
    if (i < 128) T.trans[i] = (uint8_t)(((i & 15) < 12) ? utf8_trans_entry(i >> 4, i & 15) : 0);
}

struct FastAuto {
    const ScanParams* P;
    int64_t base;       // window start (absolute); all *_rel are relative to it
    int mode;
    Record* wr;
    uint64_t text_off;
    int32_t slice_rel;  // slice start relative to base
    // segment / SplitStr state
    int32_t seg_rel;
    uint32_t prec;
    bool probe_pending, last_cut, at_left, cut, run_hostcarry;
    uint32_t run_n;
    int32_t run_s, run_e;
    // leftover (`again` chunk)
    bool has_left, left_hostcarry;
    uint32_t left_k;
    int32_t left_s, left_e;
    Carry slice_left;
    // outputs
    uint32_t nrec, ntext;
    // summary
    uint32_t m, a;
    bool in_first_run, s1_all_pass, s1_later_yield, s2_all_pass;

    SX_HD void init(const ScanParams* p, int md, int64_t b, int32_t srel) {
        P = p; base = b; mode = md; wr = nullptr; text_off = 0; slice_rel = srel;
        seg_rel = 0; prec = PREC_EXACT; probe_pending = false; last_cut = false; at_left = true; cut = false;
        run_hostcarry = false; run_n = 0; run_s = run_e = 0;
        has_left = false; left_hostcarry = false; left_k = 0; left_s = left_e = 0; slice_left = carry_none();
        nrec = ntext = 0; m = 0; a = 0; in_first_run = false; s1_all_pass = true; s1_later_yield = false; s2_all_pass = true;
    }
    // first segment of the window (finding_collection.rs:101-116, :211-241)
    SX_HD void first_segment(const Carry& kin, int32_t pend_len) {
        m = 1;
        const bool cont = (kin.kind == K_C);
        cut = false;
        last_cut = cont;
        at_left = true;
        run_n = 0; run_hostcarry = false;
        prec = PREC_EXACT;
        seg_rel = 0;
        probe_pending = (slice_rel == 0);
        has_left = false;
        in_first_run = true;
        if (kin.kind == K_L && kin.k > 0) {
            run_n = kin.k;
            run_s = -(int32_t)kin.in_bytes;
            run_e = -pend_len;
            run_hostcarry = (kin.flags & CF_HOSTCARRY) != 0;
            prec = PREC_BEFORE;
        }
    }
    // rare paths below stay inline on purpose: taking the automaton's address for a call would push its
    // state from registers into local memory
    SX_HD void yield(bool completes, bool maybe_cut) {
        if (m == 1 && !in_first_run) s1_later_yield = true;
        const uint32_t len = (uint32_t)(run_e - run_s);
        if (mode == MODE_WRITE) {
            Record r;
            r.position = P->base_consumed + (uint64_t)(base + seg_rel);
            r.in_start = base + run_s;
            r.in_len = len;
            r.text_len = len;  // UTF-8 -> UTF-8: the text is the input range
            r.text_off = text_off;
            r.flags = (completes ? RF_COMPLETES : 0u) | (run_hostcarry ? RF_HOSTCARRY : 0u);
            r.precision = prec;
            *wr++ = r;
            text_off += len;
        }
        nrec++;
        ntext += len;
        cut = maybe_cut;
        prec = PREC_AFTER;
        has_left = false;
    }
    // end of a segment's text (helper.rs:343-431 for the run touching the right boundary)
    SX_HD void segment_end(bool invalid_after) {
        if (run_n > 0) {
            const bool completes = at_left && last_cut;
            const bool again = !completes && !invalid_after;
            if (again) {
                if (m == 1 && !in_first_run) s1_later_yield = true;
                has_left = true;
                left_k = run_n; left_s = run_s; left_e = run_e; left_hostcarry = run_hostcarry;
                cut = false;
            } else if (completes || run_n >= P->n) {
                yield(completes, !invalid_after);
            }
        }
        if (m == 1) in_first_run = false;
        run_n = 0;
    }
    SX_HD void new_segment(int32_t next_rel) {  // finding_collection.rs:240-241 + helper.rs:171-200
        m++;
        last_cut = cut;
        cut = false;
        at_left = true;
        run_n = 0; run_hostcarry = false;
        prec = PREC_EXACT;
        seg_rel = next_rel;
        probe_pending = (next_rel == slice_rel);
        has_left = false;
    }
    // a malformed sequence: the segment ends here and the next one starts at next_rel
    SX_HD void brk(int32_t next_rel) {
        if (run_n > 0 && (run_n >= P->n || (at_left && last_cut))) segment_end(true);
        if (m == 1) in_first_run = false;
        new_segment(next_rel);
    }
    SX_HD void qcut() {  // helper.rs:237 exit 2, :353-355, :418-421
        yield(at_left && last_cut, true);
        at_left = true;
        last_cut = true;
        run_n = 0; run_hostcarry = false;
    }
    SX_HD void breaker_yield() {
        yield(at_left && last_cut, false);
        last_cut = false;
    }
    SX_HD void chr(bool pass, int32_t cs, int32_t ce) {
        if (pass) {
            if (run_n == 0) run_s = cs;
            run_n++;
            run_e = ce;
            if (in_first_run && a < 0xFFFFu) a++;
            if (run_n >= P->q) qcut();
        } else {
            if (m == 1) s1_all_pass = false;
            if (m == 2) s2_all_pass = false;
            if (run_n > 0 && ((last_cut && at_left) || run_n >= P->n)) breaker_yield();
            in_first_run = false;
            run_n = 0; run_hostcarry = false;
            at_left = false;
        }
    }
    SX_HD Carry carry_out(int32_t boundary_rel) const {
        if (cut) return carry_cut();
        if (has_left) {
            Carry c;
            c.kind = K_L; c.flags = left_hostcarry ? CF_HOSTCARRY : 0; c.k = (uint16_t)left_k;
            c.in_bytes = (uint32_t)(boundary_rel - left_s);
            c.out_bytes = (uint32_t)(left_e - left_s);
            return c;
        }
        return carry_none();
    }
};

// TileSrc additionally provides `const Utf8Tables* tables()`.
template <class TileSrc>
SX_HD_NOINLINE void scan_window_fast_utf8(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo,
                                          const Carry& kin, int mode, Record* wr, uint64_t text_off, WinResult& res,
                                          WinDesc* desc) {
    const Utf8Tables& T = *tsrc.tables();
    // decoder state at the window start from the preceding bytes (DecUtf8::init), mapped onto the table DFA
    DecUtf8 d0;
    d0.init(P, tsrc, geo.ws);
    uint32_t st = 0;
    bool cur_pass = false;
    int32_t seq_s = 0;
    const int32_t pend0 = d0.pending_len();
    if (d0.need) {
        const uint32_t rem = d0.need - d0.seen;
        if (d0.seen == 0) st = utf8_neutral_entry(utf8_class(d0.lead)) & 7u;
        else st = rem;  // 1 or 2 plain continuation bytes left
        cur_pass = T.pass[d0.lead] != 0;
        seq_s = -pend0;
    }
    FastAuto A;
    A.init(&P, mode, geo.ws, (int32_t)(geo.slice_start - geo.ws));
    A.wr = wr; A.text_off = text_off;
    if (geo.ws == geo.slice_start) A.slice_left = kin;
    A.first_segment(kin, pend0);
    const int32_t wlen = (int32_t)(geo.we - geo.ws);
    tsrc.for_each_byte(geo.ws, geo.we, [&](uint32_t b, int64_t pos) {
        const int32_t p = (int32_t)(pos - geo.ws);
        const uint32_t t = T.trans[(st << 4) | T.cls[b]];
        st = t & 7u;
        if (t & FT_PRE) A.brk(p);
        const uint32_t ev = (t >> 3) & 3u;
        if (ev == FE_ASCII) {
            A.probe_pending = false;  // an ASCII first char never triggers the probe (finding_collection.rs:176)
            A.chr(T.pass[b] != 0, p, p + 1);
        } else if (ev == FE_CHAR) {
            if (A.probe_pending) {
                A.probe_pending = false;
                if (mode != MODE_STATE) {
                    const Carry sl = A.slice_left;  // a copy: the probe is out of line and takes references
                    if (probe_utf8(P, g, geo.slice_start, geo.slice_end, A.m == 1, pend0, sl)) A.prec = PREC_BEFORE;
                }
            }
            A.chr(cur_pass, seq_s, p + 1);
        } else if (ev == FE_MAL) {
            A.brk(p + 1);
        }
        if (t & FT_LEAD) { cur_pass = T.pass[b] != 0; seq_s = p; }
    });
    if (geo.final_last) {
        A.segment_end(true);
        A.cut = false;
        A.has_left = false;
        res.npend_out = 0;
    } else {
        A.segment_end(false);
        res.npend_out = st ? (wlen - seq_s) : 0;
    }
    res.out = A.carry_out(wlen);
    res.nrec = A.nrec;
    res.ntext = A.ntext;
    if (desc) {
        desc->a = (uint16_t)A.a;
        desc->nrec = (uint16_t)(A.nrec > 0xFFFFu ? 0xFFFFu : A.nrec);
        desc->ntext = A.ntext;
        desc->null_out = res.out;
        desc->t_out = 0;
        desc->pad = 0;
        const bool single_all_pass = (A.m == 1 && A.s1_all_pass);
        if (geo.final_last) desc->type = WT_CONST;
        else if (single_all_pass) {
            if (A.a < P.q) { desc->type = WT_CASEB; desc->t_out = (uint16_t)(A.has_left ? (A.left_e - A.left_s) : 0); }
            else desc->type = WT_CONST;
        } else if (A.a > 0 && !A.s1_later_yield && (A.m == 1 || (A.m == 2 && A.s2_all_pass))) desc->type = WT_DEP;
        else desc->type = WT_CONST;
    }
}

// Engine dispatch used by the kernels and the test harness: UTF-8 takes the convergent engine.
template <class Dec> struct WindowEngine {
    template <class TileSrc>
    SX_HD static void run(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo, const Carry& kin,
                          int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
        scan_window<Dec>(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
    }
};
template <> struct WindowEngine<DecUtf8> {
    template <class TileSrc>
    SX_HD static void run(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo, const Carry& kin,
                          int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
#if defined(__CUDA_ARCH__)
        scan_window_fast_utf8(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);  // device: always the convergent engine
#else
        if (tsrc.tables()) scan_window_fast_utf8(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
        else scan_window<DecUtf8>(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
#endif
    }
};

}  // namespace sx
