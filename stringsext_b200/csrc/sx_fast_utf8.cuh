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

// One lookup per byte: tt[state << 8 | byte] = bits 0-2 next state, 3-4 event, 5 pre-malformed (the byte
// offends a pending sequence: malformed first, then the byte is re-read in the neutral state), 6 lead latch,
// 7 Utf8Filter verdict for a char whose UTF-8 lead byte is this byte (mission.rs:333-348).
struct Utf8Tables {
    uint8_t tt[8 * 256];
    uint8_t cls[256];  // mask engine (sx_mask_utf8.cuh): bits 0-3 utf8_class(b), bit 4 the filter verdict for lead byte b
};
enum : uint32_t { FE_NONE = 0, FE_ASCII = 1, FE_CHAR = 2, FE_MAL = 3 };
enum : uint32_t { FT_PRE = 0x20, FT_LEAD = 0x40, FT_PASS = 0x80 };

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
    return utf8_neutral_entry(c) | FT_PRE;
}
SX_HD void utf8_tables_fill(const ScanParams& P, Utf8Tables& T, uint32_t i) {  // i in 0..2047
    const uint32_t b = i & 255u, st = i >> 8;
    uint32_t t = utf8_trans_entry(st, utf8_class(b));
    if ((b < 0x80 || b >= 0xC0) && pass_filter(P, b)) t |= FT_PASS;
    T.tt[i] = (uint8_t)t;
    if (st == 0) T.cls[b] = (uint8_t)(utf8_class(b) | ((t & FT_PASS) ? 16u : 0u));
}

// Cold state: only the out-of-line record writer touches it, so it may live in local memory while the
// hot automaton state below stays in registers.
struct FastEmit {
    const ScanParams* P;
    int64_t base;  // window start (absolute)
    int mode;
    Record* wr;
    uint64_t text_off;
    uint32_t nrec, ntext;
};
SX_HD_NOINLINE void fast_emit(FastEmit* E, int32_t seg_rel, uint32_t prec, int32_t run_s, int32_t run_e, uint32_t completes,
                              uint32_t hostcarry) {
    const uint32_t len = (uint32_t)(run_e - run_s);
    if (E->mode == MODE_WRITE || (E->mode == MODE_BUFFER && E->nrec < kBufRecs)) {
        Record r;
        r.position = E->P->base_consumed + (uint64_t)(E->base + seg_rel);  // finding_collection.rs:260
        r.in_start = E->base + run_s;
        r.in_len = len;
        r.text_len = len;  // UTF-8 -> UTF-8: the text is the input range
        r.text_off = E->text_off;
        r.flags = (completes ? RF_COMPLETES : 0u) | (hostcarry ? RF_HOSTCARRY : 0u);
        r.precision = prec;
        *E->wr++ = r;
        E->text_off += len;
    }
    E->nrec++;
    E->ntext += len;
}

// hot-state flag bits
enum : uint32_t {
    FF_LASTCUT = 1u << 0,    // SplitStr.last_s_was_maybe_cut
    FF_ATLEFT = 1u << 1,     // ok_s_p == inp_start_p for the current run
    FF_CUT = 1u << 2,        // last_window_str_was_printed_and_is_maybe_cut_str
    FF_PROBE = 1u << 3,      // Precision::Before probe still possible for this segment
    FF_HOSTCARRY = 1u << 4,  // current run starts with text carried in the host ScannerState
    FF_INFIRST = 1u << 5,    // inside the first run of segment 1
    FF_S1ALL = 1u << 6,
    FF_S1LATER = 1u << 7,
    FF_S2ALL = 1u << 8,
    FF_HASLEFT = 1u << 9,
    FF_LEFTHC = 1u << 10
};

// TileSrc additionally provides `const Utf8Tables* tables()`, `lut(i)` (= tables()->tt[i], a 32-bit shared
// memory load on the device) and `load_chunk(r16, ws, we)` (the 16 input bytes at aligned offset r16).
template <class TileSrc>
SX_HD_NOINLINE void scan_window_fast_utf8(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo,
                                          const Carry& kin, int mode, Record* wr, uint64_t text_off, WinResult& res,
                                          WinDesc* desc) {
    const uint32_t n = P.n, q = P.q;
    // decoder state at the window start from the preceding bytes (DecUtf8::init), mapped onto the table DFA
    DecUtf8 d0;
    d0.init(P, tsrc, geo.ws);
    uint32_t st = 0, cur_pass = 0;
    int32_t seq_s = 0;
    const int32_t pend0 = d0.pending_len();
    if (d0.need) {
        st = d0.seen == 0 ? (utf8_neutral_entry(utf8_class(d0.lead)) & 7u) : (d0.need - d0.seen);
        cur_pass = tsrc.lut(d0.lead) & FT_PASS;
        seq_s = -pend0;
    }
    FastEmit E;
    E.P = &P; E.base = geo.ws; E.mode = mode; E.wr = wr; E.text_off = text_off; E.nrec = 0; E.ntext = 0;
    const int32_t slice_rel = (int32_t)(geo.slice_start - geo.ws);
    const int32_t wlen = (int32_t)(geo.we - geo.ws);
    const Carry slice_left = kin;  // only read by the probe when the window starts a slice

    // ---- first segment (finding_collection.rs:101-116, :211-241) ----
    uint32_t fl = FF_ATLEFT | FF_INFIRST | FF_S1ALL | FF_S2ALL;
    if (kin.kind == K_C) fl |= FF_LASTCUT;
    if (slice_rel == 0) fl |= FF_PROBE;
    uint32_t m = 1, a = 0, run_n = 0, prec = PREC_EXACT;
    int32_t seg_rel = 0, run_s = 0, run_e = 0;
    uint32_t left_k = 0;
    int32_t left_s = 0, left_e = 0;
    if (kin.kind == K_L && kin.k > 0) {
        run_n = kin.k;
        run_s = -(int32_t)kin.in_bytes;
        run_e = -pend0;
        if (kin.flags & CF_HOSTCARRY) fl |= FF_HOSTCARRY;
        prec = PREC_BEFORE;
    }

#define SX_COMPLETES ((fl & (FF_ATLEFT | FF_LASTCUT)) == (FF_ATLEFT | FF_LASTCUT))
#define SX_YIELD(completes_, maybe_cut_)                                                              \
    do {                                                                                              \
        if (m == 1 && !(fl & FF_INFIRST)) fl |= FF_S1LATER;                                           \
        fast_emit(&E, seg_rel, prec, run_s, run_e, (completes_) ? 1u : 0u, fl & FF_HOSTCARRY);        \
        fl = (maybe_cut_) ? (fl | FF_CUT) : (fl & ~FF_CUT);                                           \
        fl &= ~FF_HASLEFT;                                                                            \
        prec = PREC_AFTER;                                                                            \
    } while (0)
    // a malformed sequence: the segment ends (invalid_after) and the next one starts at next_rel
#define SX_BRK(next_rel_)                                                                             \
    do {                                                                                              \
        if (run_n > 0 && (run_n >= n || SX_COMPLETES)) { const bool c_ = SX_COMPLETES; SX_YIELD(c_, false); } \
        m++;                                                                                          \
        fl = (fl & ~(FF_LASTCUT | FF_INFIRST | FF_HOSTCARRY | FF_HASLEFT | FF_PROBE | FF_CUT)) | FF_ATLEFT | \
             ((fl & FF_CUT) ? FF_LASTCUT : 0u) | (((next_rel_) == slice_rel) ? FF_PROBE : 0u);        \
        run_n = 0;                                                                                    \
        prec = PREC_EXACT;                                                                            \
        seg_rel = (next_rel_);                                                                        \
    } while (0)
#define SX_CHR(pass_, cs_, ce_)                                                                       \
    do {                                                                                              \
        if (pass_) {                                                                                  \
            if (run_n == 0) run_s = (cs_);                                                            \
            run_n++;                                                                                  \
            run_e = (ce_);                                                                            \
            if (fl & FF_INFIRST) a++;                                                                 \
            if (run_n >= q) { /* helper.rs:237 exit 2, :353-355, :418-421 */                          \
                const bool c_ = SX_COMPLETES;                                                         \
                SX_YIELD(c_, true);                                                                   \
                fl = (fl | FF_ATLEFT | FF_LASTCUT) & ~FF_HOSTCARRY;                                   \
                run_n = 0;                                                                            \
            }                                                                                         \
        } else {                                                                                      \
            if (m == 1) fl &= ~FF_S1ALL;                                                              \
            if (m == 2) fl &= ~FF_S2ALL;                                                              \
            if (run_n > 0 && (run_n >= n || SX_COMPLETES)) { /* helper.rs:315-322 */                   \
                const bool c_ = SX_COMPLETES;                                                         \
                SX_YIELD(c_, false);                                                                  \
                fl &= ~FF_LASTCUT;                                                                    \
            }                                                                                         \
            fl &= ~(FF_INFIRST | FF_HOSTCARRY | FF_ATLEFT);                                           \
            run_n = 0;                                                                                \
        }                                                                                             \
    } while (0)

    // ---- the byte loop: 16-byte chunks, three in flight, one copy of the body (rolled on purpose) ----
    {
        const int64_t ws = geo.ws, we = geo.we;
        int64_t r16 = ws & ~(int64_t)15;
        int32_t p = (int32_t)(r16 - ws);  // relative position of the chunk's first byte (may be negative for the head)
        uint4 c0 = tsrc.load_chunk(r16, ws, we);
        uint4 c1 = tsrc.load_chunk(r16 + 16, ws, we);
        uint4 c2 = tsrc.load_chunk(r16 + 32, ws, we);
        while (p < wlen) {
            const uint4 c3 = tsrc.load_chunk(r16 + 48, ws, we);
            uint32_t w0 = c0.x, w1 = c0.y, w2 = c0.z, w3 = c0.w;
            const int32_t pe = (wlen - p) < 16 ? (wlen - p) : 16;
#pragma unroll 1
            for (int32_t i = 0; i < pe; ++i, ++p) {
                const uint32_t b = w0 & 0xFFu;
                w0 = (w0 >> 8) | (w1 << 24);  // 128-bit shift right by one byte (SHF on the device)
                w1 = (w1 >> 8) | (w2 << 24);
                w2 = (w2 >> 8) | (w3 << 24);
                w3 >>= 8;
                if (p < 0) continue;  // bytes before the window inside the first aligned chunk
                const uint32_t t = tsrc.lut((st << 8) | b);
                st = t & 7u;
                const uint32_t ev = (t >> 3) & 3u;
                const bool pre = (t & FT_PRE) != 0;
                const bool is_ascii = ev == FE_ASCII, is_char = ev == FE_CHAR, mal = ev == FE_MAL;
                const bool pass = (is_ascii && (t & FT_PASS)) || (is_char && cur_pass);
                const bool fail = (is_ascii || is_char) && !pass;
                // Anything that prints, cuts or probes takes the (rare) general path below; every other
                // transition is computed branch-free so that the 32 lanes of a warp stay converged.
                const bool run_ends = pre || mal || fail;
                const bool slow = (run_ends && run_n > 0 && (run_n >= n || SX_COMPLETES)) || (pass && !pre && run_n + 1 >= q) ||
                                  (pass && pre && 1 >= q) || (is_char && (fl & FF_PROBE));
                if (!slow) {
                    const uint32_t nbrk = (pre ? 1u : 0u) + (mal ? 1u : 0u);
                    if (nbrk) {  // malformed sequence(s): new segment (predicated, no yield needed)
                        const int32_t nx = mal ? p + 1 : p;
                        m += nbrk;
                        fl = (fl & ~(FF_LASTCUT | FF_INFIRST | FF_HOSTCARRY | FF_HASLEFT | FF_PROBE | FF_CUT)) | FF_ATLEFT |
                             ((nbrk == 1 && (fl & FF_CUT)) ? FF_LASTCUT : 0u) | ((nx == slice_rel) ? FF_PROBE : 0u);
                        run_n = 0;
                        prec = PREC_EXACT;
                        seg_rel = nx;
                    }
                    if (is_ascii) fl &= ~FF_PROBE;
                    if (pass) {
                        const int32_t cs = is_ascii ? p : seq_s;
                        run_s = run_n == 0 ? cs : run_s;
                        run_n++;
                        run_e = p + 1;
                        a += (fl & FF_INFIRST) ? 1u : 0u;
                    }
                    if (fail) {
                        fl &= ~((m == 1 ? FF_S1ALL : 0u) | (m == 2 ? FF_S2ALL : 0u) | FF_INFIRST | FF_HOSTCARRY | FF_ATLEFT);
                        run_n = 0;
                    }
                } else {
                    if (t & FT_PRE) SX_BRK(p);
                    if (ev == FE_ASCII) {
                        fl &= ~FF_PROBE;  // an ASCII first char never triggers the probe (finding_collection.rs:176)
                        SX_CHR(t & FT_PASS, p, p + 1);
                    } else if (ev == FE_CHAR) {
                        if (fl & FF_PROBE) {
                            fl &= ~FF_PROBE;
                            if (mode != MODE_STATE && probe_utf8(P, g, geo.slice_start, geo.slice_end, m == 1, pend0, slice_left))
                                prec = PREC_BEFORE;
                        }
                        SX_CHR(cur_pass, seq_s, p + 1);
                    } else if (ev == FE_MAL) {
                        SX_BRK(p + 1);
                    }
                }
                if (t & FT_LEAD) { cur_pass = t & FT_PASS; seq_s = p; }
            }
            r16 += 16;
            c0 = c1; c1 = c2; c2 = c3;
        }
    }

    // ---- end of the window's last segment (helper.rs:343-431 for the run touching the right boundary) ----
    const bool invalid_after = geo.final_last;
    if (run_n > 0) {
        const bool completes = SX_COMPLETES;
        if (!completes && !invalid_after) {  // `again`: kept as leftover (finding_collection.rs:281-284)
            if (m == 1 && !(fl & FF_INFIRST)) fl |= FF_S1LATER;
            fl |= FF_HASLEFT;
            fl = (fl & FF_HOSTCARRY) ? (fl | FF_LEFTHC) : (fl & ~FF_LEFTHC);
            left_k = run_n; left_s = run_s; left_e = run_e;
            fl &= ~FF_CUT;
        } else if (completes || run_n >= n) {
            SX_YIELD(completes, !invalid_after);
        }
    }
#undef SX_CHR
#undef SX_BRK
#undef SX_YIELD
#undef SX_COMPLETES
    if (geo.final_last) {
        // finding_collection.rs:298-304: one flush round; lasting effects: reset decoder, cut == false
        fl &= ~(FF_CUT | FF_HASLEFT);
        res.npend_out = 0;
    } else {
        res.npend_out = st ? (wlen - seq_s) : 0;
    }
    if (fl & FF_CUT) res.out = carry_cut();
    else if (fl & FF_HASLEFT) {
        Carry c;
        c.kind = K_L; c.flags = (fl & FF_LEFTHC) ? CF_HOSTCARRY : 0; c.k = (uint16_t)left_k;
        c.in_bytes = (uint32_t)(wlen - left_s);
        c.out_bytes = (uint32_t)(left_e - left_s);
        c.aux = 0;
        res.out = c;
    } else res.out = carry_none();
    res.nrec = E.nrec;
    res.ntext = E.ntext;
    res.m = m;
    res.cut1 = 0;
    if (desc) {
        desc->a = (uint16_t)(a > 0xFFFFu ? 0xFFFFu : a);
        desc->nrec = (uint16_t)(E.nrec > 0xFFFFu ? 0xFFFFu : E.nrec);
        desc->ntext = E.ntext;
        desc->null_out = res.out;
        desc->t_out = 0;
        desc->pad = 0;
        const bool single_all_pass = (m == 1 && (fl & FF_S1ALL));
        if (geo.final_last) desc->type = WT_CONST;
        else if (single_all_pass) {
            if (a < q) { desc->type = WT_CASEB; desc->t_out = (uint16_t)((fl & FF_HASLEFT) ? (left_e - left_s) : 0); }
            else desc->type = WT_CONST;
        } else if (a > 0 && !(fl & FF_S1LATER) && (m == 1 || (m == 2 && (fl & FF_S2ALL)))) desc->type = WT_DEP;
        else desc->type = WT_CONST;
    }
}

}  // namespace sx
