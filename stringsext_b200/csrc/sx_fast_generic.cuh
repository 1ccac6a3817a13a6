// sx_fast_generic.cuh -- convergent (branch-light) window automaton for the decoders WITHOUT a bit-parallel engine
// (UTF-16LE/BE, UTF-32LE/BE, Big5, EUC-JP), plain missions only (no grep_char, no same-unicode-block, n <= q).
//
// Same semantics as scan_window<Dec> in sx_core.cuh (the streaming restatement of FindingCollection::from +
// SplitStr::next, /root/reference/src/finding_collection.rs:84-342, /root/reference/src/helper.rs:210-432), organised
// like sx_fast_utf8.cuh.  Per input byte the decoder yields at most two events ("malformed, then the byte again",
// "char + second code point of a Big5 pair", "malformed + pending BMP unit"): GenStep<Dec> computes them with selects
// and one exit (UTF-32 runs its own step() against the recording emitter), and the automaton's common transitions --
// passing char, short run dropped at a breaker, malformed sequence with nothing to print -- are blends under
// all-ones / all-zero masks, so the 32 lanes of a warp stay together; everything that prints, cuts or probes is the
// rare slow path.  ncu (profiles/r02_bytewise.txt): the generic automaton ran 116 instructions per byte and lane with
// 8.8 of 32 lanes active, this one 22 lanes and 1.7x fewer warp instructions.
// tests/emul runs plain missions of these decoders through it on the CPU, against the oracle.
#pragma once
#include "sx_fast_utf8.cuh"

namespace sx {

enum : uint32_t { GE_CHAR = 1, GE_MAL = 2 };
// The decoders' step() has one return per case; inlined, the compiler clones everything that follows into each case
// (ncu: 8.8 of 32 lanes active through the whole automaton).  Passing the recorded events through an empty asm makes
// them opaque at the join, so the automaton below is emitted once and the lanes meet again in front of it.
#define SX_OPAQUE(x) asm volatile("" : "+r"(x))
struct GenEvent {
    uint32_t kind;  // 0 none, GE_CHAR, GE_MAL
    uint32_t lb, ul;
    int32_t cs, ce;  // char: input range (relative to the window start); mal: cs = start of the next segment
    uint32_t half2;  // char: second code point of a two-code-point pair (same input bytes as the first)
};
// Dec::step's emitter interface (ch / ch2 / cp / mal), recording up to two events
struct GenRecorder {
    int64_t ws;
    uint32_t n;
    GenEvent e0, e1;
    SX_HD void put(const GenEvent& e) { if (n == 0) e0 = e; else e1 = e; ++n; }
    SX_HD void cp(uint32_t) {}
    SX_HD void ch(uint32_t lb, uint32_t ul, int64_t cs, int64_t ce) { put(GenEvent{GE_CHAR, lb, ul, (int32_t)(cs - ws), (int32_t)(ce - ws), 0u}); }
    SX_HD void ch2(uint32_t lb, uint32_t ul, int64_t cs, int64_t ce) { put(GenEvent{GE_CHAR, lb, ul, (int32_t)(cs - ws), (int32_t)(ce - ws), 1u}); }
    SX_HD void mal(int64_t next) { put(GenEvent{GE_MAL, 0u, 0u, (int32_t)(next - ws), 0, 0u}); }
};

// ---- the decoders' step() in select form: one exit, the (at most two) events of a byte computed with selects ----
// Same transitions as Dec*::step in sx_core.cuh (which the general automaton keeps using; tests/emul runs plain missions
// through these and general missions through those, both against the oracle).  p = byte offset relative to the window.
SX_HD uint32_t gen_lead_of_cp(uint32_t c) {
    const uint32_t l2 = 0xC0u | (c >> 6), l3 = 0xE0u | (c >> 12), l4 = 0xF0u | (c >> 18);
    uint32_t r = c;
    r = c >= 0x80u ? l2 : r;
    r = c >= 0x800u ? l3 : r;
    r = c >= 0x10000u ? l4 : r;
    return r;
}
SX_HD uint32_t gen_len_of_cp(uint32_t c) { return 1u + (uint32_t)(c >= 0x80u) + (uint32_t)(c >= 0x800u) + (uint32_t)(c >= 0x10000u); }

template <class Dec> struct GenStep {  // any other decoder: its own step() against the recorder
    SX_HD static void run(Dec& d, const ScanParams& P, uint32_t b, int32_t p, int64_t ws, GenEvent& e0, GenEvent& e1) {
        GenRecorder rec;
        rec.ws = ws; rec.n = 0;
        rec.e0 = GenEvent{0u, 0u, 0u, 0, 0, 0u};
        rec.e1 = rec.e0;
        d.step(P, b, ws + p, rec);
        e0 = rec.e0; e1 = rec.e1;
    }
};
template <> struct GenStep<DecBig5> {
    SX_HD static void run(DecBig5& d, const ScanParams& P, uint32_t b, int32_t p, int64_t, GenEvent& e0, GenEvent& e1) {
        const uint32_t l = d.lead;
        const bool has = l != 0, ascii = b < 0x80u;
        const bool trail = has & big5_is_trail(b);
        const uint32_t ptr = trail ? big5_pointer(l, b) : 0u;
        const bool dbl = trail & big5_is_double(ptr);
        uint32_t cp = 0;
        if (trail & !dbl) cp = P.mb_a[ptr];
        const bool mapped = cp != 0;
        const bool ch_pair = dbl | mapped;                    // the pending lead and b make a char (or two)
        const bool fall = !has | (!ch_pair & ascii);           // b is (also) read in the neutral state
        const bool bad = (b == 0x80u) | (b == 0xFFu);
        d.lead = (fall & !ascii & !bad) ? b : 0u;
        // with a pending lead: char(s), or malformed -- an ASCII byte is not consumed and starts the next segment
        const uint32_t k0_has = ch_pair ? GE_CHAR : GE_MAL;
        const uint32_t k0_neu = ascii ? GE_CHAR : (bad ? GE_MAL : 0u);
        e0.kind = has ? k0_has : k0_neu;
        e0.lb = has ? (dbl ? 0xC3u : gen_lead_of_cp(cp)) : b;  // U+00CA / U+00EA first
        e0.ul = has ? (dbl ? 2u : gen_len_of_cp(cp)) : 1u;
        e0.cs = has ? (ch_pair ? p - 1 : (ascii ? p : p + 1)) : (ascii ? p : p + 1);
        e0.ce = p + 1;
        e0.half2 = 0;
        e1.kind = (has & (dbl | (!ch_pair & ascii))) ? GE_CHAR : 0u;
        e1.lb = dbl ? 0xCCu : b;  // U+0304 / U+030C second
        e1.ul = dbl ? 2u : 1u;
        e1.cs = dbl ? p - 1 : p;
        e1.ce = p + 1;
        e1.half2 = dbl ? 1u : 0u;
    }
};
template <> struct GenStep<DecEucJp> {
    SX_HD static void run(DecEucJp& d, const ScanParams& P, uint32_t b, int32_t p, int64_t, GenEvent& e0, GenEvent& e1) {
        const uint32_t l = d.lead, j = d.j0212;
        const bool ascii = b < 0x80u, a1fe = (b - 0xA1u) <= 0x5Du, runb = eucjp_is_run_byte(b);
        const bool caseA = (l == 0x8Eu) & (j == 0) & (b >= 0xA1u) & (b <= 0xDFu);  // half-width katakana U+FF61..U+FF9F
        const bool caseB = (l == 0x8Fu) & (j == 0) & a1fe;                         // JIS X 0212 lead
        const bool caseC = (l != 0) & !caseA & !caseB;
        const bool pair = caseC & ((l - 0xA1u) <= 0x5Du) & a1fe;
        uint32_t cp = 0;
        if (pair) cp = (j ? P.mb_b : P.mb_a)[(l - 0xA1u) * 94u + (b - 0xA1u)];
        const bool mapped = cp != 0;
        const bool fall = (l == 0) | (caseC & !mapped & ascii);
        d.lead = caseB ? b : ((fall & runb) ? b : 0u);
        d.j0212 = caseB ? 1u : 0u;
        const uint32_t kC = mapped ? GE_CHAR : GE_MAL;
        const uint32_t kN = ascii ? GE_CHAR : (runb ? 0u : GE_MAL);
        e0.kind = caseA ? GE_CHAR : (caseB ? 0u : (caseC ? kC : kN));
        e0.lb = caseA ? 0xEFu : (caseC ? gen_lead_of_cp(cp) : b);
        e0.ul = caseA ? 3u : (caseC ? gen_len_of_cp(cp) : 1u);
        const int32_t csC = mapped ? p - (j ? 2 : 1) : (ascii ? p : p + 1);
        e0.cs = caseA ? p - 1 : (caseC ? csC : (ascii ? p : p + 1));
        e0.ce = p + 1;
        e0.half2 = 0;
        e1.kind = (caseC & !mapped & ascii) ? GE_CHAR : 0u;
        e1.lb = b; e1.ul = 1u; e1.cs = p; e1.ce = p + 1; e1.half2 = 0;
    }
};
template <bool BE> struct GenStep<DecUtf16<BE>> {
    SX_HD static void run(DecUtf16<BE>& d, const ScanParams&, uint32_t b, int32_t p, int64_t, GenEvent& e0, GenEvent& e1) {
        e0 = GenEvent{0u, 0u, 0u, 0, p + 1, 0u};
        e1 = e0;
        if (!d.has_lead) { d.has_lead = 1; d.lead_byte = b; return; }  // same parity in every lane of a warp (W is even)
        d.has_lead = 0;
        const uint32_t cu = DecUtf16<BE>::unit(d.lead_byte, b), hb = cu & 0xFC00u, ls = d.lead_sur;
        const bool isH = hb == 0xD800u, isL = hb == 0xDC00u, bmp = !isH & !isL, pend = ls != 0;
        d.lead_sur = isH ? cu : 0u;
        const uint32_t cp4 = 0x10000u + ((ls - 0xD800u) << 10) + (cu - 0xDC00u);
        // high surrogate: malformed when one is pending already; low: completes the pair or is malformed; BMP unit after
        // a pending high surrogate: Malformed(2,2), then the unit itself first thing in the next segment
        e0.kind = isH ? (pend ? GE_MAL : 0u) : (isL ? (pend ? GE_CHAR : GE_MAL) : (pend ? GE_MAL : GE_CHAR));
        const bool c4 = isL & pend, cb = bmp & !pend;
        e0.lb = c4 ? (0xF0u | (cp4 >> 18)) : gen_lead_of_cp(cu);
        e0.ul = c4 ? 4u : gen_len_of_cp(cu);
        e0.cs = c4 ? p - 3 : (cb ? p - 1 : p + 1);
        e1.kind = (bmp & pend) ? GE_CHAR : 0u;
        e1.lb = gen_lead_of_cp(cu); e1.ul = gen_len_of_cp(cu); e1.cs = p - 1;
    }
};

struct GenEmit {
    const ScanParams* P;
    int64_t base;
    int mode;
    Record* wr;
    uint64_t text_off;
    uint32_t nrec, ntext;
};
SX_HD_NOINLINE void gen_emit(GenEmit* E, int32_t seg_rel, uint32_t prec, int32_t run_s, int32_t run_e, uint32_t text_len, uint32_t flags) {
    if (E->mode == MODE_WRITE || (E->mode == MODE_BUFFER && E->nrec < kBufRecs)) {
        Record r;
        r.position = E->P->base_consumed + (uint64_t)(E->base + seg_rel);  // finding_collection.rs:260
        r.in_start = E->base + run_s;
        r.in_len = (uint32_t)(run_e - run_s);
        r.text_len = text_len;
        r.text_off = E->text_off;
        r.flags = flags;
        r.precision = prec;
        *E->wr++ = r;
        E->text_off += text_len;
    }
    E->nrec++;
    E->ntext += text_len;
}

enum : uint32_t { GF_HALF = 1u << 11, GF_LEFTHALF = 1u << 12 };  // on top of the FF_* bits of sx_fast_utf8.cuh

template <class Dec, class TileSrc>
SX_HD_NOINLINE void scan_window_fast_generic(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo,
                                             const Carry& kin, int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
    const uint32_t n = P.n, q = P.q;
    // pass_filter (sx_core.cuh; Utf8Filter::pass_af_filter / pass_ubf_filter, mission.rs:333-348) as one indexed word:
    // lead bytes 00-7F index the 128-bit ASCII filter, 80-FF the 64-bit unicode-block filter with (lb & 0x3f)
    const uint32_t ftab[8] = {(uint32_t)P.af_lo, (uint32_t)(P.af_lo >> 32), (uint32_t)P.af_hi, (uint32_t)(P.af_hi >> 32),
                              (uint32_t)P.ubf,   (uint32_t)(P.ubf >> 32),   (uint32_t)P.ubf,   (uint32_t)(P.ubf >> 32)};
    Dec dec;
    dec.init(P, tsrc, geo.ws);
    ProbeCtx<Dec> pc;
    pc.P = &P; pc.g = &g; pc.geo = &geo; pc.pend0 = 0; pc.s0 = pc.s1 = pc.s2 = 0;
    probe_capture(pc, dec);
    const int32_t pend0 = dec.pending_len();
    GenEmit E;
    E.P = &P; E.base = geo.ws; E.mode = mode; E.wr = wr; E.text_off = text_off; E.nrec = 0; E.ntext = 0;
    const int32_t slice_rel = (int32_t)(geo.slice_start - geo.ws);
    const int32_t wlen = (int32_t)(geo.we - geo.ws);
    const Carry slice_left = kin;  // only read by the probe when the window starts a slice

    // ---- first segment (finding_collection.rs:101-116, :211-241) ----
    uint32_t fl = FF_ATLEFT | FF_INFIRST | FF_S1ALL | FF_S2ALL;
    if (kin.kind == K_C) fl |= FF_LASTCUT;
    if (slice_rel == 0 && Dec::kStateful) fl |= FF_PROBE;
    uint32_t m = 1, a = 0, run_n = 0, run_t = 0, prec = PREC_EXACT;
    int32_t seg_rel = 0, run_s = 0, run_e = 0;
    uint32_t left_k = 0, left_t = 0;
    int32_t left_s = 0;
    if (kin.kind == K_L && kin.k > 0) {
        run_n = kin.k;
        run_t = kin.out_bytes;
        run_s = -(int32_t)kin.in_bytes;
        run_e = -pend0;
        if (kin.flags & CF_HOSTCARRY) fl |= FF_HOSTCARRY;
        if (kin.flags & CF_HALF) fl |= GF_HALF;
        prec = PREC_BEFORE;
    }

#define SXG_COMPLETES ((fl & (FF_ATLEFT | FF_LASTCUT)) == (FF_ATLEFT | FF_LASTCUT))
#define SXG_YIELD(completes_, maybe_cut_)                                                                        \
    do {                                                                                                         \
        if (m == 1 && !(fl & FF_INFIRST)) fl |= FF_S1LATER;                                                      \
        gen_emit(&E, seg_rel, prec, run_s, run_e, run_t,                                                         \
                 ((completes_) ? (uint32_t)RF_COMPLETES : 0u) | ((fl & FF_HOSTCARRY) ? (uint32_t)RF_HOSTCARRY : 0u) | \
                     ((fl & GF_HALF) ? (uint32_t)RF_HALFSTART : 0u));                                             \
        fl = (maybe_cut_) ? (fl | FF_CUT) : (fl & ~FF_CUT);                                                      \
        fl &= ~FF_HASLEFT;                                                                                       \
        prec = PREC_AFTER;                                                                                       \
    } while (0)

    // One decoder event through the automaton.  The common transitions are written as blends under all-ones / all-zero
    // masks that are opaque to the compiler: written with bools and ifs it split the lambda into one copy per kind of
    // event again (ncu: 10 of 32 lanes active), this way they are straight-line code for the whole warp.
#define SXG_BLEND(x, y, k) ((x) ^ (((x) ^ (y)) & (k)))
#define SXG_MASK(cond) (0u - (uint32_t)(cond))
    auto feed = [&](const GenEvent& ev) {
        uint32_t k_char = SXG_MASK(ev.kind == GE_CHAR), k_mal = SXG_MASK(ev.kind == GE_MAL);
        const uint32_t k_pf = 0u - ((ftab[(ev.lb >> 5) & 7u] >> (ev.lb & 31u)) & 1u);
        uint32_t k_pass = k_char & k_pf, k_fail = k_char & ~k_pf;
        const uint32_t k_ends = k_mal | k_fail;
        const uint32_t k_print = SXG_MASK(run_n >= n) | SXG_MASK((fl & (FF_ATLEFT | FF_LASTCUT)) == (FF_ATLEFT | FF_LASTCUT));
        uint32_t k_slow = (k_ends & SXG_MASK(run_n > 0) & k_print) | (k_pass & SXG_MASK(run_n + 1 >= q)) |
                          (k_char & SXG_MASK((fl & FF_PROBE) != 0) & SXG_MASK(ev.lb >= 0x80));
        SX_OPAQUE(k_slow); SX_OPAQUE(k_char); SX_OPAQUE(k_mal); SX_OPAQUE(k_pass); SX_OPAQUE(k_fail);
        if (k_slow == 0) {
            // malformed sequence: new segment, nothing to print
            m -= k_mal;
            const uint32_t fl_mal = (fl & ~(FF_LASTCUT | FF_INFIRST | FF_HOSTCARRY | FF_HASLEFT | FF_PROBE | FF_CUT | GF_HALF)) | FF_ATLEFT |
                                    ((fl & FF_CUT) ? FF_LASTCUT : 0u) | ((ev.cs == slice_rel && Dec::kStateful) ? FF_PROBE : 0u);
            fl = SXG_BLEND(fl, fl_mal, k_mal);
            prec = SXG_BLEND(prec, (uint32_t)PREC_EXACT, k_mal);
            seg_rel = (int32_t)SXG_BLEND((uint32_t)seg_rel, (uint32_t)ev.cs, k_mal);
            fl &= ~(FF_PROBE & k_char);  // only the first char of a segment can trigger the probe (finding_collection.rs:176)
            // passing char
            const uint32_t k_first = k_pass & SXG_MASK(run_n == 0);
            run_s = (int32_t)SXG_BLEND((uint32_t)run_s, (uint32_t)ev.cs, k_first);
            fl = SXG_BLEND(fl, (fl & ~GF_HALF) | (ev.half2 ? GF_HALF : 0u), k_first);
            a += ((fl & FF_INFIRST) ? 1u : 0u) & k_pass;
            run_e = (int32_t)SXG_BLEND((uint32_t)run_e, (uint32_t)ev.ce, k_pass);
            // failing char
            const uint32_t clr = (m == 1 ? FF_S1ALL : 0u) | (m == 2 ? FF_S2ALL : 0u) | FF_INFIRST | FF_HOSTCARRY | FF_ATLEFT | GF_HALF;
            fl &= ~(clr & k_fail);
            run_n = (run_n - k_pass) & ~k_ends;
            run_t = (run_t + (ev.ul & k_pass)) & ~k_ends;
            return;
        }
        const bool mal = k_mal != 0, pass = k_pass != 0;
        if (mal) {  // the segment ends (invalid_after), the next one starts at ev.cs
            if (run_n > 0 && (run_n >= n || SXG_COMPLETES)) { const bool c_ = SXG_COMPLETES; SXG_YIELD(c_, false); }
            m++;
            fl = (fl & ~(FF_LASTCUT | FF_INFIRST | FF_HOSTCARRY | FF_HASLEFT | FF_PROBE | FF_CUT | GF_HALF)) | FF_ATLEFT |
                 ((fl & FF_CUT) ? FF_LASTCUT : 0u) | ((ev.cs == slice_rel && Dec::kStateful) ? FF_PROBE : 0u);
            run_n = 0; run_t = 0;
            prec = PREC_EXACT;
            seg_rel = ev.cs;
            return;
        }
        if (fl & FF_PROBE) {
            fl &= ~FF_PROBE;
            if (ev.lb >= 0x80 && mode != MODE_STATE && ProbeImpl<Dec>::run(pc, m == 1, slice_left)) prec = PREC_BEFORE;
        }
        if (pass) {
            if (run_n == 0) { run_s = ev.cs; fl = ev.half2 ? (fl | GF_HALF) : (fl & ~GF_HALF); }
            run_n++;
            run_t += ev.ul;
            run_e = ev.ce;
            if (fl & FF_INFIRST) a++;
            if (run_n >= q) {  // helper.rs:237 exit 2, :353-355, :418-421
                const bool c_ = SXG_COMPLETES;
                SXG_YIELD(c_, true);
                fl = (fl | FF_ATLEFT | FF_LASTCUT) & ~(FF_HOSTCARRY | GF_HALF);
                run_n = 0; run_t = 0;
            }
        } else {
            if (m == 1) fl &= ~FF_S1ALL;
            if (m == 2) fl &= ~FF_S2ALL;
            if (run_n > 0 && (run_n >= n || SXG_COMPLETES)) {  // helper.rs:315-322
                const bool c_ = SXG_COMPLETES;
                SXG_YIELD(c_, false);
                fl &= ~FF_LASTCUT;
            }
            fl &= ~(FF_INFIRST | FF_HOSTCARRY | FF_ATLEFT | GF_HALF);
            run_n = 0; run_t = 0;
        }
    };

    tsrc.for_each_byte(geo.ws, geo.we, [&](uint32_t b, int64_t pos) {
        GenEvent e0, e1;
        GenStep<Dec>::run(dec, P, b, (int32_t)(pos - geo.ws), geo.ws, e0, e1);
        SX_OPAQUE(e0.kind); SX_OPAQUE(e0.lb); SX_OPAQUE(e0.ul); SX_OPAQUE(e0.cs); SX_OPAQUE(e0.ce); SX_OPAQUE(e0.half2);
        SX_OPAQUE(e1.kind); SX_OPAQUE(e1.lb); SX_OPAQUE(e1.ul); SX_OPAQUE(e1.cs); SX_OPAQUE(e1.ce); SX_OPAQUE(e1.half2);
        feed(e0);  // no event: every mask is zero
        if (e1.kind) {
            // Second event of a byte: a char right behind a malformed sequence (the ASCII byte that ended a double-byte
            // sequence, the BMP unit behind a lone surrogate).  The automaton has just started a segment, so the
            // transitions are a few blends; the probe and q == 1 take the general path.
            const uint32_t k_pf = 0u - ((ftab[(e1.lb >> 5) & 7u] >> (e1.lb & 31u)) & 1u);
            const bool fresh = e0.kind == GE_MAL && run_n == 0;
            if (fresh && q > 1 && !((fl & FF_PROBE) && e1.lb >= 0x80)) {
                fl &= ~FF_PROBE;
                run_s = (int32_t)SXG_BLEND((uint32_t)run_s, (uint32_t)e1.cs, k_pf);
                run_e = (int32_t)SXG_BLEND((uint32_t)run_e, (uint32_t)e1.ce, k_pf);
                fl = SXG_BLEND(fl, (fl & ~GF_HALF) | (e1.half2 ? GF_HALF : 0u), k_pf);
                run_n = 1u & k_pf;
                run_t = e1.ul & k_pf;
                const uint32_t clr = (m == 2 ? FF_S2ALL : 0u) | FF_HOSTCARRY | FF_ATLEFT | GF_HALF;  // m >= 2, not in the first run
                fl &= ~(clr & ~k_pf);
            } else feed(e1);
        }
    });

    // ---- end of the window's last segment (helper.rs:343-431 for the run touching the right boundary) ----
    const bool invalid_after = geo.final_last;
    if (run_n > 0) {
        const bool completes = SXG_COMPLETES;
        if (!completes && !invalid_after) {  // `again`: kept as leftover (finding_collection.rs:281-284)
            if (m == 1 && !(fl & FF_INFIRST)) fl |= FF_S1LATER;
            fl |= FF_HASLEFT;
            fl = (fl & FF_HOSTCARRY) ? (fl | FF_LEFTHC) : (fl & ~FF_LEFTHC);
            fl = (fl & GF_HALF) ? (fl | GF_LEFTHALF) : (fl & ~GF_LEFTHALF);
            left_k = run_n; left_s = run_s; left_t = run_t;
            fl &= ~FF_CUT;
        } else if (completes || run_n >= n) {
            SXG_YIELD(completes, !invalid_after);
        }
    }
#undef SXG_YIELD
#undef SXG_BLEND
#undef SXG_MASK
#undef SXG_COMPLETES
    if (geo.final_last) {
        // finding_collection.rs:298-304: one flush round; lasting effects: reset decoder, cut == false
        fl &= ~(FF_CUT | FF_HASLEFT);
        res.npend_out = 0;
    } else {
        res.npend_out = dec.pending_len();
    }
    if (fl & FF_CUT) res.out = carry_cut();
    else if (fl & FF_HASLEFT) {
        Carry c;
        c.kind = K_L;
        c.flags = (uint8_t)(((fl & FF_LEFTHC) ? CF_HOSTCARRY : 0) | ((fl & GF_LEFTHALF) ? CF_HALF : 0));
        c.k = (uint16_t)left_k;
        c.in_bytes = (uint32_t)(wlen - left_s);
        c.out_bytes = left_t;
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
            if (a < q) { desc->type = WT_CASEB; desc->t_out = (uint16_t)((fl & FF_HASLEFT) ? left_t : 0u); }
            else desc->type = WT_CONST;
        } else if (a > 0 && !(fl & FF_S1LATER) && (m == 1 || (m == 2 && (fl & FF_S2ALL)))) desc->type = WT_DEP;
        else desc->type = WT_CONST;
    }
}

}  // namespace sx
