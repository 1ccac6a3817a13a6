// sx_core.cuh -- per-window scanner automaton shared by the CUDA kernels (sx_scan.cu, sx_exact.cuh,
// sx_sparse_utf8.cuh) and by the CPU test harness (tests/emul).
//
// One "window" is one decoder_input_window of the reference (2 * output_line_char_nb_max input
// bytes, /root/reference/src/finding_collection.rs:120-131).  The reference walks windows
// strictly sequentially; here every window is processed independently by one GPU lane:
//
//   * the decoder state at the window start is re-derived from the preceding bytes
//     (UTF-8 / UTF-16 are self-synchronising under the WHATWG decoders, see Dec*::init; the
//     lead/trail encodings Big5 / EUC-JP walk back to the last byte that cannot be a trail),
//   * the scanner carry between windows (`Carry`: leftover run or "cut" flag, i.e.
//     ScannerState.last_scan_run_leftover / last_run_str_was_printed_and_is_maybe_cut_str,
//     /root/reference/src/scanner.rs:45-68) is resolved along runs of adjacent listed windows by
//     the exact stage (sx_exact.cuh block kernel / sx_sparse_utf8.cuh per-stage pipeline),
//   * the SplitStr iterator (/root/reference/src/helper.rs:210-432) and the chunk loop of
//     FindingCollection::from (finding_collection.rs:246-290) are restated as ONE streaming
//     automaton over decoder events (`WinAuto`), so no decoded text is ever materialised.
//
// WinAuto is the GENERAL automaton: it also implements grep_char (helper.rs:215-331 incl. the
// early-termination quirk), require_same_unicode_block (helper.rs:287-292) and
// chars_min_nb > output_line_char_nb_max; the C ABI only rejects chars_min_nb == 0 and
// output_line_char_nb_max outside 6..8192.
//
// Everything is SX_HD so the host-side *test* harness (tests/emul/) can run the same code on
// the CPU to debug the decomposition without a GPU.  The product never does that.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SX_HD __host__ __device__ __forceinline__
#define SX_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define SX_HD inline
#define SX_HD_NOINLINE inline
#endif

namespace sx {

enum : uint32_t {
    ENC_XUD = 0,      // x-user-defined (also the `ascii` emulation, mission.rs:623-679)
    ENC_UTF8 = 1,
    ENC_UTF16LE = 2,
    ENC_UTF16BE = 3,
    ENC_SB = 4,       // single-byte, table driven
    ENC_UTF32LE = 5,  // extension (no reference semantics)
    ENC_UTF32BE = 6,
    ENC_BIG5 = 7,     // WHATWG Big5 (lead 81-FE, trail 40-7E | A1-FE), table: sx_mb_tables.inc
    ENC_EUCJP = 8     // WHATWG EUC-JP (jis0208, 8E kana, 8F + jis0212)
};

enum : uint8_t { PREC_BEFORE = 0, PREC_EXACT = 1, PREC_AFTER = 2 };  // finding.rs:34-46

// Carry between segments / windows / slices.
enum : uint8_t { K_L = 0, K_C = 1, K_UNKNOWN = 2 };
enum : uint8_t {
    CF_HOSTCARRY = 1,  // leftover starts with text carried in the host ScannerState
    CF_GREP = 2,       // leftover contains the mission's grep_char (helper.rs:252-254)
    CF_HALF = 4        // leftover starts with the SECOND code point of a Big5 two-code-point pair (its first input bytes
                       // are the pair; the pair's first code point is not part of the text)
};
struct Carry {
    uint8_t kind;        // K_L: leftover of k chars (k may be 0) and cut == false; K_C: cut == true
    uint8_t flags;
    uint16_t k;          // chars in the leftover
    uint32_t in_bytes;   // distance from the leftover's first input byte to the boundary
    uint32_t out_bytes;  // UTF-8 bytes of the leftover (device-decoded part only)
    uint32_t aux;        // UTF-8 lead byte of the leftover's last multi-byte char (same-unicode-block rule), else 0
};
SX_HD Carry carry_none() { return Carry{K_L, 0, 0, 0, 0, 0}; }
SX_HD Carry carry_cut() { return Carry{K_C, 0, 0, 0, 0, 0}; }
SX_HD Carry carry_unknown() { return Carry{K_UNKNOWN, 0, 0, 0, 0, 0}; }
SX_HD bool carry_is_null(const Carry& c) { return c.kind == K_L && c.k == 0; }

// One finding as it leaves the GPU (before text materialisation).
enum : uint32_t { RF_COMPLETES = 1, RF_HOSTCARRY = 2, RF_LEFTOVER = 4,
                  RF_HALFSTART = 8 /* the text starts with the second code point of the pair at in_start (Big5) */ };
struct Record {
    uint64_t position;  // Finding.position (absolute, includes counter_offset)   finding.rs:61
    int64_t in_start;   // first input byte of the text, buffer-relative (may reach into the virtual prefix)
    uint64_t text_off;  // offset of the UTF-8 text in the text arena
    uint32_t in_len;    // contiguous input range holding exactly the chars of the text
    uint32_t text_len;  // UTF-8 bytes (device-decoded part)
    uint32_t flags;     // RF_*
    uint32_t precision; // PREC_*
};

struct ScanParams {
    const uint8_t* in;  // device pointer to stream offset 0 of this call
    int64_t len;
    uint32_t slice_len, W, q, n;
    uint32_t enc;
    uint32_t align;      // UTF-16/32: units start at offsets o with (o - align) % unit == 0
    uint64_t af_lo, af_hi, ubf;
    uint64_t base_consumed;  // ScannerState.consumed_bytes at offset 0
    int32_t npend;           // virtual prefix: raw bytes still inside the decoder at offset 0
    int32_t is_last;
    uint8_t pend[8];         // pend[8-npend .. 8) are the bytes at offsets [-npend, 0)
    uint8_t carry_text8[8];  // first bytes of the host-carried leftover text (Precision::Before probe)
    uint32_t carry_text_len;
    Carry k0;                // carry at offset 0
    int32_t grep_char;       // Utf8Filter.grep_char, -1 = None
    uint32_t same_block;     // Mission.require_same_unicode_block
    uint32_t general;        // grep_char / same_block / chars_min_nb > q: general automaton + dual-simulation classification
    uint16_t sb_table[128];
    // Big5 / EUC-JP index tables (device memory in the kernels, host memory in the test harness): mb_a = Big5 index
    // or jis0208, mb_b = jis0212
    const uint32_t* mb_a;
    const uint32_t* mb_b;
    uint32_t* mb_fail;  // set to 1 when a decoder-state walk ran into its bound (the call then fails loudly)
};

// Byte source over the whole stream incl. the virtual prefix.  Used on the rare paths
// (look-back across tiles, probe, text materialisation); the hot loop reads shared memory.
struct GlobalSrc {
    const uint8_t* in;
    const uint8_t* pend;  // 8 bytes
    SX_HD uint8_t get(int64_t off) const { return off >= 0 ? in[off] : pend[8 + off]; }
};

SX_HD bool pass_filter(const ScanParams& P, uint32_t lb) {
    // Utf8Filter::pass_af_filter / pass_ubf_filter, mission.rs:333-348
    if (lb < 0x80) return ((lb < 64 ? (P.af_lo >> lb) : (P.af_hi >> (lb - 64))) & 1) != 0;
    return ((P.ubf >> (lb & 0x3f)) & 1) != 0;
}
SX_HD uint32_t utf8_len_of_cp(uint32_t c) { return c < 0x80 ? 1 : c < 0x800 ? 2 : c < 0x10000 ? 3 : 4; }
SX_HD uint32_t utf8_lead_of_cp(uint32_t c) {
    return c < 0x80 ? c : c < 0x800 ? (0xC0 | (c >> 6)) : c < 0x10000 ? (0xE0 | (c >> 12)) : (0xF0 | (c >> 18));
}
SX_HD uint32_t put_utf8(uint8_t* d, uint32_t c) {
    if (c < 0x80) { d[0] = (uint8_t)c; return 1; }
    if (c < 0x800) { d[0] = (uint8_t)(0xC0 | (c >> 6)); d[1] = (uint8_t)(0x80 | (c & 0x3F)); return 2; }
    if (c < 0x10000) {
        d[0] = (uint8_t)(0xE0 | (c >> 12)); d[1] = (uint8_t)(0x80 | ((c >> 6) & 0x3F));
        d[2] = (uint8_t)(0x80 | (c & 0x3F)); return 3;
    }
    d[0] = (uint8_t)(0xF0 | (c >> 18)); d[1] = (uint8_t)(0x80 | ((c >> 12) & 0x3F));
    d[2] = (uint8_t)(0x80 | ((c >> 6) & 0x3F)); d[3] = (uint8_t)(0x80 | (c & 0x3F)); return 4;
}

// ------------------------------------------------------------------------------------------
// Window summary (stage A) -> transfer-function descriptor used by the tile-level resolve.
// ------------------------------------------------------------------------------------------
enum : uint8_t { WT_CONST = 0, WT_CASEB = 1, WT_DEP = 2, WT_GUARD = 3 };
struct WinDesc {
    uint8_t type;        // WT_*
    uint8_t pad;
    uint16_t a;          // chars of the run touching the window's left boundary (capped)
    uint16_t t_out;      // WT_CASEB: UTF-8 bytes of the window's text
    uint16_t nrec;       // yields under the null carry
    uint32_t ntext;      // UTF-8 bytes of those yields
    Carry null_out;      // carry out under the null carry
};

struct WinResult {
    Carry in;           // mask engine: the carry-in it worked under (given, or derived from the 32 bytes before the window)
    Carry out;
    uint32_t nrec;
    uint32_t ntext;
    int32_t npend_out;  // bytes still inside the decoder at the window end
    uint32_t m;         // segments in the window
    uint32_t cut1;      // carry flag handed from segment 1 to segment 2
    // mask engine: caseb != 0: the window is ONE run of `a` (< q) passing chars with `t_out` text bytes covering it
    // completely, so its carry-out follows eval_caseb() from any carry-in (WinDesc WT_CASEB without a pass)
    uint16_t caseb, a, t_out;
};

// MODE_BUFFER: count, and also write the first kBufRecs records (text_off relative to the window's first
// record) into a small per-lane staging buffer, so most windows need no second (write) pass.
enum : int { MODE_STATE = 0, MODE_COUNT = 1, MODE_WRITE = 2, MODE_BUFFER = 3 };
constexpr uint32_t kBufRecs = 2;

// ------------------------------------------------------------------------------------------
// The streaming SplitStr + chunk-loop automaton.
// ------------------------------------------------------------------------------------------
struct WinAuto {
    const ScanParams* P;
    int mode;
    Record* wr;        // MODE_WRITE: next record slot
    uint64_t text_off; // MODE_WRITE: next text offset
    int64_t slice_start;
    bool probe_possible;  // stateful decoder: the Precision::Before probe can fire
    int force_s2_cont;    // classification aid: -1, or the `cont` value forced at the start of segment 2

    // segment / SplitStr state
    int64_t seg_pos;
    uint8_t prec;
    bool probe_pending;
    bool last_cut;   // SplitStr.last_s_was_maybe_cut
    bool at_left;    // ok_s_p == inp_start_p for the current run
    bool cut;        // last_window_str_was_printed_and_is_maybe_cut_str
    bool run_hostcarry;
    bool grep_ok;    // helper.rs:215: current run holds the grep_char (always true without one)
    bool qfull;      // the run reached q chars; whether it touches the right boundary is known at the next event
    bool dead;       // SplitStr::next returned None (helper.rs:410-415): the rest of the segment is ignored
    bool next_half2;   // the char about to be reported is the second code point of a two-code-point pair (Big5)
    bool run_half;     // the current run starts with such a char
    bool left_half;
    uint32_t last_mb;  // helper.rs:221: lead byte of the last multi-byte char this SplitStr::next() call saw (may be stale)
    uint32_t run_mb;   // ... of the current run only: what a re-scan of the run as leftover will see
    uint32_t run_n, run_out;
    int64_t run_in_start, run_in_end;
    // leftover produced by an `again` chunk
    bool has_left, left_hostcarry, left_grep;
    uint32_t left_k, left_out, left_mb;
    int64_t left_in_start;
    // probe context: leftover present at the slice start (its text may sit at out[0..])
    bool slice_left_present;
    Carry slice_left;

    // outputs
    uint32_t nrec, ntext;

    // summary instrumentation
    uint32_t m;             // segments so far
    uint32_t a;             // chars in the first run of segment 1
    uint32_t s1_out;        // UTF-8 bytes of passing chars in segment 1
    bool in_first_run;      // still inside the first run of segment 1
    bool s1_all_pass, s1_later_yield, s2_all_pass;
    bool cut1;              // carry flag handed from segment 1 to segment 2 (classification)
    bool first_ended;       // the first run of segment 1 ended inside the segment ...
    bool first_clean;       // ... in a way that leaves no stale lead byte behind (general missions, WT_GUARD)

    SX_HD void init(const ScanParams* p, int md, int64_t slice_st, bool probe_ok) {
        P = p; mode = md; wr = nullptr; text_off = 0; slice_start = slice_st; probe_possible = probe_ok;
        force_s2_cont = -1; cut1 = false;
        cut = false; has_left = false; left_hostcarry = false; left_grep = false; left_k = left_out = left_mb = 0; left_in_start = 0;
        nrec = ntext = 0; m = 0; a = 0; s1_out = 0; in_first_run = false;
        s1_all_pass = true; s1_later_yield = false; s2_all_pass = true; first_ended = false; first_clean = false;
        slice_left_present = false; slice_left = carry_none();
        run_n = run_out = 0; run_in_start = run_in_end = 0; run_hostcarry = false;
        grep_ok = true; qfull = false; dead = false; last_mb = 0; run_mb = 0;
        seg_pos = 0; prec = PREC_EXACT; probe_pending = false; last_cut = false; at_left = true;
        next_half2 = false; run_half = false; left_half = false;
    }
    SX_HD bool grep_reset() const { return P->grep_char < 0; }  // helper.rs:215 / :330

    // finding_collection.rs:211-241 + helper.rs:171-200.  `kin` only for the first segment of a window.
    SX_HD void segment_start(int64_t pos, const Carry* kin, int32_t pend_len) {
        seg_pos = pos;
        m++;
        if (kin) cut = (kin->kind == K_C);
        if (m == 2) {
            cut1 = cut;
            if (force_s2_cont >= 0) cut = force_s2_cont != 0;
        }
        const bool cont = cut;  // finding_collection.rs:240-241: used once
        cut = false;
        last_cut = cont;
        at_left = true;
        run_n = 0; run_out = 0; run_hostcarry = false; run_half = false;
        grep_ok = grep_reset(); last_mb = 0; run_mb = 0; qfull = false; dead = false;
        prec = PREC_EXACT;
        probe_pending = probe_possible && (pos == slice_start);
        has_left = false;
        if (kin && kin->kind == K_L && kin->k > 0) {  // finding_collection.rs:214-221: the leftover is re-scanned
            run_n = kin->k;
            run_out = kin->out_bytes;
            run_in_start = pos - (int64_t)kin->in_bytes;
            run_in_end = pos - pend_len;
            run_hostcarry = (kin->flags & CF_HOSTCARRY) != 0;
            run_half = (kin->flags & CF_HALF) != 0;
            grep_ok = grep_reset() || (kin->flags & CF_GREP) != 0;
            last_mb = kin->aux;
            run_mb = kin->aux;
            qfull = run_n >= P->q;  // only possible with a grep_char (helper.rs:389-392)
            prec = PREC_BEFORE;
        }
        if (m == 1) in_first_run = true;
    }

    SX_HD void yield(bool completes, bool maybe_cut) {
        if (m == 1 && !in_first_run) s1_later_yield = true;
        if (mode == MODE_WRITE || (mode == MODE_BUFFER && nrec < kBufRecs)) {
            Record r;
            r.position = P->base_consumed + (uint64_t)seg_pos;  // finding_collection.rs:260
            r.in_start = run_in_start;
            r.in_len = (uint32_t)(run_in_end - run_in_start);
            r.text_len = run_out;
            r.text_off = text_off;
            r.flags = (completes ? RF_COMPLETES : 0u) | (run_hostcarry ? RF_HOSTCARRY : 0u) | (run_half ? (uint32_t)RF_HALFSTART : 0u);
            r.precision = prec;
            *wr++ = r;
            text_off += run_out;
        }
        nrec++;
        ntext += run_out;
        cut = maybe_cut;     // finding_collection.rs:268
        prec = PREC_AFTER;   // finding_collection.rs:289
        has_left = false;
    }
    SX_HD void keep_leftover() {  // an `again` chunk, finding_collection.rs:281-284
        if (m == 1 && !in_first_run) s1_later_yield = true;
        has_left = true;
        left_k = run_n; left_out = run_out; left_in_start = run_in_start; left_hostcarry = run_hostcarry; left_half = run_half;
        left_grep = grep_ok && P->grep_char >= 0;
        left_mb = P->same_block ? run_mb : 0u;  // only the same-unicode-block rule reads it; the leftover is re-scanned from scratch
        cut = false;
        prec = PREC_AFTER;
    }
    SX_HD void new_run() {  // a fresh SplitStr::next() call
        run_n = 0; run_out = 0; run_hostcarry = false; run_half = false;
        grep_ok = grep_reset();
        last_mb = 0;
        run_mb = 0;
    }
    // The run holds q chars (helper.rs:237 loop exit 2): evaluate the chunk now that `right` is known.
    SX_HD void resolve_qfull(bool right, bool invalid_after) {
        qfull = false;
        const bool completes = at_left && last_cut;                               // helper.rs:355
        const bool again = !completes && right && !invalid_after && !grep_ok;     // helper.rs:389-392 (ok_n == q)
        if (again) { keep_leftover(); run_n = 0; run_out = 0; return; }
        if (!completes && (!grep_ok || run_n < P->n)) { dead = true; run_n = 0; run_out = 0; return; }  // helper.rs:410-415
        yield(completes, true);                                                   // helper.rs:353
        at_left = true;   // helper.rs:418-420: inp_start_p = p
        last_cut = true;
        new_run();
    }

    template <class ProbeFn>
    SX_HD void on_char(uint32_t lb, uint32_t ul, int64_t cstart, int64_t cend, ProbeFn&& probe) {
        if (probe_pending) {  // finding_collection.rs:176: only the first char of the segment matters
            probe_pending = false;
            if (lb >= 0x80 && mode != MODE_STATE) {
                if (probe(m == 1, slice_left)) prec = PREC_BEFORE;
            }
        }
        if (qfull) resolve_qfull(false, false);
        if (dead) return;
        bool pass;
        if (lb < 0x80) {
            if (!grep_ok && P->grep_char == (int32_t)lb) grep_ok = true;  // helper.rs:252-254, before the filter
            pass = pass_filter(*P, lb);
        } else if (pass_filter(*P, lb)) {
            if (!P->same_block || lb == last_mb || last_mb == 0) { last_mb = lb; run_mb = lb; pass = true; }
            else {
                // helper.rs:287-292: a passing char of another block ends the run and is scanned again
                if (m == 1) s1_all_pass = false;
                if (m == 2) s2_all_pass = false;
                if (run_n > 0 && ((last_cut && at_left) || (run_n >= P->n && grep_ok))) {
                    yield(at_left && last_cut, false);
                    last_cut = false;
                }
                if (m == 1 && in_first_run) { first_ended = true; first_clean = true; }
                if (m == 1) in_first_run = false;
                at_left = false;
                new_run();
                last_mb = lb;
                run_mb = lb;
                pass = true;  // now the first char of a new run
            }
        } else {
            last_mb = 0;
            pass = false;
        }
        if (pass) {
            if (run_n == 0) { run_in_start = cstart; run_half = next_half2; }
            run_n++;
            run_out += ul;
            run_in_end = cend;
            if (m == 1) { s1_out += ul; if (in_first_run && a < 0xFFFFu) a++; }
            if (run_n >= P->q) qfull = true;  // decided at the next event (helper.rs:351: touches the right boundary?)
        } else {
            if (m == 1) { s1_all_pass = false; }
            if (m == 2) s2_all_pass = false;
            bool broke = false;
            if (run_n > 0) {
                // helper.rs:315-322 exit 3 / exit 4
                if ((last_cut && at_left) || (run_n >= P->n && grep_ok)) {
                    yield(at_left && last_cut, false);
                    last_cut = false;
                    broke = true;
                }
            }
            if (m == 1 && in_first_run) { first_ended = true; first_clean = !P->same_block || lb >= 0x80; }
            if (m == 1) in_first_run = false;
            at_left = false;
            // helper.rs:327-330: without a `break` the same next() call goes on and keeps its (possibly stale)
            // last_multi_char_leading_byte; after a `break` the following next() call starts from 0
            const uint32_t keep_mb = broke ? 0u : last_mb;
            new_run();
            last_mb = keep_mb;
        }
    }

    // End of the segment text.  invalid_after: finding_collection.rs:234-237.
    SX_HD void segment_end(bool invalid_after) {
        if (!dead) {
            if (qfull) resolve_qfull(true, invalid_after);
            else if (run_n > 0) {
                const bool completes = at_left && last_cut;
                const bool again = !completes && !invalid_after;  // helper.rs:389-392 (run_n < q here)
                if (again) keep_leftover();
                else if (completes || (run_n >= P->n && grep_ok)) yield(completes, !invalid_after);  // helper.rs:353-354, :410-415
            }
        }
        if (m == 1) in_first_run = false;
        run_n = 0; run_out = 0; qfull = false;
    }

    SX_HD void on_malformed(int64_t next) {
        segment_end(true);
        segment_start(next, nullptr, 0);
    }

    SX_HD Carry carry_out(int64_t boundary) const {
        if (cut) return carry_cut();
        if (has_left) {
            Carry c;
            c.kind = K_L;
            c.flags = (uint8_t)((left_hostcarry ? CF_HOSTCARRY : 0) | (left_grep ? CF_GREP : 0) | (left_half ? CF_HALF : 0));
            c.k = (uint16_t)left_k;
            c.in_bytes = (uint32_t)(boundary - left_in_start);
            c.out_bytes = left_out;
            c.aux = left_mb;
            return c;
        }
        return carry_none();
    }
};

// ------------------------------------------------------------------------------------------
// Decoders (restating encoding_rs 0.8.34 semantics event-wise; see DESIGN.md section "Decoders").
// step(b, pos, emit): emit.ch(lb, ul, cstart, cend) for a decoded char, emit.mal(next) for a
// malformed sequence after which the next decoder call (segment) starts at `next`.
// ------------------------------------------------------------------------------------------
struct DecXud {
    template <class S> SX_HD void init(const ScanParams&, const S&, int64_t) {}
    SX_HD int32_t pending_len() const { return 0; }
    static constexpr bool kStateful = false;
    template <class E> SX_HD void step(const ScanParams&, uint32_t b, int64_t pos, E& e) {
        if (b < 0x80) e.ch(b, 1, pos, pos + 1);
        else e.ch(0xEF, 3, pos, pos + 1);  // U+F780 + (b - 0x80): UTF-8 lead byte 0xEF
    }
    template <class E> SX_HD void eof(E&) {}
};

struct DecSb {
    template <class S> SX_HD void init(const ScanParams&, const S&, int64_t) {}
    SX_HD int32_t pending_len() const { return 0; }
    static constexpr bool kStateful = false;
    template <class E> SX_HD void step(const ScanParams& P, uint32_t b, int64_t pos, E& e) {
        if (b < 0x80) { e.ch(b, 1, pos, pos + 1); return; }
        const uint32_t cp = P.sb_table[b - 0x80];
        if (cp == 0) e.mal(pos + 1);
        else e.ch(utf8_lead_of_cp(cp), utf8_len_of_cp(cp), pos, pos + 1);
    }
    template <class E> SX_HD void eof(E&) {}
};

struct DecUtf8 {
    uint32_t need, seen, lead, lo, hi;
    static constexpr bool kStateful = true;
    SX_HD void reset() { need = 0; seen = 0; lead = 0; lo = 0x80; hi = 0xBF; }
    SX_HD int32_t pending_len() const { return need ? (int32_t)seen + 1 : 0; }
    SX_HD void start(uint32_t b) {  // b is a valid lead byte C2..F4
        lead = b; seen = 0; lo = 0x80; hi = 0xBF;
        if (b < 0xE0) need = 1;
        else if (b < 0xF0) { need = 2; if (b == 0xE0) lo = 0xA0; else if (b == 0xED) hi = 0x9F; }
        else { need = 3; if (b == 0xF0) lo = 0x90; else if (b == 0xF4) hi = 0x8F; }
    }
    // Decoder state at `ws` from the <= 3 preceding bytes: the WHATWG UTF-8 decoder is in the
    // neutral state after any byte outside 0x80..0xBF that is not a lead, and a lead byte is
    // always (re-)read in the neutral state, so the state only depends on the bytes since the
    // last non-continuation byte.
    template <class S> SX_HD void init(const ScanParams& P, const S& src, int64_t ws) {
        reset();
        const int64_t lo_off = -(int64_t)P.npend;
        for (int j = 1; j <= 3; ++j) {
            const int64_t o = ws - j;
            if (o < lo_off) return;
            const uint32_t b = src.get(o);
            if ((b & 0xC0) == 0x80) continue;  // continuation byte: keep looking for the start byte
            if (b < 0xC2 || b > 0xF4) return;  // ASCII or invalid lead: neutral
            start(b);
            if ((int)need < j) { reset(); return; }  // sequence already complete (or overrun)
            for (int i = j - 1; i >= 1; --i) {       // replay the j-1 continuation bytes
                const uint32_t c = src.get(ws - i);
                if (c < lo || c > hi) { reset(); return; }
                lo = 0x80; hi = 0xBF; seen++;
            }
            if (seen == need) reset();
            return;
        }
    }
    template <class E> SX_HD void step(const ScanParams&, uint32_t b, int64_t pos, E& e) {
        if (need != 0) {
            if (b >= lo && b <= hi) {
                lo = 0x80; hi = 0xBF; seen++;
                if (seen == need) { const uint32_t l = lead, nn = need; reset(); e.ch(l, nn + 1, pos - nn, pos + 1); }
                return;
            }
            reset();
            e.mal(pos);  // the offending byte is not consumed: it starts the next segment
        }
        if (b < 0x80) { e.ch(b, 1, pos, pos + 1); return; }
        if (b < 0xC2 || b > 0xF4) { e.mal(pos + 1); return; }
        start(b);
    }
    template <class E> SX_HD void eof(E&) { reset(); }
};

template <bool BE>
struct DecUtf16 {
    uint32_t has_lead, lead_byte, lead_sur;  // lead_sur != 0: pending high surrogate
    static constexpr bool kStateful = true;
    SX_HD int32_t pending_len() const { return (lead_sur ? 2 : 0) + (has_lead ? 1 : 0); }
    SX_HD static uint32_t unit(uint32_t first, uint32_t second) { return BE ? ((first << 8) | second) : ((second << 8) | first); }
    // The unit grid is global ((o - align) even); a high surrogate is always pending after it has
    // been processed, whatever preceded it; a pending BMP unit never survives a decoder call
    // boundary because the reference flushes it with an empty call (finding_collection.rs:134-143).
    template <class S> SX_HD void init(const ScanParams& P, const S& src, int64_t ws) {
        has_lead = 0; lead_byte = 0; lead_sur = 0;
        const int64_t lo_off = -(int64_t)P.npend;
        int64_t ub = ws;  // start of the unit containing / following ws
        if (((ws - (int64_t)P.align) & 1) != 0) {
            if (ws - 1 >= lo_off) { has_lead = 1; lead_byte = src.get(ws - 1); ub = ws - 1; }
        }
        if (ub - 2 >= lo_off) {
            const uint32_t u = unit(src.get(ub - 2), src.get(ub - 1));
            if ((u & 0xFC00) == 0xD800) lead_sur = u;
        }
    }
    template <class E> SX_HD void step(const ScanParams&, uint32_t b, int64_t pos, E& e) {
        if (!has_lead) { has_lead = 1; lead_byte = b; return; }
        has_lead = 0;
        const uint32_t cu = unit(lead_byte, b);
        const uint32_t hb = cu & 0xFC00;
        if (hb == 0xD800) {
            if (lead_sur) { lead_sur = cu; e.mal(pos + 1); return; }
            lead_sur = cu;
            return;
        }
        if (hb == 0xDC00) {
            if (!lead_sur) { e.mal(pos + 1); return; }
            const uint32_t cp = 0x10000u + ((lead_sur - 0xD800u) << 10) + (cu - 0xDC00u);
            lead_sur = 0;
            e.ch(0xF0 | (cp >> 18), 4, pos - 3, pos + 1);
            return;
        }
        if (lead_sur) {
            // Malformed(2,2): the BMP unit is consumed and written first thing by the next call
            lead_sur = 0;
            e.mal(pos + 1);
            e.ch(utf8_lead_of_cp(cu), utf8_len_of_cp(cu), pos - 1, pos + 1);
            return;
        }
        e.ch(utf8_lead_of_cp(cu), utf8_len_of_cp(cu), pos - 1, pos + 1);
    }
    template <class E> SX_HD void eof(E&) { has_lead = 0; lead_sur = 0; }
};


template <bool BE>
struct DecUtf32 {  // EXTENSION: no reference semantics (mission.rs:681-688 rejects utf-32)
    uint32_t nb, acc;
    static constexpr bool kStateful = true;
    SX_HD int32_t pending_len() const { return (int32_t)nb; }
    template <class S> SX_HD void init(const ScanParams& P, const S& src, int64_t ws) {
        nb = 0; acc = 0;
        const int64_t lo_off = -(int64_t)P.npend;
        int64_t r = (ws - (int64_t)P.align) & 3;  // bytes of the current unit already before ws
        if (ws - r < lo_off) r = ws - lo_off;     // stream start inside the unit (cannot happen with a sane align)
        for (int64_t o = ws - r; o < ws; ++o) push(src.get(o));
    }
    SX_HD void push(uint32_t b) { acc = BE ? ((acc << 8) | b) : (acc | (b << (8 * nb))); nb++; }
    template <class E> SX_HD void step(const ScanParams&, uint32_t b, int64_t pos, E& e) {
        push(b);
        if (nb < 4) return;
        const uint32_t c = acc;
        nb = 0; acc = 0;
        if (c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) { e.mal(pos + 1); return; }
        e.ch(utf8_lead_of_cp(c), utf8_len_of_cp(c), pos - 3, pos + 1);
    }
    template <class E> SX_HD void eof(E&) { nb = 0; acc = 0; }
};

// ------------------------------------------------------------------------------------------
// Big5 and EUC-JP (WHATWG Encoding Standard decoders, as encoding_rs implements them; tables from
// tools/gen_multibyte_tables.py -- parity "self-consistent").
// Lead / trail roles depend on the parse history, so the state at a window start is found by walking back to the last
// byte that leaves the decoder neutral whatever came before it (Big5: a byte outside 81-FE; EUC-JP: a byte outside
// {8E, 8F, A1-FE}) and parsing forward from there.  The walk is bounded: kMbMaxBack bytes without
// such a byte (never seen outside adversarial input) raise SX_ERR_UNSUPPORTED through mb_walk_failed.
// ------------------------------------------------------------------------------------------
constexpr int64_t kMbMaxBack = 256 << 10;
SX_HD bool big5_is_run_byte(uint32_t b) { return b >= 0x81 && b <= 0xFE; }
SX_HD bool big5_is_trail(uint32_t b) { return (b >= 0x40 && b <= 0x7E) || (b >= 0xA1 && b <= 0xFE); }
SX_HD uint32_t big5_pointer(uint32_t lead, uint32_t trail) { return (lead - 0x81u) * 157u + (trail - (trail < 0x7Fu ? 0x40u : 0x62u)); }
// the four pointers that decode to TWO code points (U+00CA / U+00EA followed by U+0304 / U+030C)
SX_HD bool big5_is_double(uint32_t ptr) { return ptr == 1133u || ptr == 1135u || ptr == 1164u || ptr == 1166u; }
SX_HD uint32_t big5_double_first(uint32_t ptr) { return ptr < 1150u ? 0xCAu : 0xEAu; }
SX_HD uint32_t big5_double_second(uint32_t ptr) { return (ptr == 1133u || ptr == 1164u) ? 0x304u : 0x30Cu; }
SX_HD bool eucjp_is_run_byte(uint32_t b) { return b == 0x8E || b == 0x8F || (b >= 0xA1 && b <= 0xFE); }

struct DecBig5 {
    uint32_t lead;  // pending lead byte, 0: neutral
    static constexpr bool kStateful = true;
    SX_HD int32_t pending_len() const { return lead ? 1 : 0; }
    template <class S> SX_HD void init(const ScanParams& P, const S& src, int64_t ws) {
        lead = 0;
        const int64_t lo_off = -(int64_t)P.npend;
        int64_t o = ws - 1;
        while (o >= lo_off && big5_is_run_byte(src.get(o))) {
            if (ws - o > kMbMaxBack) { if (P.mb_fail) *P.mb_fail = 1u; break; }
            --o;
        }
        // bytes (o, ws) are all 81-FE and pair up, mapped or not (an unmapped pair with a non-ASCII trail is consumed whole)
        if (((ws - 1 - o) & 1) != 0) lead = src.get(ws - 1);
    }
    template <class E> SX_HD void step(const ScanParams& P, uint32_t b, int64_t pos, E& e) {
        if (lead) {
            const uint32_t l = lead;
            lead = 0;
            if (big5_is_trail(b)) {
                const uint32_t ptr = big5_pointer(l, b);
                if (big5_is_double(ptr)) {
                    e.cp(big5_double_first(ptr));
                    e.ch(0xC3, 2, pos - 1, pos + 1);   // U+00CA / U+00EA
                    e.cp(big5_double_second(ptr));
                    e.ch2(0xCC, 2, pos - 1, pos + 1);  // U+0304 / U+030C
                    return;
                }
                const uint32_t cp = P.mb_a[ptr];
                if (cp) { e.cp(cp); e.ch(utf8_lead_of_cp(cp), utf8_len_of_cp(cp), pos - 1, pos + 1); return; }
            }
            if (b >= 0x80) { e.mal(pos + 1); return; }
            e.mal(pos);  // an ASCII byte is not consumed: it starts the next segment
        }
        if (b < 0x80) { e.cp(b); e.ch(b, 1, pos, pos + 1); return; }
        if (b == 0x80 || b == 0xFF) { e.mal(pos + 1); return; }
        lead = b;
    }
    template <class E> SX_HD void eof(E&) { lead = 0; }
};

struct DecEucJp {
    uint32_t lead;   // pending byte: 8E, 8F or A1-FE (with j0212: the byte after 8F), 0: neutral
    uint32_t j0212;
    static constexpr bool kStateful = true;
    SX_HD int32_t pending_len() const { return lead ? (j0212 ? 2 : 1) : 0; }
    // state transition without events (used by init to parse forward from the last neutral point)
    SX_HD void advance(uint32_t b) {
        if (lead == 0x8F && !j0212 && b >= 0xA1 && b <= 0xFE) { j0212 = 1; lead = b; return; }
        if (lead) { lead = 0; j0212 = 0; return; }  // whatever run byte follows a lead is consumed with it
        lead = b;
    }
    template <class S> SX_HD void init(const ScanParams& P, const S& src, int64_t ws) {
        lead = 0; j0212 = 0;
        const int64_t lo_off = -(int64_t)P.npend;
        int64_t o = ws - 1;
        while (o >= lo_off && eucjp_is_run_byte(src.get(o))) {
            if (ws - o > kMbMaxBack) { if (P.mb_fail) *P.mb_fail = 1u; break; }
            --o;
        }
        for (int64_t p = o + 1; p < ws; ++p) advance(src.get(p));
    }
    template <class E> SX_HD void step(const ScanParams& P, uint32_t b, int64_t pos, E& e) {
        if (lead == 0x8E && !j0212 && b >= 0xA1 && b <= 0xDF) {  // U+FF61..U+FF9F
            lead = 0;
            e.cp(0xFF61u - 0xA1u + b);
            e.ch(0xEF, 3, pos - 1, pos + 1);
            return;
        }
        if (lead == 0x8F && !j0212 && b >= 0xA1 && b <= 0xFE) { j0212 = 1; lead = b; return; }
        if (lead) {
            const uint32_t l = lead, three = j0212;
            lead = 0; j0212 = 0;
            if (l >= 0xA1 && l <= 0xFE && b >= 0xA1 && b <= 0xFE) {
                const uint32_t cp = (three ? P.mb_b : P.mb_a)[(l - 0xA1u) * 94u + (b - 0xA1u)];
                if (cp) { e.cp(cp); e.ch(utf8_lead_of_cp(cp), utf8_len_of_cp(cp), pos - (three ? 2 : 1), pos + 1); return; }
            }
            if (b >= 0x80) { e.mal(pos + 1); return; }
            e.mal(pos);  // an ASCII byte is not consumed: it starts the next segment
        }
        if (b < 0x80) { e.cp(b); e.ch(b, 1, pos, pos + 1); return; }
        if (eucjp_is_run_byte(b)) { lead = b; return; }
        e.mal(pos + 1);
    }
    template <class E> SX_HD void eof(E&) { lead = 0; j0212 = 0; }
};
SX_HD void mb_set_state(DecBig5& d, uint32_t lead, uint32_t) { d.lead = lead; }
SX_HD void mb_set_state(DecEucJp& d, uint32_t lead, uint32_t j) { d.lead = lead; d.j0212 = j; }

// ------------------------------------------------------------------------------------------
// Precision::Before probe (finding_collection.rs:176-207), literal emulation of the 8-byte
// re-decode with a fresh decoder.  Only reached for the first char of a segment that starts
// at a slice start and whose first decoded byte is >= 0x80 -- rare, so clarity over speed.
// ------------------------------------------------------------------------------------------
SX_HD uint32_t utf8_valid_up_to8(const uint8_t* s, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        const uint32_t b = s[i];
        if (b < 0x80) { i++; continue; }
        if (b < 0xC2) return i;
        if (b < 0xE0) { if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return i; i += 2; continue; }
        if (b < 0xF0) {
            const uint32_t lo = b == 0xE0 ? 0xA0 : 0x80, hi = b == 0xED ? 0x9F : 0xBF;
            if (i + 2 >= n || s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80) return i;
            i += 3; continue;
        }
        if (b < 0xF5) {
            const uint32_t lo = b == 0xF0 ? 0x90 : 0x80, hi = b == 0xF4 ? 0x8F : 0xBF;
            if (i + 3 >= n || s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80 || (s[i + 3] & 0xC0) != 0x80) return i;
            i += 4; continue;
        }
        return i;
    }
    return i;
}

// `written` of a fresh encoding_rs UTF-8 decoder over src[0..slen) into an 8-byte buffer, last = true.
// For UTF-8 the output equals the input bytes, so only the count is needed.
SX_HD uint32_t utf8_fresh_written8(const uint8_t* src, uint32_t slen) {
    uint32_t sp = 0, dp = 0, need = 0, seen = 0, lo = 0x80, hi = 0xBF;
    for (;;) {
        if (need == 0) {
            const uint32_t n = (slen - sp) < (8 - dp) ? (slen - sp) : (8 - dp);
            const uint32_t v = utf8_valid_up_to8(src + sp, n);
            sp += v; dp += v;
        }
        if (sp >= slen) return dp;
        if (!(dp + 3 < 8)) return dp;
        const uint32_t b = src[sp++];
        if (need == 0) {
            if (b < 0x80) { dp++; continue; }
            if (b < 0xC2 || b > 0xF4) return dp;
            seen = 0; lo = 0x80; hi = 0xBF;
            if (b < 0xE0) need = 1;
            else if (b < 0xF0) { need = 2; if (b == 0xE0) lo = 0xA0; else if (b == 0xED) hi = 0x9F; }
            else { need = 3; if (b == 0xF0) lo = 0x90; else if (b == 0xF4) hi = 0x8F; }
            continue;
        }
        if (b < lo || b > hi) return dp;
        lo = 0x80; hi = 0xBF; seen++;
        if (seen == need) { dp += need + 1; need = 0; }
    }
}

// UTF-8 probe.  `first_segment`: the probing segment is the first one of the slice.
// Returns true for Precision::Before.
SX_HD_NOINLINE bool probe_utf8(const ScanParams& P, const GlobalSrc& g, int64_t slice_start, int64_t slice_end, bool first_segment,
                      int32_t slice_pending, const Carry& slice_left) {
    const bool has_left = slice_left.kind == K_L && slice_left.k > 0;
    if (first_segment) return has_left || slice_pending > 0;  // fresh decoder hits a continuation byte: written == 0
    if (!has_left) return false;  // same bytes, same neutral state: identical
    // Second segment at offset 0 (after a Malformed with read == 0): the stale leftover text still
    // sits at out[0..], the new text follows it.  Compare literally.
    uint8_t src8[16], out8[24];
    uint32_t sl = 0;
    for (; sl < 16 && slice_start + sl < slice_end; ++sl) src8[sl] = g.get(slice_start + sl);
    const uint32_t w2 = utf8_fresh_written8(src8, sl);
    if (w2 == 0) return true;
    uint32_t ol = 0;
    if (slice_left.flags & CF_HOSTCARRY)
        for (uint32_t i = 0; i < P.carry_text_len && i < 8 && ol < 8; ++i) out8[ol++] = P.carry_text8[i];
    // device part of the leftover text = raw input bytes (UTF-8 -> UTF-8)
    {
        const int64_t ls = slice_start - (int64_t)slice_left.in_bytes;
        for (uint32_t i = 0; i < slice_left.out_bytes && ol < 8; ++i) out8[ol++] = g.get(ls + i);
    }
    for (uint32_t i = 0; ol < 8 && i < sl; ++i) out8[ol++] = src8[i];  // (over-)approximates the new text; only w2 valid bytes are compared
    for (uint32_t i = 0; i < w2; ++i)
        if ((i < ol ? out8[i] : 0) != src8[i]) return true;
    return false;
}

// Decode UTF-16 units from `pos` with the given decoder state into at most `cap` bytes the way
// encoding_rs does (space check dp + 3 < cap before every byte read when `checked`), stopping at
// the first malformed sequence or `end`.  Returns bytes written.
template <bool BE>
SX_HD uint32_t utf16_decode_some(const GlobalSrc& g, int64_t pos, int64_t end, uint32_t has_lead, uint32_t lead_byte,
                                 uint32_t lead_sur, uint8_t* dst, uint32_t cap, bool checked) {
    uint32_t dp = 0;
    for (; pos < end; ++pos) {
        if (checked ? !(dp + 3 < cap) : (dp + 4 > cap)) break;
        const uint32_t b = g.get(pos);
        if (!has_lead) { has_lead = 1; lead_byte = b; continue; }
        has_lead = 0;
        const uint32_t cu = BE ? ((lead_byte << 8) | b) : ((b << 8) | lead_byte);
        const uint32_t hb = cu & 0xFC00;
        if (hb == 0xD800) { if (lead_sur) break; lead_sur = cu; continue; }
        if (hb == 0xDC00) {
            if (!lead_sur) break;
            dp += put_utf8(dst + dp, 0x10000u + ((lead_sur - 0xD800u) << 10) + (cu - 0xDC00u));
            lead_sur = 0;
            continue;
        }
        if (lead_sur) break;
        dp += put_utf8(dst + dp, cu);
    }
    return dp;
}

template <bool BE>
SX_HD_NOINLINE bool probe_utf16(const ScanParams& P, const GlobalSrc& g, int64_t slice_start, int64_t slice_end, int64_t win_end,
                       uint32_t has_lead, uint32_t lead_byte, uint32_t lead_sur, const Carry& slice_left) {
    (void)P;
    if (slice_left.kind == K_L && slice_left.k > 0) return true;  // leftover prepended: Before anyway
    if (!has_lead && !lead_sur) return false;                     // aligned and neutral: identical
    uint8_t fb[16], ob[16];
    for (int i = 0; i < 16; ++i) { fb[i] = 0; ob[i] = 0; }
    const uint32_t w2 = utf16_decode_some<BE>(g, slice_start, slice_end, 0, 0, 0, fb, 8, true);
    if (w2 == 0) return true;
    // the real first decoder call of the slice covers the first window only
    (void)utf16_decode_some<BE>(g, slice_start, win_end, has_lead, lead_byte, lead_sur, ob, 12, false);
    for (uint32_t i = 0; i < w2; ++i)
        if (ob[i] != fb[i]) return true;
    return false;
}

template <bool BE>
SX_HD_NOINLINE bool probe_utf32(const GlobalSrc& g, int64_t slice_start, int64_t slice_end, int64_t win_end, uint32_t nb,
                       const Carry& slice_left) {
    if (slice_left.kind == K_L && slice_left.k > 0) return true;
    if (nb == 0 && win_end - slice_start >= 20) return false;  // aligned, and the window covers all the probe reads
    // decode both ways (fresh: 4-byte units from the slice start into an 8-byte buffer; real: first
    // decoder call of the slice = first window only, the rest of `out` is still zero)
    uint8_t fb[16], ob[16];
    for (int i = 0; i < 16; ++i) { fb[i] = 0; ob[i] = 0; }
    uint32_t w2 = 0;
    {
        uint32_t dp = 0, k = 0, acc = 0;
        for (int64_t p = slice_start; p < slice_end; ++p) {
            if (!(dp + 3 < 8)) break;
            const uint32_t b = g.get(p);
            acc = BE ? ((acc << 8) | b) : (acc | (b << (8 * k)));
            if (++k < 4) continue;
            const uint32_t c = acc; k = 0; acc = 0;
            if (c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) break;
            dp += put_utf8(fb + dp, c);
        }
        w2 = dp;
    }
    if (w2 == 0) return true;
    {
        uint32_t dp = 0, k = 0, acc = 0;
        for (int64_t p = slice_start - nb; p < win_end && dp + 4 <= 12; ++p) {
            const uint32_t b = g.get(p);
            acc = BE ? ((acc << 8) | b) : (acc | (b << (8 * k)));
            if (++k < 4) continue;
            const uint32_t c = acc; k = 0; acc = 0;
            if (c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) break;
            dp += put_utf8(ob + dp, c);
        }
    }
    for (uint32_t i = 0; i < w2; ++i)
        if (ob[i] != fb[i]) return true;
    return false;
}

// Text of decoded chars (Big5 / EUC-JP): the decoders report every code point through cp() before ch().
struct MbTextEmit {
    uint8_t* dst;
    uint32_t dp, cap;
    uint32_t skip;   // code points to drop at the front (a text that starts with the second half of a Big5 pair)
    uint32_t last;
    bool stop;
    SX_HD void cp(uint32_t c) { last = c; }
    SX_HD void ch(uint32_t, uint32_t ul, int64_t, int64_t) {
        if (stop) return;
        if (skip) { --skip; return; }
        if (dp + ul > cap) { stop = true; return; }
        dp += put_utf8(dst + dp, last);
    }
    SX_HD void ch2(uint32_t lb, uint32_t ul, int64_t a, int64_t b) { ch(lb, ul, a, b); }
    SX_HD void mal(int64_t) { stop = true; }
};
// Decode from `pos` with decoder state `d` into at most `cap` bytes, stopping at the first malformed sequence or `end`.
// need_free != 0: like encoding_rs, require that many free bytes before every byte read (check_space_bmp: 3,
// check_space_astral: 4); otherwise stop when a char no longer fits.  Returns the bytes written.
template <class Dec>
SX_HD uint32_t mb_decode_some(const ScanParams& P, const GlobalSrc& g, Dec d, int64_t pos, int64_t end, uint8_t* dst, uint32_t cap,
                              uint32_t need_free) {
    MbTextEmit te{dst, 0, cap, 0, 0, false};
    for (; pos < end && !te.stop; ++pos) {
        if (need_free ? (te.dp + need_free > cap) : (te.dp + 4 > cap)) break;
        d.step(P, g.get(pos), pos, te);
    }
    return te.dp;
}
template <class Dec> struct MbTraits { static constexpr uint32_t kNeedFree = 3; };
template <> struct MbTraits<DecBig5> { static constexpr uint32_t kNeedFree = 4; };  // astral / two code points
// Precision::Before probe for the lead / trail encodings (finding_collection.rs:176-207), literal: a fresh decoder over
// the slice into 8 bytes vs what the real first decoder call of the slice (the first window, real pending state) wrote.
template <class Dec>
SX_HD_NOINLINE bool probe_mb(const ScanParams& P, const GlobalSrc& g, int64_t slice_start, int64_t slice_end, int64_t win_end,
                             uint32_t lead0, uint32_t j0, const Carry& slice_left) {
    if (slice_left.kind == K_L && slice_left.k > 0) return true;  // leftover prepended: Before anyway
    if (lead0 == 0 && win_end - slice_start >= 24) return false;  // same state, same bytes, and the window covers the probe
    uint8_t fb[16], ob[16];
    for (int i = 0; i < 16; ++i) { fb[i] = 0; ob[i] = 0; }
    Dec fresh, real;
    mb_set_state(fresh, 0, 0);
    mb_set_state(real, lead0, j0);
    const uint32_t w2 = mb_decode_some<Dec>(P, g, fresh, slice_start, slice_end, fb, 8, MbTraits<Dec>::kNeedFree);
    if (w2 == 0) return true;
    (void)mb_decode_some<Dec>(P, g, real, slice_start, win_end, ob, 12, 0);
    for (uint32_t i = 0; i < w2; ++i)
        if (ob[i] != fb[i]) return true;
    return false;
}

// ------------------------------------------------------------------------------------------
// Window driver.
// ------------------------------------------------------------------------------------------
struct WinGeom {
    int64_t ws, we;                 // window [ws, we) in buffer offsets
    int64_t slice_start, slice_end; // enclosing slice
    bool final_last;                // last window of the stream and is_last_input_buffer
    int force_s2_cont;              // classification aid (general missions): -1 or the carry flag forced into segment 2
};

template <class Dec>
struct ProbeCtx {
    const ScanParams* P;
    const GlobalSrc* g;
    const WinGeom* geo;
    // decoder state at the slice start (only meaningful when geo->ws == geo->slice_start)
    int32_t pend0;
    uint32_t s0, s1, s2;
};

// The probes are out of line and take plain values, so the automaton state never has its address
// taken and stays in registers.
template <class Dec> struct ProbeImpl {
    SX_HD static bool run(const ProbeCtx<Dec>&, bool, const Carry&) { return false; }
};
template <> struct ProbeImpl<DecUtf8> {
    SX_HD static bool run(const ProbeCtx<DecUtf8>& c, bool first_segment, const Carry& slice_left) {
        return probe_utf8(*c.P, *c.g, c.geo->slice_start, c.geo->slice_end, first_segment, c.pend0, slice_left);
    }
};
template <bool BE> struct ProbeImpl<DecUtf16<BE>> {
    SX_HD static bool run(const ProbeCtx<DecUtf16<BE>>& c, bool, const Carry& slice_left) {
        return probe_utf16<BE>(*c.P, *c.g, c.geo->slice_start, c.geo->slice_end, c.geo->we, c.s0, c.s1, c.s2, slice_left);
    }
};
template <bool BE> struct ProbeImpl<DecUtf32<BE>> {
    SX_HD static bool run(const ProbeCtx<DecUtf32<BE>>& c, bool, const Carry& slice_left) {
        return probe_utf32<BE>(*c.g, c.geo->slice_start, c.geo->slice_end, c.geo->we, c.s0, slice_left);
    }
};

template <> struct ProbeImpl<DecBig5> {
    SX_HD static bool run(const ProbeCtx<DecBig5>& c, bool, const Carry& slice_left) {
        return probe_mb<DecBig5>(*c.P, *c.g, c.geo->slice_start, c.geo->slice_end, c.geo->we, c.s0, c.s1, slice_left);
    }
};
template <> struct ProbeImpl<DecEucJp> {
    SX_HD static bool run(const ProbeCtx<DecEucJp>& c, bool, const Carry& slice_left) {
        return probe_mb<DecEucJp>(*c.P, *c.g, c.geo->slice_start, c.geo->slice_end, c.geo->we, c.s0, c.s1, slice_left);
    }
};

template <class Dec> SX_HD void probe_capture(ProbeCtx<Dec>& c, const Dec& d) { (void)c; (void)d; }
SX_HD void probe_capture(ProbeCtx<DecBig5>& c, const DecBig5& d) { c.s0 = d.lead; c.s1 = 0; }
SX_HD void probe_capture(ProbeCtx<DecEucJp>& c, const DecEucJp& d) { c.s0 = d.lead; c.s1 = d.j0212; }
SX_HD void probe_capture(ProbeCtx<DecUtf8>& c, const DecUtf8& d) { c.pend0 = d.pending_len(); }
template <bool BE> SX_HD void probe_capture(ProbeCtx<DecUtf16<BE>>& c, const DecUtf16<BE>& d) {
    c.s0 = d.has_lead; c.s1 = d.lead_byte; c.s2 = d.lead_sur;
}
template <bool BE> SX_HD void probe_capture(ProbeCtx<DecUtf32<BE>>& c, const DecUtf32<BE>& d) { c.s0 = d.nb; }

template <class Dec>
struct Emit {
    WinAuto* A;
    const ProbeCtx<Dec>* pc;
    SX_HD void ch(uint32_t lb, uint32_t ul, int64_t cs, int64_t ce) {
        const ProbeCtx<Dec>* c = pc;
        A->on_char(lb, ul, cs, ce, [c](bool first_segment, Carry slice_left) { return ProbeImpl<Dec>::run(*c, first_segment, slice_left); });
    }
    SX_HD void cp(uint32_t) {}  // the scanner only needs the UTF-8 lead byte and length of a char
    // second code point of a two-code-point pair (Big5): same input bytes as the first one
    SX_HD void ch2(uint32_t lb, uint32_t ul, int64_t cs, int64_t ce) {
        A->next_half2 = true;
        ch(lb, ul, cs, ce);
        A->next_half2 = false;
    }
    SX_HD void mal(int64_t next) { A->on_malformed(next); }
};

// TileSrc concept: uint8_t get(int64_t off) for any stream offset; load16(int64_t off16, uint32_t w[4])
// loads the aligned 16-byte chunk starting at tile-relative-aligned offset (off16 % 16 == 0 relative
// to the tile base) -- both provided by the kernel / the emulation harness.
// NOT inlined on the device: the kernels call it from several places and one copy of the byte loop
// keeps the hot code inside the instruction cache (ncu: stall_no_instruction dominated otherwise).
template <class Dec, class TileSrc>
SX_HD_NOINLINE void scan_window(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo, const Carry& kin,
                       int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
    Dec dec;
    dec.init(P, tsrc, geo.ws);
    ProbeCtx<Dec> pc;
    pc.P = &P; pc.g = &g; pc.geo = &geo; pc.pend0 = 0; pc.s0 = pc.s1 = pc.s2 = 0;
    probe_capture(pc, dec);
    WinAuto A;
    A.init(&P, mode, geo.slice_start, Dec::kStateful);
    A.wr = wr; A.text_off = text_off;
    A.force_s2_cont = geo.force_s2_cont;
    if (geo.ws == geo.slice_start) { A.slice_left_present = true; A.slice_left = kin; }
    A.segment_start(geo.ws, &kin, dec.pending_len());
    Emit<Dec> em{&A, &pc};
    tsrc.for_each_byte(geo.ws, geo.we, [&](uint32_t b, int64_t pos) { dec.step(P, b, pos, em); });
    if (geo.final_last) {
        // finding_collection.rs:234-237 / :298-304: invalid_after for every segment of the last window,
        // then one flush round whose only lasting effects are a reset decoder and cut == false.
        A.segment_end(true);
        dec.eof(em);
        A.cut = false;
        A.has_left = false;
        res.npend_out = 0;
    } else {
        A.segment_end(false);
        res.npend_out = dec.pending_len();
    }
    res.out = A.carry_out(geo.we);
    res.nrec = A.nrec;
    res.ntext = A.ntext;
    res.m = A.m;
    res.cut1 = A.cut1 ? 1u : 0u;
    if (desc) {
        desc->a = (uint16_t)A.a;
        desc->nrec = (uint16_t)(A.nrec > 0xFFFFu ? 0xFFFFu : A.nrec);
        desc->ntext = A.ntext;
        desc->null_out = res.out;
        desc->t_out = 0;
        desc->pad = 0;
        const bool single_all_pass = (A.m == 1 && A.s1_all_pass);
        if (geo.final_last) desc->type = WT_CONST;
        else if (P.general) {
            // >= 2 segments: refined by classify_general() with a second pass.  One segment whose first run ends inside
            // the window under the null carry with a < q chars: past that point the automaton holds nothing of the
            // carry-in (run state reset, cut == false, no leftover; `first_clean`: no stale lead byte either) unless
            // the carried leftover drives the first run to q chars (k + a >= q: forced cut, possibly a dead segment,
            // helper.rs:389-415).  k <= q - 1 without a grep_char, <= q with one.
            desc->type = WT_DEP;
            if (A.m == 1 && A.first_ended && A.first_clean && A.a < P.q) {
                const uint32_t kmax = P.grep_char >= 0 ? P.q : P.q - 1;
                desc->type = (A.a + kmax < P.q) ? WT_CONST : WT_GUARD;
            }
        }
        else if (single_all_pass) {
            if (A.a < P.q) { desc->type = WT_CASEB; desc->t_out = (uint16_t)A.s1_out; }
            else desc->type = WT_CONST;
        } else if (A.a > 0 && !A.s1_later_yield && (A.m == 1 || (A.m == 2 && A.s2_all_pass))) desc->type = WT_DEP;
        else desc->type = WT_CONST;
    }
}

// Transfer function of a WT_CASEB window (single segment, every char passes, t = a < q chars).
SX_HD Carry eval_caseb(const ScanParams& P, const WinDesc& d, const Carry& kin, uint32_t wlen) {
    if (kin.kind == K_UNKNOWN) return kin;
    if (kin.kind == K_C) return d.a >= 1 ? carry_cut() : carry_none();
    if (kin.k == 0) return d.null_out;
    if ((uint32_t)kin.k + d.a < P.q) {
        Carry c = kin;
        c.k = (uint16_t)(kin.k + d.a);
        c.in_bytes = kin.in_bytes + wlen;
        c.out_bytes = kin.out_bytes + d.t_out;
        return c;
    }
    return carry_cut();
}

// Window / slice geometry (finding_collection.rs:124-131, input.rs:22).
struct Geometry {
    int64_t len;
    uint32_t slice_len, W, wps;  // wps = windows per full slice
    int32_t is_last;
    SX_HD void init(const ScanParams& P) {
        len = P.len; slice_len = P.slice_len; W = P.W; is_last = P.is_last;
        wps = (slice_len + W - 1) / W;
    }
    // window index -> geometry; false when the window lies beyond the stream
    SX_HD bool window(int64_t widx, WinGeom& g) const {
        const int64_t s = widx / wps;
        const int64_t j = widx - s * wps;
        g.slice_start = s * (int64_t)slice_len;
        if (g.slice_start >= len) return false;
        g.slice_end = g.slice_start + slice_len < len ? g.slice_start + slice_len : len;
        g.ws = g.slice_start + j * (int64_t)W;
        if (g.ws >= g.slice_end) return false;
        g.we = g.ws + W < g.slice_end ? g.ws + W : g.slice_end;
        g.final_last = is_last && g.we == len;
        g.force_s2_cont = -1;
        return true;
    }
};

// Carry into a listed window whose predecessor is NOT listed (prefilter: the predecessor holds no run
// of >= T good bytes, so a definite breaker lies within its last T bytes and its carry-out does not
// depend on its own carry-in): replay the predecessor's tail from the null carry.
SX_HD WinGeom preroll_geom(const Geometry& geo, int64_t w, uint32_t pre_bytes) {
    WinGeom pg;
    geo.window(w - 1, pg);
    // at least pre_bytes; rounded down to a 16-byte boundary so the replay can use aligned vector loads
    const int64_t s = (pg.we - (int64_t)pre_bytes) & ~(int64_t)15;
    if (s > pg.ws) pg.ws = s;
    pg.final_last = false;
    return pg;
}

// The unlisted successor of a listed window whose carry-out needs an extension: only the run touching its left
// boundary can print (the carry is consumed by the first segment), and that run is shorter than
// max(T, chars_min_nb * maxlen) bytes because the window is not listed; the event that ends it follows within one
// char.  So the extension pass only reads the first `pre_bytes + 8` bytes of the window.
SX_HD WinGeom ext_geom(const Geometry& geo, int64_t w, uint32_t pre_bytes) {
    WinGeom xg;
    geo.window(w, xg);
    const int64_t lim = xg.ws + (int64_t)pre_bytes + 8;
    if (lim < xg.we) xg.we = lim;
    xg.final_last = false;
    return xg;
}

// A carry-out that can make an UNLISTED successor print something (DESIGN.md "Extension rule").
// General missions (grep_char, same-unicode-block, chars_min_nb > q): a window with >= 2 segments depends on its
// carry-in only through the one flag handed from segment 1 to segment 2, so it is CONSTANT iff forcing that
// flag the other way leaves the carry-out unchanged.  `r1` is the result of the pass under the null carry.
template <class RunFn>
SX_HD uint8_t classify_general(const WinGeom& geo, const WinResult& r1, RunFn&& run_forced) {
    if (geo.final_last) return WT_CONST;
    if (r1.m < 2) return WT_DEP;
    WinGeom g2 = geo;
    g2.force_s2_cont = r1.cut1 ? 0 : 1;
    const Carry o2 = run_forced(g2);
    const Carry& o1 = r1.out;
    const bool same = o1.kind == o2.kind && o1.flags == o2.flags && o1.k == o2.k && o1.in_bytes == o2.in_bytes &&
                      o1.out_bytes == o2.out_bytes && o1.aux == o2.aux;
    return same ? WT_CONST : WT_DEP;
}

// WT_GUARD: the carry-out is the one seen under the null carry unless the carried leftover fills the first run up.
SX_HD bool guard_benign(const ScanParams& P, const WinDesc& d, const Carry& kin) {
    return kin.kind != K_L || (uint32_t)kin.k + d.a < P.q;
}

// A WT_GUARD window behind an adjacent WT_GUARD / WT_CONST window `dp` of the same general mission, carry-in unknown:
// whatever reached `dp`, its carry-out is its null_out, "none" (the leftover killed its segment, helper.rs:410-415) or
// "cut" (the forced cut was the last event, helper.rs:353) -- so if dp.null_out is benign for `d`, every possible
// carry-in is, and d's carry-out is d.null_out.  This is what lets a block find a known carry a few windows back.
SX_HD bool guard_known_behind(const ScanParams& P, const WinDesc& d, const WinDesc& dp) {
    return d.type == WT_GUARD && (dp.type == WT_GUARD || dp.type == WT_CONST) && guard_benign(P, d, dp.null_out);
}

SX_HD bool carry_needs_extension(const ScanParams& P, const Carry& k) {
    return k.kind == K_C || (k.kind == K_L && k.k >= P.n);
}

// Does window `d` emit anything given its real carry-in?  (see DESIGN.md "Emit rule")
SX_HD bool needs_emit(const ScanParams& P, const WinDesc& d, const Carry& kin) {
    if (d.nrec > 0) return true;
    if (P.general) return !carry_is_null(kin);  // no closed form with grep_char / same block: just look
    if (kin.kind == K_C) return d.a > 0;
    return kin.k > 0 && (uint32_t)kin.k + d.a >= P.n;
}

// ------------------------------------------------------------------------------------------
// Text materialisation: decode input[in_start, in_start + in_len) (whole chars only) to UTF-8.
// ------------------------------------------------------------------------------------------
SX_HD uint32_t transcode_range(const ScanParams& P, const GlobalSrc& g, int64_t s, uint32_t len, uint8_t* dst) {
    uint32_t dp = 0;
    const int64_t e = s + len;
    switch (P.enc) {
    case ENC_UTF8:
        for (int64_t p = s; p < e; ++p) dst[dp++] = g.get(p);
        break;
    case ENC_XUD:
        for (int64_t p = s; p < e; ++p) { const uint32_t b = g.get(p); dp += put_utf8(dst + dp, b < 0x80 ? b : b + 0xF700u); }
        break;
    case ENC_SB:
        for (int64_t p = s; p < e; ++p) { const uint32_t b = g.get(p); dp += put_utf8(dst + dp, b < 0x80 ? b : P.sb_table[b - 0x80]); }
        break;
    case ENC_UTF16LE:
    case ENC_UTF16BE: {
        const bool be = P.enc == ENC_UTF16BE;
        uint32_t hs = 0;
        for (int64_t p = s; p + 1 < e; p += 2) {
            const uint32_t b0 = g.get(p), b1 = g.get(p + 1);
            const uint32_t cu = be ? ((b0 << 8) | b1) : ((b1 << 8) | b0);
            if ((cu & 0xFC00) == 0xD800) { hs = cu; continue; }
            if ((cu & 0xFC00) == 0xDC00) { dp += put_utf8(dst + dp, 0x10000u + ((hs - 0xD800u) << 10) + (cu - 0xDC00u)); hs = 0; continue; }
            dp += put_utf8(dst + dp, cu);
        }
        break;
    }
    case ENC_BIG5: {
        DecBig5 d;
        d.lead = 0;
        MbTextEmit te{dst, 0, 0xFFFFFFFFu, 0, 0, false};
        for (int64_t p = s; p < e && !te.stop; ++p) d.step(P, g.get(p), p, te);
        dp = te.dp;
        break;
    }
    case ENC_EUCJP: {
        DecEucJp d;
        d.lead = 0; d.j0212 = 0;
        MbTextEmit te{dst, 0, 0xFFFFFFFFu, 0, 0, false};
        for (int64_t p = s; p < e && !te.stop; ++p) d.step(P, g.get(p), p, te);
        dp = te.dp;
        break;
    }
    case ENC_UTF32LE:
    case ENC_UTF32BE: {
        const bool be = P.enc == ENC_UTF32BE;
        for (int64_t p = s; p + 3 < e; p += 4) {
            const uint32_t b0 = g.get(p), b1 = g.get(p + 1), b2 = g.get(p + 2), b3 = g.get(p + 3);
            const uint32_t c = be ? ((b0 << 24) | (b1 << 16) | (b2 << 8) | b3) : ((b3 << 24) | (b2 << 16) | (b1 << 8) | b0);
            dp += put_utf8(dst + dp, c);
        }
        break;
    }
    }
    return dp;
}



// The text of one record.  Big5: a text may start with the second code point of a two-code-point pair (RF_HALFSTART) or
// end with the first one (then text_len stops the output early); every other text is exactly its input range.
SX_HD uint32_t transcode_record(const ScanParams& P, const GlobalSrc& g, const Record& r, uint8_t* dst) {
    if (P.enc != ENC_BIG5) return transcode_range(P, g, r.in_start, r.in_len, dst);
    DecBig5 d;
    d.lead = 0;
    MbTextEmit te{dst, 0, r.text_len, (r.flags & RF_HALFSTART) ? 1u : 0u, 0, false};
    const int64_t e = r.in_start + r.in_len;
    for (int64_t p = r.in_start; p < e && !te.stop; ++p) d.step(P, g.get(p), p, te);
    return te.dp;
}

// ------------------------------------------------------------------------------------------
// Prefilter (kernel sx_prefilter_kernel): a conservative per-byte "could belong to a passing
// char" flag G at 32-byte-block granularity of the byte value (top 3 bits), cheap enough for
// SWAR.  A window is INTERESTING when a run of >= T good bytes (T = chars_min_nb * unit) can
// exist in it or across its left boundary; only interesting windows, their two neighbours and
// the first/last window of every 256-window tile are handed to the exact kernel.  See
// DESIGN.md "Prefilter" for the proof that every other window emits nothing and does not
// influence any carry.
// ------------------------------------------------------------------------------------------
enum : uint32_t { PF_BYTE = 0, PF_UTF8 = 1, PF_UNIT = 2, PF_PAIR = 3 };
constexpr uint32_t kPrefTileWin = 256;
struct PrefCfg {
    uint32_t enabled;
    uint32_t family;  // PF_*
    uint32_t blkA;    // bit k: bytes of block k (k = byte >> 5, k < 4) may be passing ASCII
    uint32_t blkH;    // PF_BYTE: blocks 4..7 that may map to passing chars; PF_UTF8: lead blocks 6,7 that may pass;
                      // PF_UNIT: blocks of the most significant unit byte that may belong to a passing char
    uint32_t multi;   // PF_UTF8: 3/4-byte leads may pass (a continuation byte may follow a continuation byte)
    uint32_t T;       // run threshold in bytes
    uint32_t unit;    // PF_UNIT: 2 or 4
    uint32_t hi_pos;  // PF_UNIT: offset of the most significant byte inside a unit
    uint32_t n_chars; // chars_min_nb
    uint32_t refine;  // PF_UTF8: a long good-byte run only counts if it holds >= n_chars non-continuation bytes
    uint32_t pre_bytes;  // pre-roll length: longest possible trailing good run of an uninteresting window + slack
    uint32_t kill_trail; // --grep-char missions: a window behind >= this many trailing good bytes is listed (0: rule off)
    uint32_t sb_rule;    // --same-unicode-block missions: a window is listed when its own trailing good run and the one of
                         // the window before it may both hold a multi-byte char (PrefWin::trail_hi), see make_pref_cfg
    // PF_PAIR (Big5, EUC-JP): per byte value, bit 0: a passing single-byte char, bit 1: may be the lead (or, EUC-JP, the
    // middle byte) of a passing multi-byte char, bit 2: may be a trail byte.  A byte is good when it is a passing
    // single-byte char, a lead candidate followed by a trail candidate, or a trail candidate behind a lead candidate --
    // whatever the real parse is, every byte of a passing char is good.
    uint8_t pair_cls[256];
};

// Reference (byte-wise) definition of G; the SWAR kernel must produce exactly these flags.
template <class S>
SX_HD bool pref_good(const ScanParams& P, const PrefCfg& c, const S& src, int64_t i, int64_t ws, int64_t we) {
    const uint32_t b = src.get(i);
    if (c.family == PF_BYTE) return (((b < 0x80 ? c.blkA : c.blkH) >> (b >> 5)) & 1u) != 0;
    if (c.family == PF_UTF8) {
        if (b < 0x80) return ((c.blkA >> (b >> 5)) & 1u) != 0;
        const bool is_cn = b < 0xC0;
        if (!is_cn) {  // lead candidate: good when its block may pass and a continuation byte follows
            if (!((c.blkH >> (b >> 5)) & 1u)) return false;
            if (i + 1 >= we) return true;
            const uint32_t nb = src.get(i + 1);
            return nb >= 0x80 && nb < 0xC0;
        }
        if (i - 1 < ws) return true;
        const uint32_t pb = src.get(i - 1);
        if (pb >= 0xC0) return ((c.blkH >> (pb >> 5)) & 1u) != 0;
        return c.multi && pb >= 0x80;
    }
    if (c.family == PF_PAIR) {
        const uint32_t k = c.pair_cls[b];
        if (k & 1u) return true;
        if ((k & 2u) && (i + 1 >= we || (c.pair_cls[src.get(i + 1)] & 4u))) return true;
        if ((k & 4u) && (i - 1 < ws || (c.pair_cls[src.get(i - 1)] & 2u))) return true;
        return false;
    }
    // PF_UNIT
    int64_t rel = i - (int64_t)P.align;
    int64_t k = rel >= 0 ? rel / (int64_t)c.unit : -((-rel + (int64_t)c.unit - 1) / (int64_t)c.unit);
    const int64_t t = (int64_t)P.align + k * (int64_t)c.unit + (int64_t)c.hi_pos;
    if (t < ws || t >= we) return true;
    return ((c.blkH >> (src.get(t) >> 5)) & 1u) != 0;
}

struct PrefWin { uint32_t lead, trail, maxrun, maxchars, kill, trail_hi, lead_hi; };  // trail_hi, lead_hi: see PrefCfg::sb_rule  // kill: see PrefCfg::kill_trail  // maxchars: most non-continuation bytes in a run of >= T bytes
// General missions that may use the prefilter: --grep-char and --same-unicode-block, each with one more listing rule
// (make_pref_cfg: kill_trail, sb_rule).  chars_min_nb > q drops whole segments (helper.rs:410-415): every window scanned.
SX_HD bool pref_general_ok(const ScanParams& P) {
    return P.n <= P.q && (P.grep_char >= 0 || P.same_block);
}

template <class S>
SX_HD PrefWin pref_window_ref(const ScanParams& P, const PrefCfg& c, const S& src, int64_t ws, int64_t we) {
    PrefWin r{0, 0, 0, 0, 0, 0, 0};
    uint32_t run = 0, chars = 0;
    bool seen_bad = false, run_hi = false;
    auto close_run = [&]() { if (run >= c.T && chars > r.maxchars) r.maxchars = chars; };
    for (int64_t i = ws; i < we; ++i) {
        if (pref_good(P, c, src, i, ws, we)) {
            run++;
            const uint32_t b = src.get(i);
            if (!(b >= 0x80 && b < 0xC0)) chars++;
            if (b >= 0x80 || c.family == PF_UNIT) run_hi = true;  // may belong to a char outside ASCII
            if (run > r.maxrun) r.maxrun = run;
            // a good run of >= kill_trail bytes, or one that began at (before) the window's first byte, reaching one of
            // the last 4 bytes: the decoded text may end in >= q passing chars (up to 3 bytes stay pending in the decoder)
            if (c.kill_trail != 0 && i + 4 >= we && (run >= c.kill_trail || !seen_bad)) r.kill = 1;
        } else {
            close_run();
            if (!seen_bad) { r.lead = run; r.lead_hi = (run > 0 && run_hi) ? 1u : 0u; seen_bad = true; }
            run = 0; chars = 0; run_hi = false;
        }
    }
    close_run();
    if (!seen_bad) r.lead = run;
    r.trail = run;
    // the trailing good run may hold a char outside ASCII -- or reaches back beyond the window, where it may
    r.trail_hi = (run > 0 && (run_hi || !seen_bad)) ? 1u : 0u;
    return r;
}


// Host-side: derive the prefilter configuration of a mission (used by the C ABI and by the
// test harness, so both classify identically).
// mb_a / mb_b: HOST copies of the Big5 / EUC-JP index tables (P.mb_a / P.mb_b are device pointers in the product).
inline PrefCfg make_pref_cfg(const ScanParams& P, bool input_16b_aligned, const uint32_t* mb_a = nullptr, const uint32_t* mb_b = nullptr) {
    PrefCfg c;
    c.enabled = 0; c.family = PF_BYTE; c.blkA = 0; c.blkH = 0; c.multi = 0; c.T = P.n; c.unit = 1; c.hi_pos = 0;
    c.n_chars = P.n; c.refine = 0; c.pre_bytes = 0; c.kill_trail = 0; c.sb_rule = 0;
    for (uint32_t k = 0; k < 256; ++k) c.pair_cls[k] = 0;
    for (uint32_t k = 0; k < 4; ++k) {
        const uint64_t word = k < 2 ? P.af_lo : P.af_hi;
        if ((word >> ((k & 1) * 32)) & 0xFFFFFFFFull) c.blkA |= 1u << k;
    }
    auto pass_lead = [&](uint32_t lead) { return ((P.ubf >> (lead & 0x3f)) & 1ull) != 0; };
    auto pass_cp = [&](uint32_t cp) { return cp < 0x80 ? pass_filter(P, cp) : pass_lead(utf8_lead_of_cp(cp)); };
    switch (P.enc) {
    case ENC_XUD:
        c.family = PF_BYTE;
        c.blkH = pass_lead(0xEF) ? 0xF0u : 0u;
        break;
    case ENC_SB:
        c.family = PF_BYTE;
        for (uint32_t b = 0x80; b < 0x100; ++b) {
            const uint32_t cp = P.sb_table[b - 0x80];
            if (cp != 0 && pass_cp(cp)) c.blkH |= 1u << (b >> 5);
        }
        break;
    case ENC_UTF8:
        c.family = PF_UTF8;
        if (P.ubf & 0xFFFFFFFFull) c.blkH |= 1u << 6;
        if (P.ubf >> 32) { c.blkH |= 1u << 7; c.multi = 1; }
        break;
    case ENC_UTF16LE:
    case ENC_UTF16BE: {
        c.family = PF_UNIT; c.unit = 2; c.hi_pos = P.enc == ENC_UTF16LE ? 1 : 0;
        const bool astral = ((P.ubf >> 48) & 0x1Full) != 0;
        for (uint32_t hi = 0; hi < 0x100; ++hi) {
            bool good = false;
            if (hi >= 0xD8 && hi <= 0xDF) good = astral;
            else for (uint32_t lo = 0; lo < 0x100 && !good; ++lo) good = pass_cp((hi << 8) | lo);
            if (good) c.blkH |= 1u << (hi >> 5);
        }
        break;
    }
    case ENC_UTF32LE:
    case ENC_UTF32BE:
        c.family = PF_UNIT; c.unit = 4; c.hi_pos = P.enc == ENC_UTF32LE ? 3 : 0;
        c.blkH = (P.af_lo | P.af_hi | P.ubf) ? 1u : 0u;
        break;
    case ENC_BIG5:
        c.family = PF_PAIR;
        for (uint32_t b = 0; b < 0x80; ++b) c.pair_cls[b] = pass_filter(P, b) ? 1u : 0u;
        for (uint32_t b = 0x40; b <= 0xFE; ++b) if (big5_is_trail(b)) c.pair_cls[b] |= 4u;
        for (uint32_t lead = 0x81; lead <= 0xFE && mb_a; ++lead)
            for (uint32_t t = 0x40; t <= 0xFE; ++t) {
                if (!big5_is_trail(t)) continue;
                const uint32_t ptr = big5_pointer(lead, t);
                const bool ok = big5_is_double(ptr) ? (pass_lead(0xC3) || pass_lead(0xCC)) : (mb_a[ptr] != 0 && pass_cp(mb_a[ptr]));
                if (ok) { c.pair_cls[lead] |= 2u; break; }
            }
        break;
    case ENC_EUCJP:
        c.family = PF_PAIR;
        for (uint32_t b = 0; b < 0x80; ++b) c.pair_cls[b] = pass_filter(P, b) ? 1u : 0u;
        for (uint32_t b = 0xA1; b <= 0xFE; ++b) c.pair_cls[b] |= 4u;
        if (pass_lead(0xEF)) c.pair_cls[0x8E] |= 2u;  // half-width katakana U+FF61..U+FF9F
        for (uint32_t lead = 0xA1; lead <= 0xFE && mb_a && mb_b; ++lead) {
            bool ok8 = false, ok12 = false;
            for (uint32_t t = 0; t < 94; ++t) {
                const uint32_t c8 = mb_a[(lead - 0xA1u) * 94u + t], c12 = mb_b[(lead - 0xA1u) * 94u + t];
                ok8 = ok8 || (c8 != 0 && pass_cp(c8));
                ok12 = ok12 || (c12 != 0 && pass_cp(c12));
            }
            if (ok8 || ok12) c.pair_cls[lead] |= 2u;      // lead of a jis0208 char / middle byte of a jis0212 char
            if (ok12) c.pair_cls[0x8F] |= 2u;
        }
        break;
    }
    c.T = P.n * c.unit;
    if (c.T > P.W) c.T = P.W;  // a run covering a whole window is always interesting
    // UTF-8: a run of good bytes holds at most one char per non-continuation byte, so a long byte run with
    // fewer than n such bytes cannot hold n chars.  An uninteresting window can then end in a good run of up
    // to (n-1) chars = (n-1) * maxlen bytes, which the pre-roll has to cover.
    c.refine = (c.family == PF_UTF8) ? 1u : 0u;
    {
        const uint32_t maxlen = c.family == PF_UTF8 ? (c.multi ? 4u : 2u) : 1u;
        c.pre_bytes = (c.refine ? P.n * maxlen : c.T) + 3 + c.unit;
    }
    c.enabled = (P.W % 16 == 0 && P.W <= 128 && P.slice_len % P.W == 0 && input_16b_aligned) ? 1u : 0u;
    // --grep-char (chars_min_nb <= q): a leftover of q chars is possible (helper.rs:389-392)
    // and is evaluated at the next window's first char -- dropped together with the rest of that segment if it lacks the
    // grep char (helper.rs:410-415), printed as a "maybe cut" string otherwise -- so the window behind >= q good chars
    // (>= q * unit bytes) has to be scanned exactly even if it holds no run itself.  With that window listed, no unlisted
    // window ever sees its first run filled up to q chars (trail + lead >= q * unit >= T lists it), and past its first
    // run an unlisted window holds nothing of its carry-in.  chars_min_nb > q: no prefilter (pref_general_ok).
    if (P.grep_char >= 0 && P.n <= P.q) c.kill_trail = P.q * c.unit;
    // --same-unicode-block (chars_min_nb <= q; with --grep-char both rules apply): SplitStr keeps the lead byte of the last multi-byte
    // char it saw across failing ASCII chars and short runs (helper.rs:221, :287-292, :327-330), and a leftover hands its
    // own to the next window.  An unlisted window prints nothing (its runs are pieces of the plain mission's runs), but
    // such a stale lead byte decides where its TRAILING run -- the carry-out -- begins: "abcΓΔ|" is kept whole behind a
    // clean start and as "ΓΔ" behind a Cyrillic char.  That needs a multi-byte char in the leftover coming in (the trailing
    // good run of the window before holds a byte >= 0x80, or covers that whole window) -- or one in the window's leading
    // run, which a "cut" carry prints (the `break` forgets the lead byte, helper.rs:315-330) and any other carry drops
    // (the lead byte stays); the cut flag survives junk, so the window before says nothing about it -- and one in the
    // window's own trailing run: such a window is listed.  Every other unlisted window has a carry-out that depends on its own bytes
    // only -- on ALL of them (the stale byte may stem from the window's first chars), so the pre-roll of a head is the
    // whole window before it.
    if (P.same_block && P.n <= P.q) { c.sb_rule = 1; c.pre_bytes = P.W; }
    return c;
}

// Reference INTERESTING / E classification for window `w` of a stream (spec of the SWAR kernel).
// forced: first and last window of the stream, windows shorter than W, the is_last final window.
struct PrefPlan {
    int64_t total_windows;
};
template <class S>
SX_HD bool pref_interesting_ref(const ScanParams& P, const PrefCfg& c, const Geometry& geo, const S& src, int64_t w,
                                int64_t total_windows) {
    WinGeom g;
    if (!geo.window(w, g)) return false;
    if (w == 0 || w == total_windows - 1 || (uint32_t)(g.we - g.ws) < P.W || g.final_last) return true;
    const PrefWin a = pref_window_ref(P, c, src, g.ws, g.we);
    if (c.refine ? (a.maxrun >= c.T && a.maxchars >= c.n_chars) : (a.maxrun >= c.T)) return true;
    if ((w % kPrefTileWin) == 0) return a.lead >= 1 || c.kill_trail != 0 || (c.sb_rule && a.trail_hi);  // previous window unknown to the tile
    WinGeom gp;
    geo.window(w - 1, gp);
    const PrefWin b = pref_window_ref(P, c, src, gp.ws, gp.we);
    if (b.kill) return true;
    if (c.sb_rule && a.trail_hi && (b.trail_hi || a.lead_hi)) return true;
    // A run of >= T good bytes touches the window's left boundary.  This includes a run that only starts at the
    // window's first byte (b.trail == 0): whatever its char count, it is the run a "cut" carry would complete,
    // and listing it keeps the carry-out of every UNLISTED window independent of its carry-in.
    return a.lead >= 1 && b.trail + a.lead >= c.T;
}

}  // namespace sx
