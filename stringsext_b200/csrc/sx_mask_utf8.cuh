// sx_mask_utf8.cuh -- bit-parallel window engine for UTF-8 missions (the common case of the exact kernel).
//
// Same semantics as scan_window_fast_utf8 / scan_window<DecUtf8> (FindingCollection::from + SplitStr::next,
// /root/reference/src/finding_collection.rs:84-342, /root/reference/src/helper.rs:210-432, decoder = WHATWG
// UTF-8 as restated in DecUtf8), but without a byte loop: the window (<= 128 bytes plus 3 bytes of look-back)
// is turned into bit masks -- one bit per byte -- of the decoder events, and the few runs that can print are
// located with bit scans.
//
//   class planes   every byte -> utf8_class (4 bits) + filter verdict (1 bit) through a 256-entry table, the five
//                  flag bits of 32 bytes gathered into five mask words (dp4a on the device)
//   decoder        a lead byte ALWAYS starts a sequence and a non-continuation byte ALWAYS ends one, so whether
//                  a continuation byte is accepted depends on the <= 3 preceding bytes only:
//                    ok1 = Cn & (range allowed by the lead one byte back), ok2 = Cn & ok1<<1 & len>=3 two back, ...
//                  malformed events: X bytes, continuation bytes that are not accepted (`mal`: the next segment
//                  starts after the byte) and bytes that arrive while a sequence is pending without being accepted
//                  (`pre`: the next segment starts AT the byte, which is read again)
//   runs           bytes of passing chars form maximal runs; a run prints iff it holds >= chars_min_nb chars and
//                  ends inside the window; its position is the last segment start at or before it
//
// Everything unusual returns false and the caller takes the byte-wise engine: a carry-in other than a short
// leftover, a run that could reach output_line_char_nb_max chars, a run covering the whole window, the last
// window of a flushed stream, a finding whose precision depends on the Precision::Before probe.
// tests/emul runs this code on the CPU against the oracle (it is tried first by WindowEngine<DecUtf8>).
#pragma once
#include "sx_fast_utf8.cuh"
#include "sx_fast_generic.cuh"

namespace sx {

template <int NW> struct MK { uint32_t w[NW + 1]; };  // bit 32 + p of the mask = byte p of the window, p in -32 .. 32*NW-1
// (NW = data words: 4 for a full 128-byte window, 2 for the short pre-roll / extension windows)

SX_HD uint32_t sx_fsl(uint32_t lo, uint32_t hi, uint32_t k) {  // (hi:lo) << k, upper word; 1 <= k <= 31
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, k);
#else
    return (hi << k) | (lo >> (32 - k));
#endif
}
SX_HD uint32_t sx_fsr(uint32_t lo, uint32_t hi, uint32_t k) {  // (hi:lo) >> k, lower word; 1 <= k <= 31
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, k);
#else
    return (lo >> k) | (hi << (32 - k));
#endif
}
SX_HD uint32_t sx_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
SX_HD uint32_t sx_clz(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz(x);
#else
    return (uint32_t)__builtin_clz(x);
#endif
}
SX_HD uint32_t sx_ctz(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs(x) - 1);
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

template <int K, int NW> SX_HD MK<NW> m5_shl(const MK<NW>& a) {  // towards higher byte positions
    MK<NW> r;
    r.w[0] = a.w[0] << K;
#pragma unroll
    for (int i = 1; i < NW + 1; ++i) r.w[i] = sx_fsl(a.w[i - 1], a.w[i], K);
    return r;
}
template <int K, int NW> SX_HD MK<NW> m5_shr(const MK<NW>& a) {
    MK<NW> r;
#pragma unroll
    for (int i = 0; i < NW; ++i) r.w[i] = sx_fsr(a.w[i], a.w[i + 1], K);
    r.w[NW] = a.w[NW] >> K;
    return r;
}
template <int NW> SX_HD bool m5_bit(const MK<NW>& a, uint32_t B) {  // static-index only (no local memory)
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < NW + 1; ++i) if ((B >> 5) == (uint32_t)i) v = a.w[i];
    return ((v >> (B & 31)) & 1u) != 0;
}
// highest set bit at index <= B, or -1
template <int NW> SX_HD int32_t m5_high_le(const MK<NW>& a, uint32_t B) {
    int32_t r = -1;
#pragma unroll
    for (int i = 0; i < NW + 1; ++i) {
        uint32_t v = a.w[i];
        const uint32_t lo = (uint32_t)i * 32u;
        if (B < lo) v = 0;
        else if (B < lo + 31u) v &= (2u << (B - lo)) - 1u;
        if (v) r = (int32_t)(lo + 31u - sx_clz(v));
    }
    return r;
}
// lowest set bit at index >= B, or 32 * (NW + 1)
template <int NW> SX_HD uint32_t m5_low_ge(const MK<NW>& a, uint32_t B) {
    uint32_t r = 32 * (NW + 1);
#pragma unroll
    for (int i = NW; i >= 0; --i) {
        uint32_t v = a.w[i];
        const uint32_t lo = (uint32_t)i * 32u;
        if (B >= lo + 32u) v = 0;
        else if (B > lo) v &= ~((1u << (B - lo)) - 1u);
        if (v) r = lo + sx_ctz(v);
    }
    return r;
}
// set bits with index in [A, B] (inclusive)
template <int NW> SX_HD uint32_t m5_count(const MK<NW>& a, uint32_t A, uint32_t B) {
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < NW + 1; ++i) {
        uint32_t v = a.w[i];
        const uint32_t lo = (uint32_t)i * 32u;
        if (B < lo || A >= lo + 32u) v = 0;
        else {
            if (A > lo) v &= ~((1u << (A - lo)) - 1u);
            if (B < lo + 31u) v &= (2u << (B - lo)) - 1u;
        }
        n += sx_popc(v);
    }
    return n;
}

// index of the j-th (j >= 1) set bit at index >= B; the caller guarantees it exists
template <int NW> SX_HD uint32_t m5_select(const MK<NW>& a, uint32_t B, uint32_t j) {
    uint32_t r = 32 * (NW + 1);
    bool found = false;
#pragma unroll
    for (int i = 0; i < NW + 1; ++i) {
        uint32_t v = a.w[i];
        const uint32_t lo = (uint32_t)i * 32u;
        if (B >= lo + 32u) v = 0;
        else if (B > lo) v &= ~((1u << (B - lo)) - 1u);
        const uint32_t c = sx_popc(v);
        if (!found) {
            if (j <= c) {
#if defined(__CUDA_ARCH__)
                r = lo + __fns(v, 0, (int)j);
#else
                uint32_t t = v;
                for (uint32_t k = 1; k < j; ++k) t &= t - 1;
                r = lo + sx_ctz(t);
#endif
                found = true;
            } else j -= c;
        }
    }
    return r;
}

// The five class planes of a window.  TileSrc: load_chunk(r16, ws, we) (16 bytes, zero outside [ws, we)),
// cls(b) (table lookup), get(off).
// LB32: word 0 holds all 32 bytes before the window (requires ws >= 32), otherwise only the last three.
template <int NW, bool LB32, class TileSrc>
SX_HD void utf8_class_planes(const ScanParams& P, const TileSrc& tsrc, int64_t ws, int64_t we, MK<NW>* pl) {
#pragma unroll
    for (int t = 0; t < 5; ++t) pl[t].w[0] = 0;
    if (!LB32) {
        // look-back: the three bytes before the window (bytes before the stream start leave the decoder neutral)
#pragma unroll
        for (int j = 1; j <= 3; ++j) {
            const int64_t o = ws - j;
            if (o >= -(int64_t)P.npend) {
                const uint32_t c = tsrc.cls(tsrc.get(o));
#pragma unroll
                for (int t = 0; t < 5; ++t) pl[t].w[0] |= ((c >> t) & 1u) << (32 - j);
            }
        }
    }
    const int64_t lo_bound = LB32 ? ws - 32 : ws;
#pragma unroll
    for (int grp = LB32 ? -1 : 0; grp < NW; ++grp) {
#if defined(__CUDA_ARCH__)
        uint32_t acc[5][4];
#pragma unroll
        for (int t = 0; t < 5; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0; }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t r16 = ws + (int64_t)(grp * 32 + h * 16);
            if (r16 < we) {
                const uint4 v = tsrc.load_chunk(r16, lo_bound, we);
                const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t x = xs[j];
                    const uint32_t cw = tsrc.cls(x & 0xFFu) | (tsrc.cls((x >> 8) & 0xFFu) << 8) | (tsrc.cls((x >> 16) & 0xFFu) << 16) |
                                        (tsrc.cls(x >> 24) << 24);
                    const int k = 4 * h + j;
                    const uint32_t wt = (k & 1) ? 0x80402010u : 0x08040201u;
#pragma unroll
                    for (int t = 0; t < 5; ++t) {
                        uint32_t r;
                        asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(cw & (0x01010101u << t)), "r"(wt), "r"(acc[t][k >> 1]));
                        acc[t][k >> 1] = r;  // 8 flags at bits t .. t+7
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 5; ++t)
            pl[t].w[grp + 1] = ((acc[t][0] | (acc[t][1] << 8) | (acc[t][2] << 16)) >> t) | (acc[t][3] << (24 - t));
#else
#pragma unroll
        for (int t = 0; t < 5; ++t) pl[t].w[grp + 1] = 0;
        for (int i = 0; i < 32; ++i) {
            const int64_t o = ws + grp * 32 + i;
            if (o >= we) break;
            const uint32_t c = tsrc.cls(tsrc.get(o));
            for (int t = 0; t < 5; ++t) pl[t].w[grp + 1] |= ((c >> t) & 1u) << i;
        }
#endif
    }
}

struct MaskEmit {
    const ScanParams* P;
    int64_t base;
    int mode;
    Record* wr;
    uint64_t text_off;
    uint32_t nrec, ntext;
};
SX_HD void mask_emit(MaskEmit& E, int32_t seg_rel, uint32_t prec, int32_t run_s, int32_t run_e, uint32_t text_len, uint32_t flags) {
    const uint32_t len = (uint32_t)(run_e - run_s);
    if (E.mode == MODE_WRITE || (E.mode == MODE_BUFFER && E.nrec < kBufRecs)) {
        Record r;
        r.position = E.P->base_consumed + (uint64_t)(E.base + seg_rel);  // finding_collection.rs:260
        r.in_start = E.base + run_s;
        r.in_len = len;
        r.text_len = text_len;
        r.text_off = E.text_off;
        r.flags = flags;
        r.precision = prec;
        *E.wr++ = r;
        E.text_off += text_len;
    }
    E.nrec++;
    E.ntext += text_len;
}

// Returns false when the window needs the byte-wise engine (nothing has been written in that case that the
// byte-wise engine would not overwrite).
// LB32: the carry-in is NOT given but derived from the 32 bytes before the window (the pre-roll of a head folded into
// the same pass): valid when the caller knows that the run touching the window's left boundary is shorter than 28
// bytes (the predecessor window is unlisted and pre_bytes <= 28); `kin_arg` is ignored, res.in tells what was derived.
// SBYTE: single-byte family (x-user-defined, table driven): no decoder state, a char is a byte, the class table gives
// verdict / unmapped / UTF-8 length of the mapped char; everything after the event masks is shared with UTF-8.
template <int NW, bool LB32, bool SBYTE, class TileSrc>
SX_HD_NOINLINE bool mask_window_nw(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, const Carry& kin_arg, int mode,
                            Record* wr, uint64_t text_off, WinResult& res) {
    const int64_t ws = geo.ws, we = geo.we;
    const int32_t wlen = (int32_t)(we - ws);
    const uint32_t n = P.n, q = P.q;
    if (wlen < 4 || wlen > 32 * NW || (ws & 15) != 0 || geo.final_last) return false;
    if (LB32 && ws < 32) return false;
    Carry kin = LB32 ? carry_none() : kin_arg;
    if (kin.kind == K_UNKNOWN) return false;
    bool kc = kin.kind == K_C;              // the first run completes a cut finding whatever its length
    uint32_t k_in = kc ? 0u : kin.k;        // chars of the leftover the first run continues

    using M5 = MK<NW>;
    M5 pl[5];
    utf8_class_planes<NW, LB32>(P, tsrc, ws, we, pl);
    // positions >= wlen: class 0 (ASCII), verdict 0 (zero fill read through the table may say otherwise)
    M5 V;  // valid window positions
    V.w[0] = 0;
#pragma unroll
    for (int i = 1; i < NW + 1; ++i) {
        const int32_t lo = (i - 1) * 32;
        V.w[i] = wlen >= lo + 32 ? 0xFFFFFFFFu : (wlen > lo ? ((1u << (wlen - lo)) - 1u) : 0u);
    }
#pragma unroll
    for (int t = 0; t < 5; ++t)
#pragma unroll
        for (int i = 1; i < NW + 1; ++i) pl[t].w[i] &= V.w[i];

    // event masks shared by both families: acc / pendm (UTF-8 decoder), pre / seg (segment starts), pe (passing char
    // ends), R (bytes of passing chars), lead (UTF-8 lead bytes), TL0 / TL1 (single-byte: UTF-8 length - 1 of the char)
    M5 acc, pendm, pre, seg, pe, R, lead, TL0, TL1, mal;
    if (SBYTE) {
        // class table bits: 0 verdict, 1 unmapped (malformed, DecSb), 2-3 UTF-8 length - 1 of the mapped char
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) {
            const uint32_t pv = (LB32 && i == 0) ? 0xFFFFFFFFu : V.w[i];
            R.w[i] = pl[0].w[i] & pv;
            pe.w[i] = R.w[i];
            mal.w[i] = i == 0 ? 0u : pl[1].w[i];
            TL0.w[i] = pl[2].w[i];
            TL1.w[i] = pl[3].w[i];
            acc.w[i] = 0; pendm.w[i] = 0; pre.w[i] = 0; lead.w[i] = 0;
        }
        const M5 s1mal = m5_shl<1>(mal);
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) seg.w[i] = s1mal.w[i] & V.w[i];
    } else {
        // ---- class masks (utf8_class: 0 A, 1 80-8F, 2 90-9F, 3 A0-BF, 4 X, 5 L2, 6 E0, 7 E1-EC/EE/EF, 8 ED, 9 F0, 10 F1-F3, 11 F4)
        M5 Cn, C80, C90, CA0, X, L2, plain, E0, ED, F0, F4, len34, len4, PS;
    #pragma unroll
        for (int i = 0; i < NW + 1; ++i) {
            const uint32_t c0 = pl[0].w[i], c1 = pl[1].w[i], c2 = pl[2].w[i], c3 = pl[3].w[i];
            const uint32_t hi0 = ~c3 & ~c2;  // classes 0..3
            C80.w[i] = hi0 & ~c1 & c0;
            C90.w[i] = hi0 & c1 & ~c0;
            CA0.w[i] = hi0 & c1 & c0;
            Cn.w[i] = hi0 & (c1 | c0);
            X.w[i] = ~c3 & c2 & ~c1 & ~c0;                       // 4
            L2.w[i] = ~c3 & c2 & ~c1 & c0;                       // 5
            E0.w[i] = ~c3 & c2 & c1 & ~c0;                       // 6
            const uint32_t l3n = ~c3 & c2 & c1 & c0;             // 7
            ED.w[i] = c3 & ~c2 & ~c1 & ~c0;                      // 8
            F0.w[i] = c3 & ~c2 & ~c1 & c0;                       // 9
            const uint32_t l4n = c3 & ~c2 & c1 & ~c0;            // 10
            F4.w[i] = c3 & ~c2 & c1 & c0;                        // 11
            plain.w[i] = L2.w[i] | l3n | l4n;
            len4.w[i] = F0.w[i] | l4n | F4.w[i];
            len34.w[i] = E0.w[i] | l3n | ED.w[i] | len4.w[i];
            PS.w[i] = pl[4].w[i];
        }
        // ---- decoder: accepted continuation bytes -----------------------------------------------------------------
        M5 ok1, ok2, ok3;
        {
            const M5 s_plain = m5_shl<1>(plain), s_e0 = m5_shl<1>(E0), s_ed = m5_shl<1>(ED), s_f0 = m5_shl<1>(F0), s_f4 = m5_shl<1>(F4);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i)
                ok1.w[i] = Cn.w[i] & (s_plain.w[i] | (s_e0.w[i] & CA0.w[i]) | (s_ed.w[i] & (C80.w[i] | C90.w[i])) |
                                      (s_f0.w[i] & (C90.w[i] | CA0.w[i])) | (s_f4.w[i] & C80.w[i]));
            const M5 s1ok1 = m5_shl<1>(ok1), s2l34 = m5_shl<2>(len34);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) ok2.w[i] = Cn.w[i] & s1ok1.w[i] & s2l34.w[i];
            const M5 s1ok2 = m5_shl<1>(ok2), s3l4 = m5_shl<3>(len4);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) ok3.w[i] = Cn.w[i] & s1ok2.w[i] & s3l4.w[i];
            const M5 s1l34 = m5_shl<1>(len34), s2l4 = m5_shl<2>(len4);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) {
                acc.w[i] = ok1.w[i] | ok2.w[i] | ok3.w[i];
                // a sequence is pending AFTER this byte
                pendm.w[i] = L2.w[i] | len34.w[i] | (ok1.w[i] & s1l34.w[i]) | (ok2.w[i] & s2l4.w[i]);
                mal.w[i] = X.w[i] | (Cn.w[i] & ~acc.w[i]);
            }
            mal.w[0] = 0;  // what happened before the window only matters through the pending sequence
            const M5 s1pend = m5_shl<1>(pendm), s1mal = m5_shl<1>(mal);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) {
                pre.w[i] = s1pend.w[i] & ~acc.w[i] & V.w[i];
                seg.w[i] = (pre.w[i] | s1mal.w[i]) & V.w[i];  // a segment starts at this byte
            }
        }
        // ---- char events (at the last byte of the char) and the bytes of passing chars --------------------------------
        {
            const M5 s1L2 = m5_shl<1>(L2), s2l34 = m5_shl<2>(len34), s2l4 = m5_shl<2>(len4);
            const M5 s1ps = m5_shl<1>(PS), s2ps = m5_shl<2>(PS), s3ps = m5_shl<3>(PS);
            M5 p1, p2, p3, p4;
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) {
                const uint32_t a = ~(pl[0].w[i] | pl[1].w[i] | pl[2].w[i] | pl[3].w[i]);  // class 0
                const uint32_t pv = (LB32 && i == 0) ? 0xFFFFFFFFu : V.w[i];  // LB32: the chars before the window count too
                p1.w[i] = a & PS.w[i] & pv;
                p2.w[i] = ok1.w[i] & s1L2.w[i] & s1ps.w[i] & pv;
                p3.w[i] = ok2.w[i] & s2l34.w[i] & ~s2l4.w[i] & s2ps.w[i] & pv;
                p4.w[i] = ok3.w[i] & s3ps.w[i] & pv;
                pe.w[i] = p1.w[i] | p2.w[i] | p3.w[i] | p4.w[i];
            }
            const M5 r2 = m5_shr<1>(p2), r3a = m5_shr<1>(p3), r3b = m5_shr<2>(p3), r4a = m5_shr<1>(p4), r4b = m5_shr<2>(p4), r4c = m5_shr<3>(p4);
    #pragma unroll
            for (int i = 0; i < NW + 1; ++i) R.w[i] = pe.w[i] | r2.w[i] | r3a.w[i] | r3b.w[i] | r4a.w[i] | r4b.w[i] | r4c.w[i];
        }
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) { lead.w[i] = L2.w[i] | len34.w[i]; TL0.w[i] = 0; TL1.w[i] = 0; }
    }
    // text bytes (UTF-8) of the chars whose bytes are the mask bits [a, b]
    auto T = [&](uint32_t a, uint32_t b) -> uint32_t {
        return SBYTE ? (b + 1u - a) + m5_count(TL0, a, b) + 2u * m5_count(TL1, a, b) : (b + 1u - a);
    };
    // bytes inside the decoder at the window start (the straddling char, if any, starts there)
    int32_t pend0 = 0;
    if (!SBYTE && m5_bit(pendm, 31)) pend0 = m5_bit(lead, 31) ? 1 : (m5_bit(lead, 30) ? 2 : 3);
    if (LB32) {
        // the carry-in: the run of complete passing chars that ends where the pending bytes (if any) begin
        const uint32_t lastB = 31u - (uint32_t)pend0;
        if ((R.w[0] >> lastB) & 1u) {
            const uint32_t inv = ~(R.w[0] << (31u - lastB));
            const uint32_t len = inv ? sx_clz(inv) : 32u;
            if (len > lastB + 1u - 4u) return false;  // reaches the first bytes of the frame: the pre-roll must decide
            const uint32_t ls = lastB + 1u - len;
            const uint32_t k = sx_popc(pe.w[0] & (((1u << len) - 1u) << ls));
            if (k == 0) return false;
            kin.kind = K_L; kin.flags = 0; kin.k = (uint16_t)k;
            kin.in_bytes = 32u - ls;
            kin.out_bytes = T(ls, lastB);
            kin.aux = 0;
            k_in = k;
        }
        pe.w[0] = 0;
        R.w[0] &= pend0 ? (0xFFFFFFFFu << (32 - pend0)) : 0u;  // only the straddling char's bytes belong to the window's run
    }
    res.in = kin;
    // bytes still inside the decoder at the window end
    const uint32_t Blast = 31u + (uint32_t)wlen;
    int32_t npend_out = 0;
    if (!SBYTE && m5_bit(pendm, Blast)) {
        npend_out = m5_bit(lead, Blast) ? 1 : (m5_bit(lead, Blast - 1) ? 2 : 3);
        if (npend_out > wlen) return false;
    }

    MaskEmit E;
    E.P = &P; E.base = ws; E.mode = mode; E.wr = wr; E.text_off = text_off; E.nrec = 0; E.ntext = 0;
    const bool at_slice_start = geo.slice_start == ws;
    // run boundaries
    M5 RS, RE;  // first / last byte of every run
    {
        const M5 up = m5_shl<1>(R), dn = m5_shr<1>(R);
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) { RS.w[i] = R.w[i] & ~up.w[i]; RE.w[i] = R.w[i] & ~dn.w[i]; }
    }
    const uint32_t Bend = 32u + (uint32_t)wlen - (uint32_t)npend_out;  // one past the last complete char
    int32_t last_seg = -2;  // segment of the last yield (-1: the window's first segment)
    uint32_t next_B = 32;   // runs starting below this index are done

    // The Precision::Before probe (finding_collection.rs:176-207) can only change the precision of a finding of a
    // segment that starts at the slice start; with no leftover and a neutral decoder it changes nothing.
    // First segment of a slice with bytes pending from the previous slice and no leftover: the probe fires at the
    // segment's first char if that char is multi-byte -- i.e. iff the pending sequence completes (a broken one starts
    // a second segment at offset 0 instead) -- and then says Before (finding_collection.rs:183-207: the fresh decoder
    // of the probe starts on a continuation byte and writes nothing).
    bool straddle_completes = false;
    if (pend0 > 0) {
#pragma unroll
        for (uint32_t p = 0; p < 3; ++p) {
            if (!m5_bit(acc, 32 + p) || (int32_t)p >= wlen) break;
            if (!m5_bit(pendm, 32 + p)) { straddle_completes = true; break; }
        }
    }
    const bool seg1_before = k_in > 0 || (!SBYTE && at_slice_start && pend0 > 0 && straddle_completes);
    const bool probe_seg1 = false;
    // Second segment at the slice start (the pending sequence broke at byte 0, which is read again) with a leftover in
    // front: the probe fires if that segment begins with a complete multi-byte char, and compares the fresh decode
    // with what sits at the start of the output buffer (probe_utf8, the literal emulation).
    bool seg0_before = false;
    if (!SBYTE && at_slice_start && mode != MODE_STATE && k_in > 0 && m5_bit(pre, 32) && m5_bit(lead, 32)) {
        bool completes = false;
#pragma unroll
        for (uint32_t p = 1; p < 4; ++p) {
            if ((int32_t)p >= wlen || !m5_bit(acc, 32 + p)) break;
            if (!m5_bit(pendm, 32 + p)) { completes = true; break; }
        }
        if (completes) seg0_before = probe_utf8(P, tsrc.g, geo.slice_start, geo.slice_end, false, pend0, kin);
    }
    const bool probe_seg0 = false;
    const uint32_t lo_flags = (k_in > 0 && (kin.flags & CF_HOSTCARRY)) ? (uint32_t)RF_HOSTCARRY : 0u;

    // One run of passing chars [sB, eB] (bit indices of its first / last byte).  is_left: it continues whatever touches
    // the window's left boundary (leftover of k_in chars, or a cut finding).  touches_end: its last char is the last
    // complete char of the window.  Restates helper.rs:237-431 for the run: forced cut every q chars (the piece after
    // a cut always "completes"), min-length rule otherwise; at the window end the rest is kept as leftover unless it
    // completes a cut finding (finding_collection.rs:255-290).  Returns false to decline.
    res.out = carry_none();
    res.cut1 = 0;  // here: 1 when the carry-out depends on the carry-in (see below)
    res.caseb = 0; res.a = 0; res.t_out = 0;
    bool carry_done = false;
    bool has_left = false;       // a run continues what touches the left boundary and ends inside the window at left_end
    uint32_t left_end = 0;
    bool seg1_cleared = false;   // ... and a later event of the first segment resets the "maybe cut" flag it may leave behind
    // A run that ends exactly at a forced cut inside the window leaves the "maybe cut" flag set (helper.rs:353,
    // finding_collection.rs:268).  It is reset by the next finding or leftover of the same segment; at the next
    // segment start it becomes that segment's "last was cut" (finding_collection.rs:240-241): a run beginning right
    // there completes the cut finding whatever its length; with no segment start left it is the window's carry-out.
    bool cutp = false, cutp_used = false;
    int32_t cutp_seg = 0;
    uint32_t cutp_pos = 0;
    auto do_run = [&](int32_t sB, uint32_t eB, bool is_left, bool touches_end, bool lastcut_in = false) -> bool {
        const uint32_t fromB = sB < 32 ? 32u : (uint32_t)sB;  // chars are counted at their last byte
        const uint32_t k0 = is_left ? k_in : 0u;
        const uint32_t inwin = m5_count(pe, fromB, eB);
        const uint32_t total = inwin + k0;
        const bool lastcut0 = (is_left && kc) || lastcut_in;
        int32_t seg_id = -1;
        if (!is_left) {
            const int32_t sg = m5_high_le(seg, (uint32_t)sB);  // segment start at or before the run
            seg_id = sg < 32 ? -1 : sg - 32;
        }
        const int32_t start_rel = (k0 > 0) ? -(int32_t)kin.in_bytes : sB - 32;
        // text of the run: the leftover's (decoded earlier) + the chars from the straddling char / the run start on
        const uint32_t text0 = k0 > 0 ? kin.out_bytes : 0u;
        const uint32_t tfrom = k0 > 0 ? 32u - (uint32_t)pend0 : (uint32_t)sB;
        const uint32_t fl0 = (is_left ? lo_flags : 0u) | (lastcut0 ? (uint32_t)RF_COMPLETES : 0u);
        const bool yields = total >= q || (touches_end ? lastcut0 : (lastcut0 || total >= n));
        if (yields && ((seg_id == 0 && probe_seg0) || (seg_id < 0 && probe_seg1))) return false;
        uint32_t prec = (seg_id == last_seg) ? PREC_AFTER
                        : ((seg_id < 0 && seg1_before) || (seg_id == 0 && seg0_before)) ? PREC_BEFORE : PREC_EXACT;
        const int32_t seg_rel = seg_id < 0 ? 0 : seg_id;
        if (yields) last_seg = seg_id;
        if (cutp && seg_id == cutp_seg && (yields || touches_end)) cutp = false;  // reset by this finding / leftover
        // a finding, or the leftover at the window end, of the first segment after the left run (cut := false / kept)
        if (!is_left && seg_id < 0 && (touches_end || total >= n)) seg1_cleared = true;
        if (total < q) {
            if (touches_end) {
                carry_done = true;
                if (lastcut0) {  // completes the cut finding and may be cut again at the window end
                    mask_emit(E, seg_rel, prec, start_rel, (int32_t)eB + 1 - 32, text0 + T(tfrom, eB), fl0);
                    res.out = carry_cut();
                } else {         // kept as leftover
                    Carry c;
                    c.kind = K_L; c.flags = (uint8_t)((is_left && k0 > 0 && (kin.flags & CF_HOSTCARRY)) ? CF_HOSTCARRY : 0);
                    c.k = (uint16_t)total;
                    c.in_bytes = (uint32_t)(wlen - start_rel);
                    c.out_bytes = text0 + T(tfrom, Bend - 1);
                    c.aux = 0;
                    res.out = c;
                }
            } else if (yields) mask_emit(E, seg_rel, prec, start_rel, (int32_t)eB + 1 - 32, text0 + T(tfrom, eB), fl0);
            return true;
        }
        // forced cuts every q chars
        uint32_t need = q - k0, remaining = inwin, posB = fromB, flp = fl0;
        int32_t piece_s = start_rel;
        uint32_t piece_t0 = text0, piece_from = tfrom;
        int pieces = 0;
        while (remaining >= need) {
            const uint32_t endB = m5_select(pe, posB, need);
            mask_emit(E, seg_rel, prec, piece_s, (int32_t)endB + 1 - 32, piece_t0 + T(piece_from, endB), flp);
            piece_t0 = 0;
            piece_from = endB + 1;
            prec = PREC_AFTER;
            flp = RF_COMPLETES;
            remaining -= need;
            need = q;
            posB = endB + 1;
            piece_s = (int32_t)endB + 1 - 32;
            if (++pieces > 24) return false;
        }
        if (remaining == 0) {
            // the run ends exactly at a cut: the "maybe cut" flag outlives it (until the next segment / the window end)
            if (touches_end) { carry_done = true; res.out = carry_cut(); }
            else {
                if (cutp_used) return false;  // twice in one window: byte-wise engine
                cutp = true; cutp_used = true; cutp_seg = seg_id; cutp_pos = eB;
            }
        } else {
            mask_emit(E, seg_rel, prec, piece_s, (int32_t)eB + 1 - 32, piece_t0 + T(piece_from, eB), RF_COMPLETES);
            if (touches_end) { carry_done = true; res.out = carry_cut(); }
        }
        return true;
    };

    // hands a pending "maybe cut" flag to the segment that starts before bit index `upto`; *ran_to: last byte of the run
    // it processed (a run beginning at that segment start), 0 if none
    auto flush_cut = [&](uint32_t upto, uint32_t* ran_to) -> bool {
        *ran_to = 0;
        if (!cutp) return true;
        const uint32_t s1 = m5_low_ge(seg, cutp_pos + 1);
        if (s1 >= 32u + (uint32_t)wlen || s1 > upto) return true;
        cutp = false;
        if (m5_bit(R, s1)) {
            const uint32_t e2 = m5_low_ge(RE, s1);
            if (!do_run((int32_t)s1, e2, false, e2 + 1 >= Bend, true)) return false;
            *ran_to = e2;
        }
        return true;
    };

    // ---- the run touching the left boundary: continues the leftover / completes the cut finding unless a pending
    //      sequence breaks first (then the leftover ends at that break, helper.rs:315-322) -----------------------------
    if (m5_bit(R, 32) && !m5_bit(pre, 32)) {
        const uint32_t e_last = m5_low_ge(RE, 32);
        const int32_t s0 = m5_high_le(RS, 32);
        if (s0 < 29 || e_last >= 32 * (NW + 1)) return false;
        if (!do_run(s0, e_last, true, e_last + 1 >= Bend)) return false;
        next_B = e_last + 1;
        const uint32_t inwin = m5_count(pe, 32, e_last);
        if (e_last + 1 >= Bend) {
            // a run covering the window: >= q chars by itself always end in a cut at the window end; fewer follow the
            // closed-form transfer function eval_caseb (leftover accumulates, or is cut once it reaches q)
            res.cut1 = inwin >= q ? 0u : 1u;
            if (inwin < q) {
                res.caseb = 1;
                res.a = (uint16_t)inwin;
                res.t_out = (uint16_t)T(32u - (uint32_t)pend0, Bend - 1);
            }
        } else {
            has_left = true;
            left_end = e_last;
        }
    } else if (k_in >= n) {
        // the leftover alone is long enough: printed at the first event of the window
        mask_emit(E, 0, PREC_BEFORE, -(int32_t)kin.in_bytes, -pend0, kin.out_bytes, lo_flags);
        last_seg = -1;
    }
    // ---- runs of >= n bytes inside the window ------------------------------------------------------------------------
    if (!carry_done) {
        // LR: last bytes of runs holding at least n bytes (n <= 128): AND of R shifted by 0 .. n-1
        M5 LR = R;
        uint32_t have = 1;
        auto and_shl = [&](uint32_t s) {
            if (s >= 32) {
                const uint32_t ws_ = s >> 5, bs = s & 31;
#pragma unroll
                for (int i = NW; i >= 0; --i) {
                    uint32_t lo = 0, hi = 0;
#pragma unroll
                    for (int j = 0; j < NW + 1; ++j) {
                        if ((uint32_t)j + ws_ == (uint32_t)i) hi = LR.w[j];
                        if ((uint32_t)j + ws_ + 1 == (uint32_t)i) lo = LR.w[j];
                    }
                    LR.w[i] &= bs ? sx_fsl(lo, hi, bs) : hi;
                }
            } else {
#pragma unroll
                for (int i = NW; i >= 1; --i) LR.w[i] &= sx_fsl(LR.w[i - 1], LR.w[i], s);
                LR.w[0] &= LR.w[0] << s;
            }
        };
        if (n >= 2) { and_shl(1); have = 2; }
        if (n >= 4) { and_shl(2); have = 4; }
        if (n >= 8) { and_shl(4); have = 8; }
        if (n >= 16) { and_shl(8); have = 16; }
        if (n >= 32) { and_shl(16); have = 32; }
        if (n >= 64) { and_shl(32); have = 64; }
        if (n > have) and_shl(n - have);
        M5 cand;  // last bytes of the long runs
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) cand.w[i] = LR.w[i] & RE.w[i];
        int guard = 0;
        for (; guard < 12; ++guard) {
            const uint32_t e_last = m5_low_ge(cand, next_B);
            if (e_last >= 32 * (NW + 1)) break;
            next_B = e_last + 1;
            const int32_t s = m5_high_le(RS, e_last);
            if (s < 32) return false;
            uint32_t ran_to;
            if (!flush_cut((uint32_t)s, &ran_to)) return false;
            if (ran_to != e_last && !do_run(s, e_last, false, e_last + 1 >= Bend)) return false;
            if (carry_done) break;
        }
        if (guard == 12) return false;  // a window crowded with short findings: byte-wise engine
    }
    if (!carry_done) {
        uint32_t ran_to;
        if (!flush_cut(0xFFFFFFFFu, &ran_to)) return false;
    }
    // ---- carry out: a short run of complete chars touching the window end (finding_collection.rs:281-284) ---------
    if (!carry_done && Bend > 32 && m5_bit(R, Bend - 1)) {
        const int32_t ts = m5_high_le(RS, Bend - 1);
        if (ts < 32) return false;
        if (!do_run(ts, Bend - 1, false, true)) return false;
    }
    // a malformed last byte starts an (empty) segment at the window end, which swallows the flag (finding_collection.rs:134-143)
    const bool seg_at_end = m5_bit(mal, 31u + (uint32_t)wlen);
    if (cutp && !carry_done && !seg_at_end) res.out = carry_cut();  // no segment start behind the run: the flag is the carry (Q4)
    if (has_left) {
        // Does the carry-out depend on the carry-in?  The carry-in decides whether the left run is printed (no trace in
        // the carry) and, when leftover + run is an exact multiple of q, leaves the "maybe cut" flag set behind the run
        // (helper.rs:353, finding_collection.rs:268).  The flag dies at the next finding or leftover of the same segment;
        // at the next segment start it becomes that segment's "last was cut" (finding_collection.rs:240), which only
        // matters if the segment begins with a run of passing chars -- and only shows in the carry if that run (< q
        // chars) reaches the window end (cut again vs kept as leftover).
        bool dep = false;
        if (!seg1_cleared) {
            const uint32_t s1 = m5_low_ge(seg, left_end + 1);
            if (s1 >= 32u + (uint32_t)wlen) dep = !seg_at_end;  // the first segment runs to the window end: the flag is the carry
            else if (m5_bit(R, s1)) {
                const uint32_t e2 = m5_low_ge(RE, s1);
                dep = e2 + 1 >= Bend && m5_count(pe, s1, e2) < q;
            }
        }
        res.cut1 = dep ? 1u : 0u;
    }
    res.nrec = E.nrec;
    res.ntext = E.ntext;
    res.npend_out = npend_out;
    res.m = 1;
    return true;
}

// Short windows (pre-roll, extension) only need two data words.
template <bool SBYTE, class TileSrc>
SX_HD bool mask_window(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, const Carry& kin, int mode, Record* wr,
                       uint64_t text_off, WinResult& res) {
    if (geo.we - geo.ws <= 64) return mask_window_nw<2, false, SBYTE>(P, tsrc, geo, kin, mode, wr, text_off, res);
    return mask_window_nw<4, false, SBYTE>(P, tsrc, geo, kin, mode, wr, text_off, res);
}
// A head (predecessor window not listed) in ONE pass: the pre-roll region is the 32 bytes in front of the window.
constexpr uint32_t kMaskLb32MaxPre = 28;
template <bool SBYTE, class TileSrc>
SX_HD bool mask_head(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, uint32_t pre_bytes, int mode, Record* wr,
                     uint64_t text_off, WinResult& res) {
    if (pre_bytes > kMaskLb32MaxPre || geo.ws < 32 || geo.we - geo.ws <= 64) return false;
    return mask_window_nw<4, true, SBYTE>(P, tsrc, geo, carry_none(), mode, wr, text_off, res);
}
template <class TileSrc>
SX_HD bool utf8_mask_window(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, const Carry& kin, int mode, Record* wr,
                            uint64_t text_off, WinResult& res) {
    return mask_window<false>(P, tsrc, geo, kin, mode, wr, text_off, res);
}
template <class TileSrc>
SX_HD bool utf8_mask_head(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, uint32_t pre_bytes, int mode, Record* wr,
                          uint64_t text_off, WinResult& res) {
    return mask_head<false>(P, tsrc, geo, pre_bytes, mode, wr, text_off, res);
}
// which decoders the mask engine covers
template <class Dec> struct MaskFamily { static constexpr bool kHas = false, kSByte = false; };
template <> struct MaskFamily<DecUtf8> { static constexpr bool kHas = true, kSByte = false; };
template <> struct MaskFamily<DecXud> { static constexpr bool kHas = true, kSByte = true; };
template <> struct MaskFamily<DecSb> { static constexpr bool kHas = true, kSByte = true; };

// Class table of the single-byte family (Utf8Tables.cls): bit 0 filter verdict of the mapped char, bit 1 unmapped byte
// (malformed, DecSb), bits 2-3 UTF-8 length - 1 of the mapped char.
SX_HD void sbyte_cls_fill(const ScanParams& P, Utf8Tables& T, uint32_t b) {
    uint32_t pass, mal = 0, len = 1;
    if (b < 0x80) pass = pass_filter(P, b) ? 1u : 0u;
    else if (P.enc == ENC_XUD) { pass = pass_filter(P, 0xEF) ? 1u : 0u; len = 3; }  // U+F780 + (b - 0x80)
    else {
        const uint32_t cp = P.sb_table[b - 0x80];
        if (cp == 0) { mal = 1; pass = 0; }
        else { pass = pass_filter(P, utf8_lead_of_cp(cp)) ? 1u : 0u; len = utf8_len_of_cp(cp); }
    }
    T.cls[b] = (uint8_t)(pass | (mal << 1) | ((len - 1u) << 2));
}
// i in 0..2047: the tables the engines of the mission's encoding read
SX_HD void mask_tables_fill(const ScanParams& P, Utf8Tables& T, uint32_t i) {
    if (P.enc == ENC_UTF8) utf8_tables_fill(P, T, i);
    else if (i < 256) sbyte_cls_fill(P, T, i);
}

#if !defined(__CUDA_ARCH__)
// Host harness only: try the mask engine and cross-check it against the byte-wise engine `fallback` (results and records).
template <bool SBYTE, class TileSrc, class Fallback>
inline bool host_mask_try(const ScanParams& P, const TileSrc& tsrc, const WinGeom& geo, const Carry& kin, int mode, Record* wr,
                          uint64_t text_off, WinResult& res, Fallback&& fallback) {
    Record tmp[80];
    const bool wr_mode = mode == MODE_WRITE || mode == MODE_BUFFER;
    const bool ok = mask_window<SBYTE>(P, tsrc, geo, kin, mode, wr, text_off, res);
    tsrc.mask_result(ok);
    if (!ok) return false;
    WinResult r2;
    fallback(tmp, r2);
    bool same = r2.nrec == res.nrec && r2.ntext == res.ntext && r2.npend_out == res.npend_out && r2.out.kind == res.out.kind &&
                r2.out.k == res.out.k && r2.out.in_bytes == res.out.in_bytes && r2.out.out_bytes == res.out.out_bytes &&
                r2.out.flags == res.out.flags;
    if (same && wr_mode)
        for (uint32_t k = 0; k < res.nrec && (mode == MODE_WRITE || k < kBufRecs); ++k)
            same = same && tmp[k].position == wr[k].position && tmp[k].in_start == wr[k].in_start && tmp[k].in_len == wr[k].in_len &&
                   tmp[k].flags == wr[k].flags && tmp[k].precision == wr[k].precision && tmp[k].text_off == wr[k].text_off &&
                   tmp[k].text_len == wr[k].text_len;
    if (!same) tsrc.mask_mismatch(geo.ws, geo.we, kin, mode);
    return true;
}
#endif

// Engine dispatch used by the kernels and the test harness: the mask engine first where it exists (host harness; the
// kernels call it themselves and queue what it declines), then the byte-wise engines; grep_char / same-unicode-block /
// chars_min_nb > q missions take the general automaton.
template <class Dec> struct WindowEngine {
    template <class TileSrc>
    SX_HD static void run(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo, const Carry& kin,
                          int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
#if !defined(__CUDA_ARCH__)
        if (MaskFamily<Dec>::kSByte && tsrc.tables() && !P.general && !desc && tsrc.use_mask() &&
            host_mask_try<true>(P, tsrc, geo, kin, mode, wr, text_off, res, [&](Record* tmp, WinResult& r2) {
                scan_window<Dec>(P, tsrc, g, geo, kin, mode, tmp, text_off, r2, nullptr);
            }))
            return;
#endif
        if (Dec::kStateful && !P.general) scan_window_fast_generic<Dec>(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
        else scan_window<Dec>(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
    }
};
template <> struct WindowEngine<DecUtf8> {
    template <class TileSrc>
    SX_HD static void run(const ScanParams& P, const TileSrc& tsrc, const GlobalSrc& g, const WinGeom& geo, const Carry& kin,
                          int mode, Record* wr, uint64_t text_off, WinResult& res, WinDesc* desc) {
        if (tsrc.tables() && !P.general) {
#if !defined(__CUDA_ARCH__)
            if (!desc && tsrc.use_mask() &&
                host_mask_try<false>(P, tsrc, geo, kin, mode, wr, text_off, res, [&](Record* tmp, WinResult& r2) {
                    scan_window_fast_utf8(P, tsrc, g, geo, kin, mode, tmp, text_off, r2, nullptr);
                }))
                return;
#endif
            scan_window_fast_utf8(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
        } else scan_window<DecUtf8>(P, tsrc, g, geo, kin, mode, wr, text_off, res, desc);
    }
};

}  // namespace sx
