/*
 * sx_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
 * See sx_oracle.h for the scope statement and the reference file:line map.
 *
 * The control flow deliberately mirrors the reference (buffered segment text + SplitStr
 * iterator), while the CUDA product uses a streaming formulation -- so a differential test
 * between the two is meaningful.
 */
#include "sx_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* finding.rs:23 */
static size_t g_output_buf_len = 0x9192;
void sxo_set_output_buf_len(size_t n) { g_output_buf_len = n; }

enum { R_INPUT_EMPTY = 0, R_OUTPUT_FULL = 1, R_MALFORMED = 2 };

/* ------------------------------------------------------------------------------------------
 * Decoders: restatement of encoding_rs 0.8.34 `decode_to_utf8_raw` for each encoding used,
 * including its destination-space checks (check_space_bmp: pos+2 < len, check_space_astral:
 * pos+3 < len) because the reference's 8-byte Precision::Before probe
 * (finding_collection.rs:180-194) observes them through `written`.
 * ---------------------------------------------------------------------------------------- */
struct sxo_decoder {
    uint32_t enc;
    uint16_t table[128];
    /* UTF-8 (WHATWG utf-8 decoder state) */
    uint32_t cp;
    uint8_t need, seen, lo, hi;
    /* UTF-16 */
    uint16_t lead_surrogate;
    int lead_byte; /* -1 = None */
    int pending_bmp;
    /* UTF-32 (extension) */
    uint8_t u32buf[4];
    int u32n;
    /* Big5 / EUC-JP (WHATWG decoders): pending lead byte, EUC-JP jis0212 flag */
    uint8_t mb_lead;
    uint8_t mb_j0212;
};

static void dec_reset(sxo_decoder *d) {
    d->cp = 0;
    d->need = 0;
    d->seen = 0;
    d->lo = 0x80;
    d->hi = 0xBF;
    d->lead_surrogate = 0;
    d->lead_byte = -1;
    d->pending_bmp = 0;
    d->u32n = 0;
    d->mb_lead = 0;
    d->mb_j0212 = 0;
}

static void dec_init(sxo_decoder *d, uint32_t enc, const uint16_t *tab) {
    memset(d, 0, sizeof *d);
    d->enc = enc;
    if (tab) memcpy(d->table, tab, sizeof d->table);
    dec_reset(d);
}

static size_t put_utf8(uint8_t *dst, uint32_t c) {
    if (c < 0x80) {
        dst[0] = (uint8_t)c;
        return 1;
    }
    if (c < 0x800) {
        dst[0] = (uint8_t)(0xC0 | (c >> 6));
        dst[1] = (uint8_t)(0x80 | (c & 0x3F));
        return 2;
    }
    if (c < 0x10000) {
        dst[0] = (uint8_t)(0xE0 | (c >> 12));
        dst[1] = (uint8_t)(0x80 | ((c >> 6) & 0x3F));
        dst[2] = (uint8_t)(0x80 | (c & 0x3F));
        return 3;
    }
    dst[0] = (uint8_t)(0xF0 | (c >> 18));
    dst[1] = (uint8_t)(0x80 | ((c >> 12) & 0x3F));
    dst[2] = (uint8_t)(0x80 | ((c >> 6) & 0x3F));
    dst[3] = (uint8_t)(0x80 | (c & 0x3F));
    return 4;
}

/* Longest prefix of complete, valid UTF-8 sequences (encoding_rs utf8_valid_up_to). */
static size_t utf8_valid_up_to(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t b = s[i];
        if (b < 0x80) {
            i++;
            continue;
        }
        if (b < 0xC2) return i;
        if (b < 0xE0) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return i;
            i += 2;
            continue;
        }
        if (b < 0xF0) {
            uint8_t lo = (b == 0xE0) ? 0xA0 : 0x80, hi = (b == 0xED) ? 0x9F : 0xBF;
            if (i + 2 >= n || s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80) return i;
            i += 3;
            continue;
        }
        if (b < 0xF5) {
            uint8_t lo = (b == 0xF0) ? 0x90 : 0x80, hi = (b == 0xF4) ? 0x8F : 0xBF;
            if (i + 3 >= n || s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80 ||
                (s[i + 3] & 0xC0) != 0x80)
                return i;
            i += 4;
            continue;
        }
        return i;
    }
    return i;
}

/* encoding_rs utf_8.rs Utf8Decoder::decode_to_utf8_raw */
static int dec_utf8(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last, size_t *rd,
                    size_t *wr) {
    size_t sp = 0, dp = 0;
    for (;;) {
        if (d->need == 0) { /* loop preamble: Utf8Destination::copy_utf8_up_to_invalid_from */
            size_t n = slen - sp < dlen - dp ? slen - sp : dlen - dp;
            size_t v = utf8_valid_up_to(src + sp, n);
            memcpy(dst + dp, src + sp, v);
            sp += v;
            dp += v;
        }
        if (sp >= slen) {
            if (last && d->need != 0) {
                dec_reset(d);
                *rd = sp;
                *wr = dp;
                return R_MALFORMED;
            }
            *rd = sp;
            *wr = dp;
            return R_INPUT_EMPTY;
        }
        if (!(dp + 3 < dlen)) { /* check_space_astral */
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        uint8_t b = src[sp++];
        if (d->need == 0) {
            if (b < 0x80) {
                dst[dp++] = b;
                continue;
            }
            if (b < 0xC2) goto malformed;
            if (b < 0xE0) {
                d->need = 1;
                d->cp = b & 0x1F;
                continue;
            }
            if (b < 0xF0) {
                if (b == 0xED)
                    d->hi = 0x9F;
                else if (b == 0xE0)
                    d->lo = 0xA0;
                d->need = 2;
                d->cp = b & 0xF;
                continue;
            }
            if (b < 0xF5) {
                if (b == 0xF4)
                    d->hi = 0x8F;
                else if (b == 0xF0)
                    d->lo = 0x90;
                d->need = 3;
                d->cp = b & 0x7;
                continue;
            }
            goto malformed;
        }
        if (!(b >= d->lo && b <= d->hi)) {
            dec_reset(d);
            sp--; /* unread_handle.unread(): the offending byte is not consumed */
            goto malformed;
        }
        d->lo = 0x80;
        d->hi = 0xBF;
        d->cp = (d->cp << 6) | (b & 0x3F);
        d->seen++;
        if (d->seen != d->need) continue;
        dp += put_utf8(dst + dp, d->cp);
        dec_reset(d);
        continue;
    malformed:
        *rd = sp;
        *wr = dp;
        return R_MALFORMED;
    }
}

/* encoding_rs utf_16.rs Utf16Decoder (decoder_function! with check_space_astral) */
static int dec_utf16(sxo_decoder *d, int be, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last,
                     size_t *rd, size_t *wr) {
    size_t sp = 0, dp = 0;
    if (d->pending_bmp) { /* preamble */
        if (!(dp + 2 < dlen)) {
            *rd = 0;
            *wr = 0;
            return R_OUTPUT_FULL;
        }
        dp += put_utf8(dst + dp, d->lead_surrogate);
        d->pending_bmp = 0;
        d->lead_surrogate = 0;
    }
    for (;;) {
        if (sp >= slen) {
            if (last) {
                if (d->lead_surrogate != 0 || d->lead_byte >= 0) {
                    d->lead_surrogate = 0;
                    d->lead_byte = -1;
                    *rd = sp;
                    *wr = dp;
                    return R_MALFORMED;
                }
            }
            *rd = sp;
            *wr = dp;
            return R_INPUT_EMPTY;
        }
        if (!(dp + 3 < dlen)) {
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        uint8_t b = src[sp++];
        if (d->lead_byte < 0) {
            d->lead_byte = b;
            continue;
        }
        uint16_t lead = (uint16_t)d->lead_byte;
        d->lead_byte = -1;
        uint16_t cu = be ? (uint16_t)((lead << 8) | b) : (uint16_t)(((uint16_t)b << 8) | lead);
        uint16_t hb = cu & 0xFC00;
        if (hb == 0xD800) {
            if (d->lead_surrogate != 0) {
                /* previous high surrogate in error, this one becomes the pending one */
                d->lead_surrogate = cu;
                *rd = sp;
                *wr = dp;
                return R_MALFORMED;
            }
            d->lead_surrogate = cu;
            continue;
        }
        if (hb == 0xDC00) {
            if (d->lead_surrogate == 0) {
                *rd = sp;
                *wr = dp;
                return R_MALFORMED;
            }
            uint32_t c = 0x10000u + (((uint32_t)d->lead_surrogate - 0xD800u) << 10) + ((uint32_t)cu - 0xDC00u);
            dp += put_utf8(dst + dp, c);
            d->lead_surrogate = 0;
            continue;
        }
        if (d->lead_surrogate != 0) {
            /* previous high surrogate in error; this unit becomes a pending BMP character
             * that is written by the preamble of the NEXT call */
            d->lead_surrogate = cu;
            d->pending_bmp = 1;
            *rd = sp;
            *wr = dp;
            return R_MALFORMED;
        }
        dp += put_utf8(dst + dp, cu);
    }
}

/* encoding_rs x_user_defined.rs (decoder_function! with check_space_bmp) */
static int dec_xud(const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, size_t *rd, size_t *wr) {
    size_t sp = 0, dp = 0;
    for (;;) {
        if (sp >= slen) {
            *rd = sp;
            *wr = dp;
            return R_INPUT_EMPTY;
        }
        if (!(dp + 2 < dlen)) {
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        uint8_t b = src[sp++];
        if (b < 0x80)
            dst[dp++] = b;
        else
            dp += put_utf8(dst + dp, (uint32_t)b + 0xF700u);
    }
}

/* encoding_rs single_byte.rs SingleByteDecoder::decode_to_utf8_raw */
static int dec_single(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, size_t *rd,
                      size_t *wr) {
    size_t sp = 0, dp = 0;
    uint8_t non_ascii, b;
outermost: {
    size_t srem = slen - sp, drem = dlen - dp;
    int pending = drem < srem ? R_OUTPUT_FULL : R_INPUT_EMPTY;
    size_t length = drem < srem ? drem : srem, i = 0;
    while (i < length && src[sp + i] < 0x80) {
        dst[dp + i] = src[sp + i];
        i++;
    }
    sp += i;
    dp += i;
    if (i == length) {
        *rd = sp;
        *wr = dp;
        return pending;
    }
    if (dp + 2 < dlen) {
        non_ascii = src[sp++];
    } else {
        *rd = sp;
        *wr = dp;
        return R_OUTPUT_FULL;
    }
}
middle: {
    uint16_t mapped = d->table[non_ascii - 0x80];
    if (mapped == 0) {
        *rd = sp;
        *wr = dp;
        return R_MALFORMED;
    }
    dp += put_utf8(dst + dp, mapped);
    if (sp >= slen) {
        *rd = sp;
        *wr = dp;
        return R_INPUT_EMPTY;
    }
    if (!(dp + 2 < dlen)) {
        *rd = sp;
        *wr = dp;
        return R_OUTPUT_FULL;
    }
    b = src[sp++];
}
    for (;;) { /* innermost */
        if (b > 127) {
            non_ascii = b;
            goto middle;
        }
        dst[dp++] = b;
        if (b < 60) {
            if (sp >= slen) {
                *rd = sp;
                *wr = dp;
                return R_INPUT_EMPTY;
            }
            if (!(dp + 2 < dlen)) {
                *rd = sp;
                *wr = dp;
                return R_OUTPUT_FULL;
            }
            b = src[sp++];
            continue;
        }
        goto outermost;
    }
}

/* EXTENSION (no reference semantics): 4-byte units aligned to the stream,
 * valid iff <= 0x10FFFF and not a surrogate, else Malformed read past the unit. */
static int dec_utf32(sxo_decoder *d, int be, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last,
                     size_t *rd, size_t *wr) {
    size_t sp = 0, dp = 0;
    for (;;) {
        if (sp >= slen) {
            if (last && d->u32n != 0) {
                d->u32n = 0;
                *rd = sp;
                *wr = dp;
                return R_MALFORMED;
            }
            *rd = sp;
            *wr = dp;
            return R_INPUT_EMPTY;
        }
        if (!(dp + 3 < dlen)) {
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        d->u32buf[d->u32n++] = src[sp++];
        if (d->u32n < 4) continue;
        d->u32n = 0;
        uint32_t c = be ? ((uint32_t)d->u32buf[0] << 24 | (uint32_t)d->u32buf[1] << 16 | (uint32_t)d->u32buf[2] << 8 |
                           d->u32buf[3])
                        : ((uint32_t)d->u32buf[3] << 24 | (uint32_t)d->u32buf[2] << 16 | (uint32_t)d->u32buf[1] << 8 |
                           d->u32buf[0]);
        if (c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) {
            *rd = sp;
            *wr = dp;
            return R_MALFORMED;
        }
        dp += put_utf8(dst + dp, c);
    }
}

/* ------------------------------------------------------------------------------------------
 * Big5 and EUC-JP: the WHATWG Encoding Standard decoders (https://encoding.spec.whatwg.org/#big5-decoder,
 * #euc-jp-decoder) in the calling convention of encoding_rs 0.8.34 (big5.rs / euc_jp.rs): `read` ends after the bytes
 * the decoder consumed for a malformed sequence -- an ASCII byte that fails as a trail is NOT consumed (it is
 * "prepended to the stream", i.e. re-read by the next call) --, an incomplete sequence at the end of the input stays
 * pending, destination space is checked before every byte read (check_space_astral: 4 free bytes for Big5, whose
 * pointers 1133 / 1135 / 1164 / 1166 decode to TWO code points; check_space_bmp: 3 for EUC-JP).
 * The index tables are the library's (stringsext_b200/csrc/sx_mb_tables.inc, generated from CPython's codecs because
 * the WHATWG index files are not available offline): PARITY UNPINNED w.r.t. encoding_rs' tables and `written`
 * details -- the tests check "oracle == kernels on the same table" (SURVEY.md 8(c)).
 * ---------------------------------------------------------------------------------------- */
#include "../stringsext_b200/csrc/sx_mb_tables.inc"

static int dec_big5(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last, size_t *rd,
                    size_t *wr) {
    size_t sp = 0, dp = 0;
    for (;;) {
        if (sp >= slen) {
            *rd = sp;
            *wr = dp;
            if (last && d->mb_lead) {
                d->mb_lead = 0;
                return R_MALFORMED;
            }
            return R_INPUT_EMPTY;
        }
        if (!(dp + 3 < dlen)) {
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        uint8_t b = src[sp++];
        if (d->mb_lead) {
            uint32_t l = d->mb_lead;
            d->mb_lead = 0;
            if ((b >= 0x40 && b <= 0x7E) || (b >= 0xA1 && b <= 0xFE)) {
                uint32_t ptr = (l - 0x81u) * 157u + (uint32_t)(b - (b < 0x7F ? 0x40 : 0x62));
                if (ptr == 1133 || ptr == 1135 || ptr == 1164 || ptr == 1166) {
                    dp += put_utf8(dst + dp, ptr < 1150 ? 0xCA : 0xEA);
                    dp += put_utf8(dst + dp, (ptr == 1133 || ptr == 1164) ? 0x304 : 0x30C);
                    continue;
                }
                uint32_t cp = kSxBig5Index[ptr];
                if (cp) {
                    dp += put_utf8(dst + dp, cp);
                    continue;
                }
            }
            if (b < 0x80) sp--; /* prepend the ASCII byte to the stream */
            *rd = sp;
            *wr = dp;
            return R_MALFORMED;
        }
        if (b < 0x80) {
            dst[dp++] = b;
            continue;
        }
        if (b >= 0x81 && b <= 0xFE) {
            d->mb_lead = b;
            continue;
        }
        *rd = sp;
        *wr = dp;
        return R_MALFORMED;
    }
}

static int dec_eucjp(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last, size_t *rd,
                     size_t *wr) {
    size_t sp = 0, dp = 0;
    for (;;) {
        if (sp >= slen) {
            *rd = sp;
            *wr = dp;
            if (last && d->mb_lead) {
                d->mb_lead = 0;
                d->mb_j0212 = 0;
                return R_MALFORMED;
            }
            return R_INPUT_EMPTY;
        }
        if (!(dp + 2 < dlen)) {
            *rd = sp;
            *wr = dp;
            return R_OUTPUT_FULL;
        }
        uint8_t b = src[sp++];
        if (d->mb_lead == 0x8E && b >= 0xA1 && b <= 0xDF) {
            d->mb_lead = 0;
            dp += put_utf8(dst + dp, 0xFF61u - 0xA1u + b);
            continue;
        }
        if (d->mb_lead == 0x8F && b >= 0xA1 && b <= 0xFE) {
            d->mb_j0212 = 1;
            d->mb_lead = b;
            continue;
        }
        if (d->mb_lead) {
            uint32_t l = d->mb_lead, cp = 0;
            int three = d->mb_j0212;
            d->mb_lead = 0;
            d->mb_j0212 = 0;
            if (l >= 0xA1 && l <= 0xFE && b >= 0xA1 && b <= 0xFE)
                cp = (three ? kSxJis0212Index : kSxJis0208Index)[(l - 0xA1u) * 94u + (uint32_t)(b - 0xA1)];
            if (cp) {
                dp += put_utf8(dst + dp, cp);
                continue;
            }
            if (b < 0x80) sp--;
            *rd = sp;
            *wr = dp;
            return R_MALFORMED;
        }
        if (b < 0x80) {
            dst[dp++] = b;
            continue;
        }
        if (b == 0x8E || b == 0x8F || (b >= 0xA1 && b <= 0xFE)) {
            d->mb_lead = b;
            continue;
        }
        *rd = sp;
        *wr = dp;
        return R_MALFORMED;
    }
}

static int dec_decode(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last, size_t *rd,
                      size_t *wr) {
    switch (d->enc) {
    case SXO_ENC_BIG5: return dec_big5(d, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_EUC_JP: return dec_eucjp(d, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_X_USER_DEFINED: return dec_xud(src, slen, dst, dlen, rd, wr);
    case SXO_ENC_UTF_8: return dec_utf8(d, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_UTF_16LE: return dec_utf16(d, 0, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_UTF_16BE: return dec_utf16(d, 1, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_SINGLE_BYTE: return dec_single(d, src, slen, dst, dlen, rd, wr);
    case SXO_ENC_UTF_32LE: return dec_utf32(d, 0, src, slen, dst, dlen, last, rd, wr);
    case SXO_ENC_UTF_32BE: return dec_utf32(d, 1, src, slen, dst, dlen, last, rd, wr);
    }
    *rd = slen;
    *wr = 0;
    return R_INPUT_EMPTY;
}

sxo_decoder *sxo_decoder_new(uint32_t enc, const uint16_t *tab) {
    sxo_decoder *d = (sxo_decoder *)malloc(sizeof *d);
    dec_init(d, enc, tab);
    return d;
}
void sxo_decoder_free(sxo_decoder *d) { free(d); }
int sxo_decoder_decode(sxo_decoder *d, const uint8_t *src, size_t slen, uint8_t *dst, size_t dlen, int last, size_t *rd,
                       size_t *wr) {
    return dec_decode(d, src, slen, dst, dlen, last, rd, wr);
}

/* ------------------------------------------------------------------------------------------
 * Utf8Filter (mission.rs:329-349)
 * ---------------------------------------------------------------------------------------- */
static int pass_af(uint64_t lo, uint64_t hi, uint8_t b) { /* b <= 0x7f */
    return b < 64 ? (int)((lo >> b) & 1) : (int)((hi >> (b - 64)) & 1);
}
static int pass_ubf(uint64_t ubf, uint8_t b) { /* b > 0x7f */ return (int)((ubf >> (b & 0x3f)) & 1); }

/* ------------------------------------------------------------------------------------------
 * SplitStr (helper.rs:58-433)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *inp;
    size_t end, start_p, p;
    uint8_t chars_min_nb;
    int same_block, last_cut, invalid_after;
    uint64_t af_lo, af_hi, ubf;
    int grep; /* -1 None */
    size_t s_char_nb_max;
} splitstr;

/* helper.rs:210-432; returns 1 for Some(result) */
static int splitstr_next(splitstr *it, sxo_split_result *r) {
    int grep_ok = it->grep < 0;
    size_t ok_p = it->p, ok_len = 0, ok_n = 0;
    uint8_t last_mb = 0;
    size_t ok_max = it->s_char_nb_max;

    while (it->p < it->end && ok_n < ok_max) { /* helper.rs:237 */
        uint8_t lb = it->inp[it->p];
        size_t cl;
        if ((lb & 0x80) == 0) {
            if (!grep_ok && it->grep == (int)lb) grep_ok = 1;
            cl = 1;
        } else if ((lb & 0xE0) == 0xC0)
            cl = 2;
        else if ((lb & 0xF0) == 0xE0)
            cl = 3;
        else if ((lb & 0xF8) == 0xF0)
            cl = 4;
        else
            cl = 1;
        int ok, advance;
        if (cl == 1) { /* helper.rs:275-276; NB a stray byte >=0x80 with cl==1 "should never occur" */
            ok = (lb & 0x80) ? 0 : pass_af(it->af_lo, it->af_hi, lb);
            advance = 1;
        } else if (pass_ubf(it->ubf, lb)) {
            if (!it->same_block || lb == last_mb || last_mb == 0) {
                last_mb = lb;
                ok = 1;
                advance = 1;
            } else {
                last_mb = lb;
                ok = 0;
                advance = 0;
            }
        } else {
            last_mb = 0;
            ok = 0;
            advance = 1;
        }
        if (ok) {
            ok_len += cl;
            ok_n += 1;
            it->p += cl;
        } else {
            if (advance) it->p += cl;
            if ((it->last_cut && ok_n > 0 && ok_p == it->start_p) || (ok_n >= it->chars_min_nb && grep_ok)) break;
            ok_len = 0;
            ok_n = 0;
            ok_p = it->p;
            grep_ok = it->grep < 0;
        }
    }
    if (ok_len == 0) return 0; /* helper.rs:343 */
    int left = ok_p == it->start_p;
    int right = ok_p + ok_len >= it->end;
    int maybe_cut = ok_n >= ok_max || (right && !it->invalid_after);
    int completes = left && it->last_cut;
    int again = !completes && right && !it->invalid_after && (ok_n < it->s_char_nb_max || !grep_ok);
    int min_ok = ok_n >= it->chars_min_nb;
    if (!completes && !again && (!grep_ok || !min_ok)) return 0; /* helper.rs:410-415 */
    if (ok_n >= ok_max) it->start_p = it->p;
    it->last_cut = maybe_cut;
    r->s_off = (uint32_t)ok_p;
    r->s_len = (uint32_t)ok_len;
    r->completes = (uint8_t)completes;
    r->maybe_cut = (uint8_t)maybe_cut;
    r->again = (uint8_t)again;
    r->min_ok = (uint8_t)min_ok;
    r->grep_ok = (uint8_t)grep_ok;
    return 1;
}

size_t sxo_split_str(const uint8_t *s, size_t len, uint8_t n, int same_block, int last_cut, int invalid_after,
                     uint64_t af_lo, uint64_t af_hi, uint64_t ubf, int grep_char, size_t qmax, sxo_split_result *out,
                     size_t max) {
    splitstr it = {s, len, 0, 0, n, same_block, last_cut, invalid_after, af_lo, af_hi, ubf, grep_char, qmax};
    size_t k = 0;
    while (k < max && splitstr_next(&it, &out[k])) k++;
    return k;
}

size_t sxo_char_count(const uint8_t *s, size_t len) {
    size_t n = 0, i = 0;
    while (i < len) {
        uint8_t c = s[i];
        i += (c & 0x80) == 0 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 1;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * ScannerState (scanner.rs:40-89) and FindingCollection (finding_collection.rs:31-63)
 * ---------------------------------------------------------------------------------------- */
struct sxo_state {
    sxo_mission m;
    sxo_decoder dec;
    uint8_t *leftover;
    size_t leftover_len, leftover_cap;
    int cut;
    uint64_t consumed;
};

struct sxo_fc {
    sxo_finding *v;
    size_t n, cap;
    uint8_t *text;
    size_t text_len, text_cap;
    uint64_t first_byte_position;
    int str_buf_overflow;
};

sxo_state *sxo_state_new(const sxo_mission *m, const uint16_t *tab) {
    sxo_state *ss = (sxo_state *)calloc(1, sizeof *ss);
    ss->m = *m;
    dec_init(&ss->dec, m->encoding_id, tab);
    ss->consumed = m->counter_offset; /* scanner.rs:86 */
    return ss;
}
void sxo_state_free(sxo_state *ss) {
    if (!ss) return;
    free(ss->leftover);
    free(ss);
}
uint64_t sxo_state_consumed(const sxo_state *ss) { return ss->consumed; }
int sxo_state_cut(const sxo_state *ss) { return ss->cut; }
size_t sxo_state_leftover(const sxo_state *ss, const uint8_t **p) {
    *p = ss->leftover;
    return ss->leftover_len;
}
size_t sxo_state_decoder_pending(const sxo_state *ss, uint8_t *o) {
    const sxo_decoder *d = &ss->dec;
    size_t k = 0;
    switch (d->enc) {
    case SXO_ENC_UTF_8:
        o[k++] = d->need;
        o[k++] = d->seen;
        o[k++] = d->lo;
        o[k++] = d->hi;
        o[k++] = (uint8_t)(d->cp & 0xFF);
        o[k++] = (uint8_t)((d->cp >> 8) & 0xFF);
        o[k++] = (uint8_t)((d->cp >> 16) & 0xFF);
        break;
    case SXO_ENC_UTF_16LE:
    case SXO_ENC_UTF_16BE:
        o[k++] = (uint8_t)(d->lead_byte >= 0);
        o[k++] = (uint8_t)(d->lead_byte >= 0 ? d->lead_byte : 0);
        o[k++] = (uint8_t)(d->lead_surrogate & 0xFF);
        o[k++] = (uint8_t)(d->lead_surrogate >> 8);
        o[k++] = (uint8_t)d->pending_bmp;
        break;
    case SXO_ENC_UTF_32LE:
    case SXO_ENC_UTF_32BE:
        o[k++] = (uint8_t)d->u32n;
        for (int i = 0; i < d->u32n; i++) o[k++] = d->u32buf[i];
        break;
    default: break;
    }
    return k;
}

static sxo_fc *fc_new(uint64_t off) {
    sxo_fc *fc = (sxo_fc *)calloc(1, sizeof *fc);
    fc->first_byte_position = off;
    return fc;
}
static void fc_push(sxo_fc *fc, const sxo_state *ss, int file_id, uint64_t pos, int prec, const uint8_t *s, size_t len,
                    int completes) {
    if (fc->n == fc->cap) {
        fc->cap = fc->cap ? fc->cap * 2 : 16;
        fc->v = (sxo_finding *)realloc(fc->v, fc->cap * sizeof *fc->v);
    }
    if (fc->text_len + len > fc->text_cap) {
        fc->text_cap = (fc->text_cap ? fc->text_cap * 2 : 256) + len;
        fc->text = (uint8_t *)realloc(fc->text, fc->text_cap);
    }
    memcpy(fc->text + fc->text_len, s, len);
    sxo_finding *f = &fc->v[fc->n++];
    f->position = pos;
    f->precision = (uint8_t)prec;
    f->completes_previous = (uint8_t)completes;
    f->input_file_id = (int16_t)file_id;
    f->mission_id = ss->m.mission_id;
    f->s_off = (uint32_t)fc->text_len;
    f->s_len = (uint32_t)len;
    fc->text_len += len;
}
size_t sxo_fc_len(const sxo_fc *fc) { return fc->n; }
const sxo_finding *sxo_fc_get(const sxo_fc *fc, size_t i) { return &fc->v[i]; }
const uint8_t *sxo_fc_text(const sxo_fc *fc) { return fc->text; }
uint64_t sxo_fc_first_byte_position(const sxo_fc *fc) { return fc->first_byte_position; }
int sxo_fc_str_buf_overflow(const sxo_fc *fc) { return fc->str_buf_overflow; }
void sxo_fc_free(sxo_fc *fc) {
    if (!fc) return;
    free(fc->v);
    free(fc->text);
    free(fc);
}

/* finding_collection.rs:84-342 */
static void from_into(sxo_fc *fc, sxo_state *ss, int file_id, const uint8_t *inp, size_t len, int is_last) {
    const size_t OUT = g_output_buf_len;
    /* finding_collection.rs:55: a zeroed buffer per call; slack so that decoders that were
     * told the true remaining length can never write past it */
    uint8_t *out = (uint8_t *)calloc(OUT + 8, 1);
    int extra_round = 0;
    size_t istart = 0, iend, ostart = 0, left_len = 0;
    if (ss->leftover_len) { /* :101-114 */
        memcpy(out, ss->leftover, ss->leftover_len);
        left_len = ss->leftover_len;
        ss->leftover_len = 0;
        ostart += left_len;
    }
    int cut = ss->cut;                                               /* :115 */
    const size_t W = 2 * (size_t)ss->m.output_line_char_nb_max;      /* :120 */
    int is_last_window = 0;
    while (istart < len) { /* :124 */
        if (istart + W < len)
            iend = istart + W;
        else {
            is_last_window = 1;
            iend = len;
        }
        for (;;) { /* 'decoder :134 */
            size_t rd = 0, wr = 0;
            int res = dec_decode(&ss->dec, inp + istart, iend - istart, out + ostart, OUT - ostart, extra_round, &rd,
                                 &wr);
            int prec = SXO_EXACT; /* :146 */
            if (wr > 0) {
                if (istart == 0 && (out[ostart] & 0x80)) { /* :176 */
                    sxo_decoder fresh;
                    dec_init(&fresh, ss->dec.enc, ss->dec.table);
                    uint8_t b8[16];
                    memset(b8, 0, sizeof b8);
                    size_t r2 = 0, w2 = 0;
                    dec_decode(&fresh, inp, len, b8, 8, 1, &r2, &w2); /* :190-194 */
                    if (w2 == 0 || memcmp(out, b8, w2) != 0) prec = SXO_BEFORE; /* :202-206 */
                }
            }
            size_t s0 = ostart, s1 = ostart + wr; /* :211-212 */
            if (left_len > 0) {                    /* :214-221 */
                s0 -= left_len;
                left_len = 0;
                prec = SXO_BEFORE;
            }
            int invalid_after = (res != R_INPUT_EMPTY && res != R_OUTPUT_FULL) || (is_last_window && is_last); /* :234 */
            int cont = cut; /* :240-241 */
            cut = 0;
            splitstr it = {out + s0,
                           s1 - s0,
                           0,
                           0,
                           ss->m.chars_min_nb,
                           ss->m.require_same_unicode_block,
                           cont,
                           invalid_after,
                           ss->m.af_lo,
                           ss->m.af_hi,
                           ss->m.ubf,
                           ss->m.grep_char,
                           ss->m.output_line_char_nb_max};
            sxo_split_result ch;
            while (splitstr_next(&it, &ch)) { /* :246 */
                if (!ch.again) {
                    fc_push(fc, ss, file_id, ss->consumed + istart, prec, out + s0 + ch.s_off, ch.s_len, ch.completes);
                    left_len = 0;
                    cut = ch.maybe_cut;
                } else {
                    left_len = ch.s_len;
                    cut = 0;
                }
                prec = SXO_AFTER; /* :289 */
            }
            ostart += wr; /* :292 */
            istart += rd; /* :294 */
            if (res == R_INPUT_EMPTY) {
                if (is_last_window && is_last && !extra_round)
                    extra_round = 1;
                else
                    break;
            } else if (res == R_OUTPUT_FULL) { /* :306-323 */
                fc->n = 0;
                fc->str_buf_overflow = 1;
                ostart = 0;
            }
        }
    }
    /* :330-338 */
    if (left_len > ss->leftover_cap) {
        ss->leftover_cap = left_len * 2 + 16;
        ss->leftover = (uint8_t *)realloc(ss->leftover, ss->leftover_cap);
    }
    if (left_len) memcpy(ss->leftover, out + ostart - left_len, left_len);
    ss->leftover_len = left_len;
    ss->cut = cut;
    ss->consumed += istart;
    free(out);
}

sxo_fc *sxo_from(sxo_state *ss, int file_id, const uint8_t *buf, size_t len, int is_last) {
    sxo_fc *fc = fc_new(ss->consumed);
    from_into(fc, ss, file_id, buf, len, is_last);
    return fc;
}

sxo_fc *sxo_scan_stream(sxo_state *ss, int file_id, const uint8_t *buf, size_t len, size_t slice_len, int is_last) {
    sxo_fc *fc = fc_new(ss->consumed);
    if (len == 0) {
        from_into(fc, ss, file_id, buf, 0, is_last);
        return fc;
    }
    for (size_t off = 0; off < len; off += slice_len) {
        size_t l = len - off < slice_len ? len - off : slice_len;
        int overflow_before = fc->str_buf_overflow;
        size_t n_before = fc->n;
        fc->str_buf_overflow = 0;
        /* a per-slice collection in the reference; overflow only clears that slice's findings */
        sxo_fc tmp = *fc;
        tmp.v = NULL;
        tmp.n = tmp.cap = 0;
        tmp.text = NULL;
        tmp.text_len = tmp.text_cap = 0;
        from_into(&tmp, ss, file_id, buf + off, l, is_last && off + l >= len);
        for (size_t i = 0; i < tmp.n; i++) {
            sxo_finding *f = &tmp.v[i];
            fc_push(fc, ss, f->input_file_id, f->position, f->precision, tmp.text + f->s_off, f->s_len,
                    f->completes_previous);
        }
        (void)n_before;
        fc->str_buf_overflow = overflow_before | tmp.str_buf_overflow;
        free(tmp.v);
        free(tmp.text);
    }
    return fc;
}

/* finding.rs:112-155 */
size_t sxo_print_finding(const sxo_finding *f, const uint8_t *text, const char *enc_name, int n_inputs, int n_missions,
                         int radix, int no_metadata, uint8_t *out, size_t cap) {
    char tmp[96];
    size_t k = 0;
#define PUT(p, n)                                                                                                      \
    do {                                                                                                               \
        if (k + (n) <= cap) memcpy(out + k, (p), (n));                                                                 \
        k += (n);                                                                                                      \
    } while (0)
    PUT("\n", 1);
    if (!no_metadata) {
        if (n_inputs > 1 && f->input_file_id >= 0) {
            tmp[0] = (char)(f->input_file_id + 64);
            tmp[1] = ' ';
            PUT(tmp, 2);
        }
        if (radix) {
            PUT(f->precision == SXO_AFTER ? ">" : f->precision == SXO_EXACT ? " " : "<", 1);
            int n = radix == 'x'   ? snprintf(tmp, sizeof tmp, "%llx", (unsigned long long)f->position)
                    : radix == 'd' ? snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)f->position)
                                   : snprintf(tmp, sizeof tmp, "%llo", (unsigned long long)f->position);
            PUT(tmp, (size_t)n);
            PUT(f->completes_previous ? "+\t" : " \t", 2);
        }
        if (n_missions > 1) {
            int n = snprintf(tmp, sizeof tmp, "(%c %s)\t", (char)(f->mission_id + 97), enc_name);
            PUT(tmp, (size_t)n);
        }
    }
    PUT(text + f->s_off, f->s_len);
#undef PUT
    return k;
}
