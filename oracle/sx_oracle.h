/*
 * sx_oracle.h -- CPU ORACLE for the stringsext scanner hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (stringsext_b200/, include/) never links, imports or calls it.
 *
 * It is a plain-C restatement of the reference algorithm (getreu/stringsext v2.3.5):
 *   FindingCollection::from      src/finding_collection.rs:84-342
 *   SplitStr::next               src/helper.rs:210-432
 *   Utf8Filter::pass_*_filter    src/mission.rs:329-349
 *   ScannerState::new            src/scanner.rs:71-89
 *   Slicer::next (geometry)      src/input.rs:104-168
 *   merge order / print          src/finding.rs:92-155, src/main.rs:103-141
 * plus a restatement of the decoders of the un-vendored crate encoding_rs 0.8.34
 * (Cargo.toml:21, Cargo.lock:147-150) that the reference calls at
 * finding_collection.rs:138-143 and :180-194.
 *
 * Parity pinning: see oracle/README.md.  UTF-8, UTF-16LE/BE and x-user-defined are pinned
 * by the reference's own unit tests and CLI golden files (tests/golden/).  Single-byte
 * legacy tables, UTF-32 (an extension, the reference has none) and a few decoder corners
 * are "parity unpinned" (listed in oracle/README.md and DESIGN.md).
 */
#ifndef SX_ORACLE_H
#define SX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Encoding ids (resolved, not labels; label parsing is CLI work and out of scope). */
enum {
    SXO_ENC_X_USER_DEFINED = 0, /* also the `ascii` emulation, mission.rs:623-679 */
    SXO_ENC_UTF_8 = 1,
    SXO_ENC_UTF_16LE = 2,
    SXO_ENC_UTF_16BE = 3,
    SXO_ENC_SINGLE_BYTE = 4, /* table driven (koi8-r, ibm866, ...); table passed at state creation */
    SXO_ENC_UTF_32LE = 5,    /* EXTENSION: not in encoding_rs / the reference */
    SXO_ENC_UTF_32BE = 6,    /* EXTENSION */
    SXO_ENC_BIG5 = 7,        /* WHATWG Big5 decoder over the library's generated index (parity unpinned, see sx_oracle.c) */
    SXO_ENC_EUC_JP = 8       /* WHATWG EUC-JP decoder, same caveat */
};

enum { SXO_BEFORE = 0, SXO_EXACT = 1, SXO_AFTER = 2 }; /* finding.rs:34-46 */

/* mission.rs:382-421 + Utf8Filter mission.rs:308-327, resolved to plain data. */
typedef struct {
    uint8_t mission_id;
    uint64_t counter_offset;
    uint32_t encoding_id;
    uint8_t chars_min_nb;
    uint8_t require_same_unicode_block;
    uint64_t af_lo, af_hi; /* u128 af, bit b = ASCII code b passes */
    uint64_t ubf;          /* bit (lead & 0x3f) passes */
    int16_t grep_char;     /* -1 = None */
    uint32_t output_line_char_nb_max;
    uint8_t print_encoding_as_ascii;
} sxo_mission;

/* finding.rs:51-74 */
typedef struct {
    uint64_t position;
    uint8_t precision;
    uint8_t completes_previous;
    int16_t input_file_id; /* -1 = None */
    uint8_t mission_id;
    uint32_t s_off; /* offset into the collection's text arena */
    uint32_t s_len;
} sxo_finding;

typedef struct sxo_state sxo_state;
typedef struct sxo_fc sxo_fc;

/* helper.rs:127-168 */
typedef struct {
    uint32_t s_off, s_len;
    uint8_t completes, maybe_cut, again, min_ok, grep_ok;
} sxo_split_result;

/* test hook: finding.rs:23-25 OUTPUT_BUF_LEN (0x9192 in production, 0x40 under cfg(test)) */
void sxo_set_output_buf_len(size_t n);

/* scanner.rs:73-88.  `sb_table`: 128 u16 code points for bytes 0x80..0xFF (0 = unmapped),
 * only read for SXO_ENC_SINGLE_BYTE. */
sxo_state *sxo_state_new(const sxo_mission *m, const uint16_t *sb_table);
void sxo_state_free(sxo_state *);
uint64_t sxo_state_consumed(const sxo_state *);
int sxo_state_cut(const sxo_state *);
size_t sxo_state_leftover(const sxo_state *, const uint8_t **p);
/* opaque decoder pending state, for hand-off tests: fills up to 8 bytes, returns count */
size_t sxo_state_decoder_pending(const sxo_state *, uint8_t *out8);

/* finding_collection.rs:84-342, one slice, literal. */
sxo_fc *sxo_from(sxo_state *, int input_file_id, const uint8_t *buf, size_t len, int is_last);
/* Fold of sxo_from over consecutive slice_len pieces (input.rs:104-168 geometry for one
 * file); `is_last` is applied to the final slice only.  len == 0 is a no-op collection. */
sxo_fc *sxo_scan_stream(sxo_state *, int input_file_id, const uint8_t *buf, size_t len, size_t slice_len,
                        int is_last);

size_t sxo_fc_len(const sxo_fc *);
const sxo_finding *sxo_fc_get(const sxo_fc *, size_t i);
const uint8_t *sxo_fc_text(const sxo_fc *);
uint64_t sxo_fc_first_byte_position(const sxo_fc *);
int sxo_fc_str_buf_overflow(const sxo_fc *);
void sxo_fc_free(sxo_fc *);

/* helper.rs:171-432 exposed for the reference's SplitStr unit tests.
 * Returns the number of results written (iteration until None), at most `max`. */
size_t sxo_split_str(const uint8_t *s, size_t len, uint8_t chars_min_nb, int same_block, int last_cut,
                     int invalid_after, uint64_t af_lo, uint64_t af_hi, uint64_t ubf, int grep_char,
                     size_t s_char_nb_max, sxo_split_result *out, size_t max);
size_t sxo_char_count(const uint8_t *s, size_t len); /* helper.rs:445-461 */

/* Raw decoder access (encoding_rs Decoder::decode_to_str_without_replacement) for tests.
 * returns 0 InputEmpty, 1 OutputFull, 2 Malformed */
typedef struct sxo_decoder sxo_decoder;
sxo_decoder *sxo_decoder_new(uint32_t encoding_id, const uint16_t *sb_table);
void sxo_decoder_free(sxo_decoder *);
int sxo_decoder_decode(sxo_decoder *, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, int last,
                       size_t *read, size_t *written);

/* finding.rs:112-155 + main.rs:116,:138.  Appends the printed bytes of one finding to
 * out (cap bytes available), returns bytes written.  radix: 'x','d','o' or 0 (None). */
size_t sxo_print_finding(const sxo_finding *f, const uint8_t *text, const char *enc_name, int n_inputs,
                         int n_missions, int radix, int no_metadata, uint8_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
