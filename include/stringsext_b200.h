/*
 * stringsext_b200.h -- C ABI of the B200-native stringsext scanner hot path.
 *
 * Drop-in boundary: the Rust function boundary the reference's driver uses at
 * /root/reference/src/main.rs:150-161, i.e.
 *     ScannerState::new(&'static Mission) -> ScannerState            scanner.rs:73-88
 *     FindingCollection::from(&mut ScannerState, Option<u8>, &[u8], bool)
 *             -> Pin<Box<FindingCollection>>                          finding_collection.rs:84-89
 *     fc.v[i].{position, position_precision, s, s_completes_previous_s, mission, input_file_id}
 *     fc.first_byte_position, fc.str_buf_overflow                     finding_collection.rs:31-50
 * Each entry point below names the reference interface it replaces.  All scanning runs in
 * hand-written sm_100a CUDA kernels; there is NO CPU fallback -- without a CUDA device every
 * scanning call fails and sx_last_error() says why.  INTEGRATION.md shows the Rust FFI stub.
 *
 * Plain pointers and sizes only; no torch / C++ types.
 */
#ifndef STRINGSEXT_B200_H
#define STRINGSEXT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Resolved encodings (Mission.encoding, mission.rs:397; `ascii` is x-user-defined + filter,
 * mission.rs:623-679).  Label parsing stays with the caller (CLI, out of scope). */
enum {
    SX_ENC_X_USER_DEFINED = 0,
    SX_ENC_UTF_8 = 1,
    SX_ENC_UTF_16LE = 2,
    SX_ENC_UTF_16BE = 3,
    SX_ENC_SINGLE_BYTE = 4, /* sb_table gives the upper half (koi8-r, ibm866, windows-125x, ...) */
    SX_ENC_UTF_32LE = 5,    /* extension: the reference rejects utf-32 (mission.rs:681-688) */
    SX_ENC_UTF_32BE = 6,
    SX_ENC_BIG5 = 7,   /* WHATWG Big5; index table generated from CPython's big5hkscs (tools/gen_multibyte_tables.py):
                        * self-consistent with the oracle, not reference-pinned (no WHATWG index offline) */
    SX_ENC_EUC_JP = 8  /* WHATWG EUC-JP; jis0208 / jis0212 from CPython's euc_jp, same caveat */
};

/* finding.rs:34-46 Precision */
enum { SX_PRECISION_BEFORE = 0, SX_PRECISION_EXACT = 1, SX_PRECISION_AFTER = 2 };

/* error codes returned by sx_last_error_code() */
enum {
    SX_OK = 0,
    SX_ERR_NO_DEVICE = 1,    /* no CUDA device / driver: the product never falls back to the CPU */
    SX_ERR_CUDA = 2,         /* a CUDA runtime call failed */
    SX_ERR_UNSUPPORTED = 3,  /* mission or geometry outside what the kernels implement */
    SX_ERR_ARGUMENT = 4
};

/* mission.rs:382-421 Mission + mission.rs:308-327 Utf8Filter, as plain data. */
typedef struct {
    uint8_t mission_id;
    uint64_t counter_offset;
    uint32_t encoding_id; /* SX_ENC_* */
    uint8_t chars_min_nb;
    uint8_t require_same_unicode_block;
    uint64_t af_lo, af_hi; /* Utf8Filter.af (u128): bit b set = ASCII code b passes */
    uint64_t ubf;          /* Utf8Filter.ubf: bit (utf8_lead_byte & 0x3f) set = passes */
    int16_t grep_char;     /* Utf8Filter.grep_char, -1 = None */
    uint32_t output_line_char_nb_max;
    uint8_t print_encoding_as_ascii;
    uint16_t sb_table[128]; /* SX_ENC_SINGLE_BYTE: code point of byte 0x80+i, 0 = unmapped */
} sx_mission;

/* finding.rs:51-74 Finding.  `s` points into the owning collection. */
typedef struct {
    uint64_t position;
    uint8_t precision;          /* SX_PRECISION_* */
    uint8_t completes_previous; /* s_completes_previous_s */
    int16_t input_file_id;      /* Option<u8>: -1 = None */
    uint8_t mission_id;
    const uint8_t* s; /* UTF-8, not NUL terminated */
    uint32_t s_len;
    int64_t in_start; /* reserved, 0 (an earlier revision reported the input range of the text here; it is no longer */
    uint32_t in_len;  /* shipped from the device: a finding crosses PCIe as 16 bytes + its text)                      */
} sx_finding;

typedef struct sx_scanner_state sx_scanner_state;
typedef struct sx_finding_collection sx_finding_collection;

/* Number of CUDA devices usable by the library; 0 when there is no driver/device. */
int sx_device_count(void);

/* ScannerState::new (scanner.rs:73-88).  `device`: CUDA ordinal the state's scans run on.
 * Returns NULL (and sets sx_last_error) for chars_min_nb == 0 and output_line_char_nb_max < 6 (options.rs:33) or
 * > 8192.  Missions with grep_char, require_same_unicode_block or chars_min_nb > output_line_char_nb_max take the
 * general automaton; of these only grep_char alone keeps the prefilter. */
sx_scanner_state* sx_scanner_state_new(const sx_mission* m, int device);
void sx_scanner_state_free(sx_scanner_state*);
/* Back to the state ScannerState::new leaves (scanner.rs:73-88), keeping the device buffers. */
void sx_scanner_state_reset(sx_scanner_state*);

/* ScannerState fields (scanner.rs:55-68). */
uint64_t sx_scanner_state_consumed_bytes(const sx_scanner_state*);
int sx_scanner_state_maybe_cut(const sx_scanner_state*); /* last_run_str_was_printed_and_is_maybe_cut_str */
size_t sx_scanner_state_leftover(const sx_scanner_state*, const uint8_t** utf8); /* last_scan_run_leftover */

/* FindingCollection::from (finding_collection.rs:84-89): ONE slice (<= 4096 bytes in the
 * reference, input.rs:22), host pointer, exact reference semantics -- executed on the GPU. */
sx_finding_collection* sx_finding_collection_from(sx_scanner_state*, int input_file_id, const uint8_t* buf,
                                                  size_t len, int is_last_input_buffer);

/* Batched form, the GPU entry point: semantically the fold of sx_finding_collection_from over
 * consecutive slice_len pieces of buf (main.rs:153-167 for one mission), findings concatenated,
 * `is_last` applied to the final slice.  buf is a host pointer (buf_is_device == 0; the copy to
 * the device is part of the call) or a device pointer on the state's device.
 * cuda_stream: a cudaStream_t or NULL for the default stream.
 * Leaves the state exactly as the fold would, so calls chain (across buffers and files). */
sx_finding_collection* sx_scan_stream(sx_scanner_state*, int input_file_id, const void* buf, size_t len,
                                      size_t slice_len, int buf_is_device, int is_last, void* cuda_stream);

/* One RANGE of a resident stream: the findings sx_scan_stream would emit while it processes the bytes [lo, hi) of buf
 * (lo, hi multiples of slice_len, or hi == len), in the same order -- so the collections of consecutive ranges
 * concatenate to the collection of the whole call.  This is how one mission is spread over several GPUs / streams
 * (SURVEY.md 8(e)(2)): the reference gives a mission one thread (main.rs:151-167), here every device scans its own
 * range.  The carry into the range's first window is derived on the device by walking back to the nearest window
 * whose carry-out does not depend on its carry-in (on binary input: the window before the range).
 * flags & SX_RANGE_PREFIX_UNKNOWN: buf does not start where the state stands but somewhere earlier in the stream
 * (a halo in front of the range, lo > 0; the state's counter_offset / consumed bytes must equal the stream offset
 * of buf[0]); the call fails (SX_ERR_UNSUPPORTED) if the walk back would have to leave the halo.
 * The ScannerState only moves on when hi == len (then exactly as sx_scan_stream would leave it). */
enum { SX_RANGE_PREFIX_UNKNOWN = 1 };
sx_finding_collection* sx_scan_range(sx_scanner_state*, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                     int buf_is_device, int is_last, size_t lo, size_t hi, int flags, void* cuda_stream);

/* Asynchronous forms.  The reference overlaps scanning with merging / printing: every mission has a scanner thread
 * that sends its collections through a bounded channel to the merger thread (main.rs:98 sync_channel, :161 tx.send,
 * :103-141 merger).  Here every state owns a worker thread: the *_async calls return a handle at once, the scans of
 * one state run in call order (so chained calls see the ScannerState the previous call left), scans of different
 * states -- on the same or on different GPUs -- run side by side.  sx_fc_wait blocks until the collection is complete,
 * returns it (NULL on error: sx_last_error* of the WAITING thread is set) and consumes the handle.  buf must stay
 * valid until the wait returns. */
typedef struct sx_pending sx_pending;
sx_pending* sx_scan_stream_async(sx_scanner_state*, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                 int buf_is_device, int is_last, void* cuda_stream);
sx_pending* sx_scan_range_async(sx_scanner_state*, int input_file_id, const void* buf, size_t len, size_t slice_len,
                                int buf_is_device, int is_last, size_t lo, size_t hi, int flags, void* cuda_stream);
int sx_pending_ready(const sx_pending*); /* 1: sx_fc_wait will not block */
sx_finding_collection* sx_fc_wait(sx_pending*);

/* Streaming driver: the loop of main.rs:143-167 over the input iterator of input.rs:104-168, for the GPU.  One input is
 * pulled through `read` (returns the bytes it stored, 0 at the end of the input; short reads are fine) in pieces of
 * chunk_bytes (a positive multiple of 4096: the slice grid inside the library is then the reference's, and restarts
 * with every input like input.rs), each piece is uploaded once per device (double buffered: the read and the upload of
 * piece k + 1 overlap the scans of piece k) and scanned by every state side by side; `batch` gets the piece's
 * collections in state order and owns them (sx_fc_free) -- sx_merge over them gives the order of main.rs:118-136;
 * a non-zero return of `batch` stops the run and is returned.  Carry, pending decoder bytes and byte counters stay in
 * the states: call once per input with the reference's 1-based file label (-1: none).  The "last input buffer" flag is
 * never set, like the reference's CLI (input.rs:130-137).  Returns 0, batch's non-zero value, or -1 on error. */
typedef size_t (*sx_read_fn)(void* user, uint8_t* dst, size_t cap);
typedef int (*sx_batch_fn)(void* user, int input_file_id, sx_finding_collection* const* fcs, size_t n_states);
int sx_scan_reader(sx_scanner_state* const* states, size_t n_states, int input_file_id, sx_read_fn read, void* read_user,
                   size_t chunk_bytes, sx_batch_fn batch, void* batch_user);
/* The same with a file (path == NULL: stdin); an unreadable file is reported on stderr and scanned as an empty input
 * (input.rs:78-84). */
int sx_scan_file(sx_scanner_state* const* states, size_t n_states, int input_file_id, const char* path, size_t chunk_bytes,
                 sx_batch_fn batch, void* batch_user);

/* FindingCollection accessors (finding_collection.rs:31-50, :371-415). */
size_t sx_fc_len(const sx_finding_collection*);
const sx_finding* sx_fc_get(const sx_finding_collection*, size_t i);
const sx_finding* sx_fc_data(const sx_finding_collection*); /* contiguous array of sx_fc_len() findings */
uint64_t sx_fc_first_byte_position(const sx_finding_collection*);
int sx_fc_str_buf_overflow(const sx_finding_collection*);
void sx_fc_free(sx_finding_collection*);

/* k-way merge of the collections of one slice batch in the order of `impl PartialOrd for
 * Finding` (finding.rs:92-109) as used by main.rs:133: fills `out` (capacity = sum of lengths)
 * with pointers to the findings; returns the count. */
size_t sx_merge(const sx_finding_collection* const* fcs, size_t n, const sx_finding** out);

/* Instrumentation of the most recent sx_scan_stream on this state (CUDA-event times of the
 * library's own kernels, measured on the stream they were launched on). */
typedef struct {
    float scan_kernel_ms;        /* prefilter + list + exact kernels, first launch to last */
    float prefilter_kernel_ms;   /* sx_prefilter_kernel: the one pass over every input byte (HBM bound) */
    float list_kernels_ms;       /* sx_list_scan_kernel + sx_list_expand_kernel */
    float exact_kernel_ms;       /* exact stage over the listed windows: sx_exact_kernel, or the sparse pipeline's kernels
                                  * incl. the host round trip for the list length */
    float materialize_kernel_ms; /* sx_materialize_kernel (finding text) */
    uint32_t kernel_launches;    /* launches of library kernels in the call */
    uint32_t relaunches;         /* pipeline re-runs caused by an output buffer that was too small */
    uint32_t prefilter_used;
    uint32_t tma_used;           /* the prefilter staged its tiles with cp.async.bulk.tensor */
    uint32_t sparse_used;        /* the exact stage ran as the barrier-free sparse-list pipeline (UTF-8, single-byte) */
    uint32_t pieces;             /* pieces the call was cut into (prefilter of piece k+1 overlaps exact stage + download of piece k) */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t n_records, text_bytes;
    uint64_t windows_total, windows_listed;
    float host_total_ms; /* wall clock of the whole call */
    float host_post_ms;  /* of which: building the collection after the last device sync */
    float sparse_stage_ms[6]; /* sparse pipeline (sparse_used), summed over the pieces: list compaction, heads, members,
                               * fix+late, ext, scan+gather kernels */
    float host_phase_ms[4]; /* wall clock: [0] call start -> kernels enqueued, [1] -> counters back (first sync),
                             * [2] -> results downloaded (second sync), [3] -> collection built */
} sx_scan_stats;
void sx_scanner_state_last_stats(const sx_scanner_state*, sx_scan_stats* out);

/* Tuning / test hooks.  The prefilter never changes results (tests compare both settings). */
void sx_scanner_state_set_prefilter(sx_scanner_state*, int enabled);
void sx_scanner_state_set_tma(sx_scanner_state*, int enabled); /* 0: stage tiles with plain vector loads */
void sx_scanner_state_set_sparse(sx_scanner_state*, int mode); /* 0: always the block kernel for the exact stage; 1 (default):
                                                                 * UTF-8 missions take the per-stage pipeline whenever its
                                                                 * per-entry state fits in memory */
void sx_scanner_state_set_direct_output(sx_scanner_state*, int enabled); /* 0: download records, convert on the host */
void sx_scanner_state_set_pieces(sx_scanner_state*, int pieces); /* sparse pipeline: cut every call into this many pieces
                                                                  * (0: automatic, by size); results never depend on it */
/* Copies the window list the prefilter built in the most recent call (ascending window indices,
 * window = decoder_input_window of finding_collection.rs:120-131) into out[0..cap); returns the
 * list length, 0 when the prefilter did not run. */
size_t sx_scanner_state_last_window_list(const sx_scanner_state*, uint32_t* out, size_t cap);

/* Test / benchmark support: fill a device buffer with the library's reproducible synthetic
 * corpus bytes (counter based; byte i depends on (seed, i) only; see bench.py). */
int sx_fill_random(void* device_buf, size_t len, uint64_t seed, uint64_t stream_offset, int device,
                   void* cuda_stream);

int sx_last_error_code(void);
const char* sx_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
