"""ctypes binding of the CPU ORACLE (oracle/sx_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under stringsext_b200/ may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsx_oracle.so")

ENC_X_USER_DEFINED, ENC_UTF_8, ENC_UTF_16LE, ENC_UTF_16BE, ENC_SINGLE_BYTE, ENC_UTF_32LE, ENC_UTF_32BE, ENC_BIG5, ENC_EUC_JP = range(9)
BEFORE, EXACT, AFTER = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sx_oracle.c")
    hdr = os.path.join(_HERE, "sx_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Mission(C.Structure):
    _fields_ = [
        ("mission_id", C.c_uint8),
        ("counter_offset", C.c_uint64),
        ("encoding_id", C.c_uint32),
        ("chars_min_nb", C.c_uint8),
        ("require_same_unicode_block", C.c_uint8),
        ("af_lo", C.c_uint64),
        ("af_hi", C.c_uint64),
        ("ubf", C.c_uint64),
        ("grep_char", C.c_int16),
        ("output_line_char_nb_max", C.c_uint32),
        ("print_encoding_as_ascii", C.c_uint8),
    ]


class _Finding(C.Structure):
    _fields_ = [
        ("position", C.c_uint64),
        ("precision", C.c_uint8),
        ("completes_previous", C.c_uint8),
        ("input_file_id", C.c_int16),
        ("mission_id", C.c_uint8),
        ("s_off", C.c_uint32),
        ("s_len", C.c_uint32),
    ]


class _Split(C.Structure):
    _fields_ = [
        ("s_off", C.c_uint32),
        ("s_len", C.c_uint32),
        ("completes", C.c_uint8),
        ("maybe_cut", C.c_uint8),
        ("again", C.c_uint8),
        ("min_ok", C.c_uint8),
        ("grep_ok", C.c_uint8),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.sxo_state_new.restype = C.c_void_p
        L.sxo_state_new.argtypes = [C.POINTER(_Mission), C.POINTER(C.c_uint16)]
        L.sxo_state_free.argtypes = [C.c_void_p]
        L.sxo_state_consumed.restype = C.c_uint64
        L.sxo_state_consumed.argtypes = [C.c_void_p]
        L.sxo_state_cut.argtypes = [C.c_void_p]
        L.sxo_state_leftover.restype = C.c_size_t
        L.sxo_state_leftover.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
        L.sxo_state_decoder_pending.restype = C.c_size_t
        L.sxo_state_decoder_pending.argtypes = [C.c_void_p, C.c_char_p]
        L.sxo_from.restype = C.c_void_p
        L.sxo_from.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_int]
        L.sxo_scan_stream.restype = C.c_void_p
        L.sxo_scan_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        L.sxo_fc_len.restype = C.c_size_t
        L.sxo_fc_len.argtypes = [C.c_void_p]
        L.sxo_fc_get.restype = C.POINTER(_Finding)
        L.sxo_fc_get.argtypes = [C.c_void_p, C.c_size_t]
        L.sxo_fc_text.restype = C.POINTER(C.c_uint8)
        L.sxo_fc_text.argtypes = [C.c_void_p]
        L.sxo_fc_first_byte_position.restype = C.c_uint64
        L.sxo_fc_first_byte_position.argtypes = [C.c_void_p]
        L.sxo_fc_str_buf_overflow.argtypes = [C.c_void_p]
        L.sxo_fc_free.argtypes = [C.c_void_p]
        L.sxo_split_str.restype = C.c_size_t
        L.sxo_split_str.argtypes = [
            C.c_char_p, C.c_size_t, C.c_uint8, C.c_int, C.c_int, C.c_int,
            C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_size_t, C.POINTER(_Split), C.c_size_t,
        ]
        L.sxo_char_count.restype = C.c_size_t
        L.sxo_char_count.argtypes = [C.c_char_p, C.c_size_t]
        L.sxo_set_output_buf_len.argtypes = [C.c_size_t]
        L.sxo_decoder_new.restype = C.c_void_p
        L.sxo_decoder_new.argtypes = [C.c_uint32, C.POINTER(C.c_uint16)]
        L.sxo_decoder_free.argtypes = [C.c_void_p]
        L.sxo_decoder_decode.argtypes = [
            C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int,
            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
        ]
        L.sxo_print_finding.restype = C.c_size_t
        L.sxo_print_finding.argtypes = [
            C.POINTER(_Finding), C.POINTER(C.c_uint8), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_char_p, C.c_size_t,
        ]
        _lib = L
    return _lib


@dataclass
class OMission:
    """Plain-data Mission (mission.rs:382-421) as the oracle sees it."""

    encoding_id: int
    chars_min_nb: int = 4
    af: int = 0
    ubf: int = 0
    grep_char: Optional[int] = None
    output_line_char_nb_max: int = 64
    require_same_unicode_block: bool = False
    counter_offset: int = 0
    mission_id: int = 0
    print_encoding_as_ascii: bool = False
    sb_table: Optional[Sequence[int]] = None  # 128 code points, 0 = unmapped
    encoding_name: str = ""

    def c(self) -> _Mission:
        return _Mission(
            self.mission_id, self.counter_offset, self.encoding_id, self.chars_min_nb,
            1 if self.require_same_unicode_block else 0,
            self.af & 0xFFFFFFFFFFFFFFFF, (self.af >> 64) & 0xFFFFFFFFFFFFFFFF, self.ubf,
            -1 if self.grep_char is None else self.grep_char, self.output_line_char_nb_max,
            1 if self.print_encoding_as_ascii else 0,
        )


@dataclass
class OFinding:
    position: int
    precision: int
    s: bytes
    completes: bool
    input_file_id: int = -1
    mission_id: int = 0

    def key(self):
        return (self.position, self.precision, self.s, self.completes)


@dataclass
class OCollection:
    v: List[OFinding] = field(default_factory=list)
    first_byte_position: int = 0
    str_buf_overflow: bool = False


def _collect(L, fc) -> OCollection:
    n = L.sxo_fc_len(fc)
    text = L.sxo_fc_text(fc)
    out = OCollection([], L.sxo_fc_first_byte_position(fc), bool(L.sxo_fc_str_buf_overflow(fc)))
    if n:
        total = 0
        last = L.sxo_fc_get(fc, n - 1).contents
        total = last.s_off + last.s_len
        blob = C.string_at(text, total)
        for i in range(n):
            f = L.sxo_fc_get(fc, i).contents
            out.v.append(
                OFinding(f.position, f.precision, blob[f.s_off : f.s_off + f.s_len], bool(f.completes_previous),
                         f.input_file_id, f.mission_id)
            )
    L.sxo_fc_free(fc)
    return out


class OState:
    """ScannerState (scanner.rs:40-89) in the oracle."""

    def __init__(self, m: OMission):
        L = lib()
        self.mission = m
        tab = None
        if m.sb_table is not None:
            tab = (C.c_uint16 * 128)(*m.sb_table)
        cm = m.c()
        self._h = L.sxo_state_new(C.byref(cm), tab)

    def __del__(self):
        try:
            if self._h:
                lib().sxo_state_free(self._h)
                self._h = None
        except Exception:
            pass

    def scan(self, buf: bytes, is_last: bool, file_id: int = -1) -> OCollection:
        """FindingCollection::from, one slice (finding_collection.rs:84)."""
        L = lib()
        return _collect(L, L.sxo_from(self._h, file_id, bytes(buf), len(buf), 1 if is_last else 0))

    def scan_stream(self, buf, is_last: bool = False, slice_len: int = 4096, file_id: int = -1) -> OCollection:
        """Fold of `scan` over slice_len pieces; buf may be bytes or a numpy uint8 array."""
        L = lib()
        if isinstance(buf, (bytes, bytearray)):
            b = bytes(buf)
            p = C.cast(C.c_char_p(b), C.c_void_p)
            n = len(b)
        else:  # numpy
            p = C.c_void_p(buf.ctypes.data)
            n = buf.size
        return _collect(L, L.sxo_scan_stream(self._h, file_id, p, n, slice_len, 1 if is_last else 0))

    @property
    def consumed_bytes(self) -> int:
        return lib().sxo_state_consumed(self._h)

    @property
    def cut(self) -> bool:
        return bool(lib().sxo_state_cut(self._h))

    @property
    def leftover(self) -> bytes:
        p = C.POINTER(C.c_uint8)()
        n = lib().sxo_state_leftover(self._h, C.byref(p))
        return C.string_at(p, n) if n else b""

    @property
    def decoder_pending(self) -> bytes:
        b = C.create_string_buffer(16)
        n = lib().sxo_state_decoder_pending(self._h, b)
        return b.raw[:n]


def split_str(s: bytes, chars_min_nb, same_block, last_cut, invalid_after, af, ubf, grep_char, s_char_nb_max):
    """SplitStr::new(...).collect() (helper.rs:171-432)."""
    L = lib()
    arr = (_Split * 256)()
    n = L.sxo_split_str(
        s, len(s), chars_min_nb, int(same_block), int(last_cut), int(invalid_after),
        af & 0xFFFFFFFFFFFFFFFF, (af >> 64) & 0xFFFFFFFFFFFFFFFF, ubf,
        -1 if grep_char is None else grep_char, s_char_nb_max, arr, 256,
    )
    return [
        dict(s=s[r.s_off : r.s_off + r.s_len], completes=bool(r.completes), maybe_cut=bool(r.maybe_cut),
             again=bool(r.again), min_ok=bool(r.min_ok), grep_ok=bool(r.grep_ok))
        for r in arr[:n]
    ]


def char_count(s: bytes) -> int:
    return lib().sxo_char_count(s, len(s))


def decode(encoding_id: int, chunks, dst_len: int = 1 << 16, last_on_final: bool = True, sb_table=None):
    """Feed chunks to one decoder; returns list of (result, read, written_bytes) per call."""
    L = lib()
    tab = (C.c_uint16 * 128)(*sb_table) if sb_table is not None else None
    d = L.sxo_decoder_new(encoding_id, tab)
    out = []
    for i, ch in enumerate(chunks):
        dst = C.create_string_buffer(dst_len + 8)
        rd, wr = C.c_size_t(0), C.c_size_t(0)
        res = L.sxo_decoder_decode(d, ch, len(ch), dst, dst_len, int(last_on_final and i == len(chunks) - 1),
                                   C.byref(rd), C.byref(wr))
        out.append((res, rd.value, dst.raw[: wr.value]))
    L.sxo_decoder_free(d)
    return out


def print_findings(findings: Sequence[OFinding], enc_names: Sequence[str], n_inputs: int, radix: Optional[str],
                   no_metadata: bool = False) -> bytes:
    """main.rs:116 + finding.rs:112-155 + main.rs:138: BOM, findings, trailing newline."""
    L = lib()
    out = bytearray(b"\xef\xbb\xbf")
    buf = C.create_string_buffer(1 << 16)
    for f in findings:
        cf = _Finding(f.position, f.precision, int(f.completes), f.input_file_id, f.mission_id, 0, len(f.s))
        txt = (C.c_uint8 * max(1, len(f.s))).from_buffer_copy(f.s if f.s else b"\0")
        n = L.sxo_print_finding(C.byref(cf), txt, enc_names[f.mission_id].encode(), n_inputs, len(enc_names),
                                ord(radix) if radix else 0, int(no_metadata), buf, len(buf))
        out += buf.raw[:n]
    out += b"\n"
    return bytes(out)


def cli_scan(missions: Sequence[OMission], files: Sequence[bytes], slice_len: int = 4096) -> List[OFinding]:
    """Driver of main.rs:93-175 with Slicer geometry (input.rs:104-168): per-file 4096-byte grid,
    carry flows across files, `is_last` never true, per-slice k-way merge by
    (position, mission_id) (finding.rs:92-109)."""
    states = [OState(m) for m in missions]
    merged: List[OFinding] = []
    for fid, data in enumerate(files, start=1):
        for off in range(0, len(data), slice_len):
            sl = data[off : off + slice_len]
            batch = []
            for ss in states:
                batch.extend(ss.scan(sl, False, fid).v)
            # each mission's list is position-monotone; stable sort == kmerge here
            batch.sort(key=lambda f: (f.position, f.mission_id))
            merged.extend(batch)
    return merged
