"""ScannerState / FindingCollection / Finding: Python mirror of the reference's scanner API,
bound to the CUDA library through its C ABI (include/stringsext_b200.h).

Reference interface mirrored here:
  * `ScannerState::new(mission)`                         /root/reference/src/scanner.rs:73-88
  * `FindingCollection::from(ss, file_id, buf, is_last)` /root/reference/src/finding_collection.rs:84-89
  * `Finding` / `Precision`                              /root/reference/src/finding.rs:34-74
  * merge order (`impl PartialOrd for Finding`)          /root/reference/src/finding.rs:92-109
  * `Finding::print`                                     /root/reference/src/finding.rs:112-155

All scanning happens in the CUDA kernels; if the library or a CUDA device is missing the calls
raise -- there is deliberately no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

from .mission import Mission

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstringsext_b200.so")


class ScannerError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"stringsext_b200 error {code}: {msg}")
        self.code = code


class Precision(enum.IntEnum):  # finding.rs:34-46
    Before = 0
    Exact = 1
    After = 2


class _CMission(C.Structure):
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
        ("sb_table", C.c_uint16 * 128),
    ]


class _CFinding(C.Structure):
    _fields_ = [
        ("position", C.c_uint64),
        ("precision", C.c_uint8),
        ("completes_previous", C.c_uint8),
        ("input_file_id", C.c_int16),
        ("mission_id", C.c_uint8),
        ("s", C.POINTER(C.c_uint8)),
        ("s_len", C.c_uint32),
        ("in_start", C.c_int64),
        ("in_len", C.c_uint32),
    ]


RANGE_PREFIX_UNKNOWN = 1  # SX_RANGE_PREFIX_UNKNOWN


class ScanStats(C.Structure):
    _fields_ = [
        ("scan_kernel_ms", C.c_float),
        ("prefilter_kernel_ms", C.c_float),
        ("list_kernels_ms", C.c_float),
        ("exact_kernel_ms", C.c_float),
        ("materialize_kernel_ms", C.c_float),
        ("kernel_launches", C.c_uint32),
        ("relaunches", C.c_uint32),
        ("prefilter_used", C.c_uint32),
        ("tma_used", C.c_uint32),
        ("sparse_used", C.c_uint32),
        ("pieces", C.c_uint32),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("n_records", C.c_uint64),
        ("text_bytes", C.c_uint64),
        ("windows_total", C.c_uint64),
        ("windows_listed", C.c_uint64),
        ("host_total_ms", C.c_float),
        ("host_post_ms", C.c_float),
        ("sparse_stage_ms", C.c_float * 6),
        ("host_phase_ms", C.c_float * 4),
    ]


READ_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)          # sx_read_fn
BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_size_t)  # sx_batch_fn

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScannerError(2, f"{LIB_PATH} is missing: build it with `python -m stringsext_b200.build` "
                              "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.sx_device_count.restype = C.c_int
    L.sx_scanner_state_new.restype = C.c_void_p
    L.sx_scanner_state_new.argtypes = [C.POINTER(_CMission), C.c_int]
    L.sx_scanner_state_free.argtypes = [C.c_void_p]
    L.sx_scanner_state_reset.argtypes = [C.c_void_p]
    L.sx_scanner_state_consumed_bytes.restype = C.c_uint64
    L.sx_scanner_state_consumed_bytes.argtypes = [C.c_void_p]
    L.sx_scanner_state_maybe_cut.argtypes = [C.c_void_p]
    L.sx_scanner_state_leftover.restype = C.c_size_t
    L.sx_scanner_state_leftover.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
    L.sx_scanner_state_last_stats.argtypes = [C.c_void_p, C.POINTER(ScanStats)]
    L.sx_scanner_state_set_prefilter.argtypes = [C.c_void_p, C.c_int]
    L.sx_scanner_state_set_tma.argtypes = [C.c_void_p, C.c_int]
    L.sx_scanner_state_set_sparse.argtypes = [C.c_void_p, C.c_int]
    L.sx_scanner_state_set_direct_output.argtypes = [C.c_void_p, C.c_int]
    L.sx_scanner_state_set_pieces.argtypes = [C.c_void_p, C.c_int]
    L.sx_scan_range.restype = C.c_void_p
    L.sx_scan_range.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                C.c_int, C.c_void_p]
    L.sx_scanner_state_last_window_list.restype = C.c_size_t
    L.sx_scanner_state_last_window_list.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t]
    L.sx_finding_collection_from.restype = C.c_void_p
    L.sx_finding_collection_from.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_int]
    L.sx_scan_stream.restype = C.c_void_p
    L.sx_scan_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.sx_scan_stream_async.restype = C.c_void_p
    L.sx_scan_stream_async.argtypes = L.sx_scan_stream.argtypes
    L.sx_scan_range_async.restype = C.c_void_p
    L.sx_scan_range_async.argtypes = L.sx_scan_range.argtypes
    L.sx_pending_ready.argtypes = [C.c_void_p]
    L.sx_fc_wait.restype = C.c_void_p
    L.sx_fc_wait.argtypes = [C.c_void_p]
    L.sx_scan_reader.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, READ_FN, C.c_void_p, C.c_size_t, BATCH_FN, C.c_void_p]
    L.sx_scan_file.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, BATCH_FN, C.c_void_p]
    L.sx_fc_len.restype = C.c_size_t
    L.sx_fc_len.argtypes = [C.c_void_p]
    L.sx_fc_get.restype = C.POINTER(_CFinding)
    L.sx_fc_get.argtypes = [C.c_void_p, C.c_size_t]
    L.sx_fc_data.restype = C.POINTER(_CFinding)
    L.sx_fc_data.argtypes = [C.c_void_p]
    L.sx_fc_first_byte_position.restype = C.c_uint64
    L.sx_fc_first_byte_position.argtypes = [C.c_void_p]
    L.sx_fc_str_buf_overflow.argtypes = [C.c_void_p]
    L.sx_fc_free.argtypes = [C.c_void_p]
    L.sx_merge.restype = C.c_size_t
    L.sx_merge.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.POINTER(_CFinding))]
    L.sx_fill_random.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    L.sx_last_error_code.restype = C.c_int
    L.sx_last_error.restype = C.c_char_p
    _lib = L
    return L


def _raise_last():
    L = load_library()
    raise ScannerError(L.sx_last_error_code(), (L.sx_last_error() or b"").decode())


def device_count() -> int:
    return load_library().sx_device_count()


def exported_symbols() -> List[str]:
    """Entry points declared in include/stringsext_b200.h (used by the ABI load test)."""
    return [
        "sx_device_count", "sx_scanner_state_new", "sx_scanner_state_free", "sx_scanner_state_reset", "sx_scanner_state_consumed_bytes",
        "sx_scanner_state_maybe_cut", "sx_scanner_state_leftover", "sx_finding_collection_from", "sx_scan_stream",
        "sx_fc_len", "sx_fc_get", "sx_fc_data", "sx_fc_first_byte_position", "sx_fc_str_buf_overflow", "sx_fc_free",
        "sx_merge", "sx_scanner_state_last_stats", "sx_scanner_state_set_prefilter", "sx_scanner_state_set_tma", "sx_scanner_state_set_sparse", "sx_scanner_state_set_direct_output", "sx_scanner_state_set_pieces", "sx_scan_range", "sx_scan_stream_async", "sx_scan_range_async", "sx_pending_ready", "sx_fc_wait",
        "sx_scan_reader", "sx_scan_file", "sx_scanner_state_last_window_list", "sx_fill_random", "sx_last_error_code", "sx_last_error",
    ]


@dataclass
class Finding:  # finding.rs:51-74
    position: int
    position_precision: Precision
    s: bytes
    s_completes_previous_s: bool
    mission: Mission
    input_file_id: Optional[int] = None
    in_start: int = 0
    in_len: int = 0

    def sort_key(self):  # finding.rs:92-109
        return (self.position, self.mission.mission_id, self.mission.filter.ubf, self.mission.filter.af)

    def print(self, n_inputs: int = 1, n_missions: int = 1, radix: Optional[str] = None,
              no_metadata: bool = False) -> bytes:
        """finding.rs:112-155 (the global ARGS become parameters)."""
        out = bytearray(b"\n")
        if not no_metadata:
            if n_inputs > 1 and self.input_file_id is not None:
                out += bytes([self.input_file_id + 64, 0x20])
            if radix is not None:
                out += {Precision.After: b">", Precision.Exact: b" ", Precision.Before: b"<"}[self.position_precision]
                out += {"x": "%x", "d": "%d", "o": "%o"}[radix.lower()].encode() % self.position
                out += b"+\t" if self.s_completes_previous_s else b" \t"
            if n_missions > 1:
                out += b"(" + bytes([self.mission.mission_id + 97]) + b" " + self.mission.printed_encoding_name.encode() + b")\t"
        out += self.s
        return bytes(out)


class FindingCollection:  # finding_collection.rs:31-50
    def __init__(self, v: List[Finding], first_byte_position: int, str_buf_overflow: bool):
        self.v = v
        self.first_byte_position = first_byte_position
        self.str_buf_overflow = str_buf_overflow

    def __iter__(self):
        return iter(self.v)

    def __len__(self):
        return len(self.v)


class ScannerState:
    """scanner.rs:40-89.  One per mission; scans run on CUDA device `device`."""

    def __init__(self, mission: Mission, device: int = 0):
        L = load_library()
        self.mission = mission
        self.device = device
        cm = _CMission()
        cm.mission_id = mission.mission_id
        cm.counter_offset = mission.counter_offset
        cm.encoding_id = mission.encoding_id
        cm.chars_min_nb = mission.chars_min_nb
        cm.require_same_unicode_block = 1 if mission.require_same_unicode_block else 0
        cm.af_lo = mission.filter.af & 0xFFFFFFFFFFFFFFFF
        cm.af_hi = (mission.filter.af >> 64) & 0xFFFFFFFFFFFFFFFF
        cm.ubf = mission.filter.ubf
        cm.grep_char = -1 if mission.filter.grep_char is None else mission.filter.grep_char
        cm.output_line_char_nb_max = mission.output_line_char_nb_max
        cm.print_encoding_as_ascii = 1 if mission.print_encoding_as_ascii else 0
        if mission.sb_table is not None:
            for i, v in enumerate(mission.sb_table):
                cm.sb_table[i] = v
        self._h = L.sx_scanner_state_new(C.byref(cm), device)
        if not self._h:
            _raise_last()

    def close(self):
        if getattr(self, "_h", None):
            load_library().sx_scanner_state_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- ScannerState fields -----------------------------------------------------------------
    @property
    def consumed_bytes(self) -> int:
        return load_library().sx_scanner_state_consumed_bytes(self._h)

    @property
    def last_run_str_was_printed_and_is_maybe_cut_str(self) -> bool:
        return bool(load_library().sx_scanner_state_maybe_cut(self._h))

    @property
    def last_scan_run_leftover(self) -> bytes:
        p = C.POINTER(C.c_uint8)()
        n = load_library().sx_scanner_state_leftover(self._h, C.byref(p))
        return C.string_at(p, n) if n else b""

    def reset(self) -> None:
        load_library().sx_scanner_state_reset(self._h)

    def set_prefilter(self, enabled: bool) -> None:
        load_library().sx_scanner_state_set_prefilter(self._h, 1 if enabled else 0)

    def set_tma(self, enabled: bool) -> None:
        load_library().sx_scanner_state_set_tma(self._h, 1 if enabled else 0)

    def set_direct_output(self, enabled: bool) -> None:
        """False: records are downloaded and converted on the host instead of being written by the GPU as findings."""
        load_library().sx_scanner_state_set_direct_output(self._h, 1 if enabled else 0)

    def set_pieces(self, pieces: int) -> None:
        """Cut every call of the sparse pipeline into this many pieces (0: automatic); results never depend on it."""
        load_library().sx_scanner_state_set_pieces(self._h, int(pieces))

    def set_sparse(self, enabled: bool) -> None:
        """False: the exact stage always runs as the block kernel (sx_exact_kernel), never as the sparse-list pipeline."""
        load_library().sx_scanner_state_set_sparse(self._h, int(enabled))  # 0 off, 1 default, 2 whenever possible

    def last_window_list(self) -> List[int]:
        L = load_library()
        n = L.sx_scanner_state_last_window_list(self._h, None, 0)
        if n == 0:
            return []
        arr = (C.c_uint32 * n)()
        L.sx_scanner_state_last_window_list(self._h, arr, n)
        return list(arr)

    @property
    def last_stats(self) -> ScanStats:
        st = ScanStats()
        load_library().sx_scanner_state_last_stats(self._h, C.byref(st))
        return st

    # -- scanning ------------------------------------------------------------------------------
    def _collect(self, fc, file_id) -> FindingCollection:
        L = load_library()
        if not fc:
            _raise_last()
        n = L.sx_fc_len(fc)
        arr = L.sx_fc_data(fc)
        v = []
        m = self.mission
        for i in range(n):
            f = arr[i]
            v.append(Finding(f.position, Precision(f.precision), C.string_at(f.s, f.s_len), bool(f.completes_previous),
                             m, None if f.input_file_id < 0 else f.input_file_id, f.in_start, f.in_len))
        out = FindingCollection(v, L.sx_fc_first_byte_position(fc), bool(L.sx_fc_str_buf_overflow(fc)))
        L.sx_fc_free(fc)
        return out

    def scan(self, input_buffer: bytes, is_last_input_buffer: bool, input_file_id: Optional[int] = None) -> FindingCollection:
        """FindingCollection::from: one slice, exact reference semantics, executed on the GPU."""
        L = load_library()
        fid = -1 if input_file_id is None else input_file_id
        b = bytes(input_buffer)
        return self._collect(L.sx_finding_collection_from(self._h, fid, b, len(b), 1 if is_last_input_buffer else 0), fid)

    def scan_stream(self, buf, is_last: bool = False, slice_len: int = 4096, input_file_id: Optional[int] = None,
                    device_ptr: Optional[int] = None, length: Optional[int] = None, cuda_stream: int = 0,
                    raw: bool = False, lo: Optional[int] = None, hi: Optional[int] = None, prefix_unknown: bool = False):
        """The fold of `scan` over slice_len pieces.  `buf`: bytes / numpy uint8 array (host) or, with
        `device_ptr`+`length`, a device pointer on this state's device.  raw=True returns the C
        collection handle wrapped in RawCollection (no per-finding Python objects)."""
        L = load_library()
        fid = -1 if input_file_id is None else input_file_id
        keep = None
        if device_ptr is not None:
            p, n, isdev = C.c_void_p(device_ptr), int(length), 1
        elif isinstance(buf, (bytes, bytearray)):
            keep = bytes(buf)
            p, n, isdev = C.cast(C.c_char_p(keep), C.c_void_p), len(keep), 0
        else:  # numpy array or anything with ctypes.data / nbytes
            p, n, isdev = C.c_void_p(buf.ctypes.data), int(buf.nbytes if length is None else length), 0
        if lo is None and hi is None:
            fc = L.sx_scan_stream(self._h, fid, p, n, slice_len, isdev, 1 if is_last else 0, C.c_void_p(cuda_stream))
        else:  # sx_scan_range: only the findings emitted while the bytes [lo, hi) are processed
            fc = L.sx_scan_range(self._h, fid, p, n, slice_len, isdev, 1 if is_last else 0, int(lo or 0), n if hi is None else int(hi),
                                 RANGE_PREFIX_UNKNOWN if prefix_unknown else 0, C.c_void_p(cuda_stream))
        del keep
        if raw:
            if not fc:
                _raise_last()
            return RawCollection(fc)
        return self._collect(fc, fid)


    def scan_stream_async(self, buf=None, is_last: bool = False, slice_len: int = 4096, input_file_id: Optional[int] = None,
                          device_ptr: Optional[int] = None, length: Optional[int] = None, cuda_stream: int = 0,
                          lo: Optional[int] = None, hi: Optional[int] = None, prefix_unknown: bool = False) -> "PendingScan":
        """sx_scan_stream_async / sx_scan_range_async: returns at once; `.wait()` gives the collection.  Scans of one state
        run in call order, scans of different states side by side (main.rs:98-167: scanner threads + channel)."""
        L = load_library()
        fid = -1 if input_file_id is None else input_file_id
        keep = None
        if device_ptr is not None:
            p, n, isdev = C.c_void_p(device_ptr), int(length), 1
        elif isinstance(buf, (bytes, bytearray)):
            keep = bytes(buf)
            p, n, isdev = C.cast(C.c_char_p(keep), C.c_void_p), len(keep), 0
        else:
            keep = buf
            p, n, isdev = C.c_void_p(buf.ctypes.data), int(buf.nbytes if length is None else length), 0
        if lo is None and hi is None:
            h = L.sx_scan_stream_async(self._h, fid, p, n, slice_len, isdev, 1 if is_last else 0, C.c_void_p(cuda_stream))
        else:
            h = L.sx_scan_range_async(self._h, fid, p, n, slice_len, isdev, 1 if is_last else 0, int(lo or 0), n if hi is None else int(hi),
                                      RANGE_PREFIX_UNKNOWN if prefix_unknown else 0, C.c_void_p(cuda_stream))
        if not h:
            _raise_last()
        return PendingScan(self, h, fid, keep)


class PendingScan:
    """Handle of an asynchronous scan (sx_pending)."""

    def __init__(self, state, h, fid, keep):
        self._state, self._h, self._fid, self._keep = state, h, fid, keep

    def ready(self) -> bool:
        return self._h is None or bool(load_library().sx_pending_ready(self._h))

    def wait(self, raw: bool = False):
        h, self._h = self._h, None
        fc = load_library().sx_fc_wait(h)
        self._keep = None
        if raw:
            if not fc:
                _raise_last()
            return RawCollection(fc)
        return self._state._collect(fc, self._fid)


class RawCollection:
    """Thin owner of a C sx_finding_collection (for large result sets / benchmarks)."""

    def __init__(self, h):
        self._h = h

    def __len__(self):
        return load_library().sx_fc_len(self._h)

    def get(self, i: int):
        f = load_library().sx_fc_get(self._h, i).contents
        return (f.position, f.precision, C.string_at(f.s, f.s_len), bool(f.completes_previous))

    def all(self):
        L = load_library()
        n = L.sx_fc_len(self._h)
        arr = L.sx_fc_data(self._h)
        return [(arr[i].position, arr[i].precision, C.string_at(arr[i].s, arr[i].s_len), bool(arr[i].completes_previous))
                for i in range(n)]

    def as_numpy(self):
        """The findings as a structured numpy array over the collection's own memory (valid until close())."""
        import numpy as np

        dt = np.dtype({"names": ["position", "precision", "completes", "file_id", "mission_id", "s", "s_len", "in_start", "in_len"],
                       "formats": ["<u8", "u1", "u1", "<i2", "u1", "<u8", "<u4", "<i8", "<u4"],
                       "offsets": [0, 8, 9, 10, 12, 16, 24, 32, 40], "itemsize": 48})
        L = load_library()
        n = L.sx_fc_len(self._h)
        if n == 0:
            return np.zeros(0, dtype=dt)
        addr = C.cast(L.sx_fc_data(self._h), C.c_void_p).value
        buf = (C.c_uint8 * (48 * n)).from_address(addr)
        return np.frombuffer(buf, dtype=dt, count=n)

    def select(self, lo: int, hi: int):
        """(position, precision, text, completes) of the findings with lo <= position < hi (positions are monotone)."""
        import numpy as np

        a = self.as_numpy()
        i0, i1 = np.searchsorted(a["position"], [lo, hi], side="left")
        return [(int(r["position"]), int(r["precision"]), C.string_at(int(r["s"]), int(r["s_len"])), bool(r["completes"]))
                for r in a[i0:i1]]

    def close(self):
        if self._h:
            load_library().sx_fc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge(collections: Sequence[FindingCollection]) -> List[Finding]:
    """itertools::kmerge over one slice batch (main.rs:133) in `impl PartialOrd for Finding` order."""
    allf = [f for fc in collections for f in fc.v]
    allf.sort(key=lambda f: (f.position, f.mission.mission_id))  # stable: keeps each mission's emission order
    return allf
