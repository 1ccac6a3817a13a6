"""TEST-ONLY: drive tests/emul/sx_emul.cpp (the product's window automaton compiled for the CPU).

Mirrors what the product's host code (stringsext_b200/csrc/sx_api.cu) does around the kernels:
parameter block, virtual prefix / carry hand-off between calls, record -> finding conversion.
Used by the CPU test-suite to validate the parallel decomposition against the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsx_emul.so")


class Carry(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("flags", C.c_uint8), ("k", C.c_uint16), ("in_bytes", C.c_uint32),
                ("out_bytes", C.c_uint32), ("aux", C.c_uint32)]


class ScanParams(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("len", C.c_int64),
        ("slice_len", C.c_uint32), ("W", C.c_uint32), ("q", C.c_uint32), ("n", C.c_uint32),
        ("enc", C.c_uint32), ("align", C.c_uint32),
        ("af_lo", C.c_uint64), ("af_hi", C.c_uint64), ("ubf", C.c_uint64),
        ("base_consumed", C.c_uint64),
        ("npend", C.c_int32), ("is_last", C.c_int32),
        ("pend", C.c_uint8 * 8), ("carry_text8", C.c_uint8 * 8),
        ("carry_text_len", C.c_uint32),
        ("k0", Carry),
        ("grep_char", C.c_int32), ("same_block", C.c_uint32), ("general", C.c_uint32),
        ("sb_table", C.c_uint16 * 128),
        ("mb_a", C.c_void_p), ("mb_b", C.c_void_p), ("mb_fail", C.c_void_p),  # set by the harness itself
    ]


class Record(C.Structure):
    _fields_ = [("position", C.c_uint64), ("in_start", C.c_int64), ("text_off", C.c_uint64), ("in_len", C.c_uint32),
                ("text_len", C.c_uint32), ("flags", C.c_uint32), ("precision", C.c_uint32)]


class EmulOut(C.Structure):
    _fields_ = [("recs", C.POINTER(Record)), ("nrecs", C.c_size_t), ("text", C.POINTER(C.c_uint8)),
                ("ntext", C.c_size_t), ("final_carry", Carry), ("final_npend", C.c_int32), ("stats", C.c_uint64 * 8),
                ("list", C.POINTER(C.c_uint32)), ("nlist", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "sx_emul.cpp")
        csrc = os.path.join(HERE, "..", "..", "stringsext_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("sx_core.cuh", "sx_fast_utf8.cuh", "sx_fast_generic.cuh", "sx_mask_utf8.cuh", "sx_mb_tables.inc")]
        if (not os.path.exists(LIB)) or max(os.path.getmtime(f) for f in deps) > os.path.getmtime(LIB):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", LIB, src])
        L = C.CDLL(LIB)
        L.sx_emul_scan.argtypes = [C.POINTER(ScanParams), C.c_int, C.POINTER(EmulOut)]
        L.sx_emul_sizeof_params.restype = C.c_size_t
        assert L.sx_emul_sizeof_params() == C.sizeof(ScanParams)
        _lib = L
    return _lib


def last_multibyte_lead(s: bytes) -> int:
    """Lead byte of the last multi-byte char of a UTF-8 string, 0 if none (helper.rs:221, same-unicode-block rule)."""
    for b in reversed(s):
        if b >= 0xC0:
            return b
    return 0


def char_count(s: bytes) -> int:
    return sum(1 for b in s if (b & 0xC0) != 0x80)


class EmulState:
    """ScannerState for the emulated product path."""

    def __init__(self, m, use_pref=True):
        self.m = m
        self.use_pref = use_pref
        self.last_list = []
        self.windows_total = 0
        self.windows_listed = 0
        self.consumed = m.counter_offset
        self.leftover = b""
        self.cut = False
        self.pend = b""  # raw bytes still inside the decoder
        self.stream_off = 0  # total bytes seen (unit alignment for UTF-16/32)
        self.stats = [0] * 8

    def scan_stream(self, buf: bytes, is_last=False, slice_len=4096):
        m = self.m
        first = self.consumed
        if len(buf) == 0:
            return [], first
        P = ScanParams()
        cbuf = C.create_string_buffer(bytes(buf), len(buf))
        P.inp = C.cast(cbuf, C.c_void_p)
        P.len = len(buf)
        P.slice_len = slice_len
        P.q = m.output_line_char_nb_max
        P.W = 2 * P.q
        P.n = m.chars_min_nb
        P.enc = m.encoding_id
        unit = 2 if m.encoding_id in (2, 3) else 4 if m.encoding_id in (5, 6) else 1
        P.align = (-len(self.pend)) % unit if unit > 1 else 0
        P.af_lo = m.filter.af & 0xFFFFFFFFFFFFFFFF
        P.af_hi = (m.filter.af >> 64) & 0xFFFFFFFFFFFFFFFF
        P.ubf = m.filter.ubf
        P.base_consumed = self.consumed
        gc = m.filter.grep_char
        P.grep_char = -1 if gc is None else gc
        P.same_block = 1 if m.require_same_unicode_block else 0
        P.general = 1 if (gc is not None or m.require_same_unicode_block or m.chars_min_nb > m.output_line_char_nb_max) else 0
        P.npend = len(self.pend)
        P.is_last = 1 if is_last else 0
        for i, b in enumerate(self.pend):
            P.pend[8 - len(self.pend) + i] = b
        for i, b in enumerate(self.leftover[:8]):
            P.carry_text8[i] = b
        P.carry_text_len = len(self.leftover)
        if self.cut:
            P.k0 = Carry(1, 0, 0, 0, 0, 0)
        elif self.leftover:
            P.k0 = Carry(0, 1 | (2 if (gc is not None and gc in self.leftover) else 0), char_count(self.leftover), len(self.pend), 0,
                         last_multibyte_lead(self.leftover) if m.require_same_unicode_block else 0)
        else:
            P.k0 = Carry(0, 0, 0, 0, 0, 0)
        if m.sb_table is not None:
            for i, v in enumerate(m.sb_table):
                P.sb_table[i] = v
        out = EmulOut()
        rc = lib().sx_emul_scan(C.byref(P), 1 if self.use_pref else 0, C.byref(out))
        assert rc == 0
        text = C.string_at(out.text, out.ntext) if out.ntext else b""
        findings = []
        new_left = b""
        for i in range(out.nrecs):
            r = out.recs[i]
            t = text[r.text_off : r.text_off + r.text_len]
            if r.flags & 2:
                t = self.leftover + t
            if r.flags & 4:
                new_left = t
                continue
            findings.append((r.position, r.precision, t, bool(r.flags & 1)))
        for i in range(7):
            self.stats[i] += out.stats[i]
        self.stats[7] = out.stats[7]
        self.last_list = [out.list[i] for i in range(out.nlist)] if out.nlist < 200000 else None
        self.windows_listed += out.nlist
        fc = out.final_carry
        self.cut = fc.kind == 1
        self.leftover = new_left if fc.kind == 0 and fc.k > 0 else b""
        allb = self.pend + bytes(buf)
        self.pend = allb[len(allb) - out.final_npend :] if out.final_npend else b""
        self.consumed += len(buf)
        lib().sx_emul_free(C.byref(out))
        return findings, first
