"""Synthetic corpora shared by the CPU and GPU tests and by bench.py.

`sx_mix_bytes` reproduces on the host (numpy) exactly what `sx_fill_random` writes on the device:
byte i of the stream is byte (i & 7), little endian, of splitmix64(seed, i >> 3).
`plant` overwrites a sparse, seeded set of positions with strings in the mission's encoding so
that encodings which find nothing in random bytes (UTF-16/32) still have parity
to check; placements straddle window and slice boundaries on purpose.
"""
from __future__ import annotations

import random

import numpy as np

_M64 = (1 << 64) - 1


def sx_mix_bytes(seed: int, offset: int, length: int) -> np.ndarray:
    """uint8 array: bytes [offset, offset+length) of the stream with the given seed."""
    w0 = offset >> 3
    w1 = (offset + length + 7) >> 3
    idx = np.arange(w0, w1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = z.astype("<u8").view(np.uint8)
    s = offset - (w0 << 3)
    return b[s : s + length].copy()


_WORDS_MB = ["\u00ca\u0304", "\u00ea\u030cx", "中文", "ｱｲｳ", "Āā"]
_WORDS = ["hello", "world", "straße", "naïve", "€uro", "日本語", "😀", "a", "xy", "longer-word-here", "ΑΒΓ", "привет",
          "Հայերեն", "עברית", "العربية"]


def encode_for(enc_id: int, s: str) -> bytes:
    if enc_id == 2:
        return s.encode("utf-16le")
    if enc_id == 3:
        return s.encode("utf-16be")
    if enc_id == 5:
        return s.encode("utf-32le")
    if enc_id == 6:
        return s.encode("utf-32be")
    if enc_id == 4:
        return s.encode("koi8-r", errors="replace")
    if enc_id == 0:
        return s.encode("ascii", errors="replace")
    if enc_id == 7:
        return s.encode("big5hkscs", errors="replace")
    if enc_id == 8:
        return s.encode("euc_jp", errors="replace")
    return s.encode("utf-8")


def planted_strings(rng: random.Random, enc_id: int, n: int, q: int):
    """Strings of the lengths SURVEY.md 8(d) lists, in the mission's encoding."""
    out = []
    alphabets = ["abcdefghijklmnopqrstuvwxyz0123456789 /-_.", "äöüßéèàç", "€日本語", "😀🎉", "բարևՀայ", "שלום", "مرحبا"]
    for ln in (n - 1, n, n + 1, q - 1, q, q + 1, 2 * q - 1, 2 * q, 2 * q + 1, 5 * q):
        if ln <= 0:
            continue
        for _ in range(2):
            k = rng.random()
            alpha = alphabets[0] if k < 0.4 else alphabets[0] + rng.choice(alphabets[1:])
            if enc_id in (0,):
                alpha = alphabets[0]
            if enc_id == 4:
                alpha = alphabets[0] + "абвгдежзийклмноп"
            if enc_id == 7:  # Big5: CJK, Greek, Cyrillic, and the pairs that decode to two code points
                alpha = alphabets[0] + rng.choice(["日本語中文字", "ΑΒΓΔαβγδ", "абвгдеж", "\u00ca\u0304\u00ea\u030c", "äöü"])
            if enc_id == 8:  # EUC-JP: jis0208 kanji / kana / Greek / Cyrillic, half-width katakana (8E), jis0212 Latin (8F)
                alpha = alphabets[0] + rng.choice(["日本語漢字かなカナ", "ΑΒΓΔαβγδ", "абвгдеж", "ｱｲｳｴｵ", "ÀÁÂãäåĀāĂ"])
            s = "".join(rng.choice(alpha) for _ in range(ln))
            out.append(encode_for(enc_id, s))
    return out


def plant(buf: np.ndarray, seed: int, enc_id: int, n: int, q: int, slice_len: int = 4096, density: int = 1 << 16):
    """Overwrite ~len/density positions (at least a dozen) with planted strings, in place."""
    rng = random.Random(seed * 7919 + enc_id)
    ln = len(buf)
    strings = planted_strings(rng, enc_id, n, q)
    W = 2 * q
    count = max(12, ln // density)
    for i in range(count):
        s = strings[i % len(strings)]
        if len(s) + 8 >= ln:
            continue
        kind = i % 6
        base = rng.randrange(0, max(1, ln - len(s) - 8))
        if kind == 1:  # straddle a window boundary
            base = (base // W) * W + W - rng.randrange(1, max(2, min(len(s), W)))
        elif kind == 2:  # straddle a slice boundary
            base = (base // slice_len) * slice_len + slice_len - rng.randrange(1, max(2, min(len(s), slice_len)))
        elif kind == 3:  # odd offset
            base |= 1
        elif kind == 4:  # right after a malformed byte
            if base > 0:
                buf[base - 1] = 0xFF
        elif kind == 5:  # ends exactly at a malformed byte
            if base + len(s) < ln:
                buf[base + len(s)] = 0xC0
        base = max(0, min(base, ln - len(s)))
        buf[base : base + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return buf


def gen(rng: random.Random, kind: str, n: int, enc: int) -> bytes:
    """Small adversarial buffers for differential fuzzing."""
    if n <= 0:
        return b""
    if kind == "rand":
        return bytes(rng.getrandbits(8) for _ in range(n))
    if kind == "lowent" and enc in (7, 8) and rng.random() < 0.5:
        alpha = (b"ab \x00\xa4\x40\xa3\x44\x88\x62\x80" if enc == 7 else b"ab \x00\xc6\xfc\x8e\xb1\x8f\xaa\xa1\x80")
        return bytes(rng.choice(alpha) for _ in range(n))
    if kind == "lowent":
        alpha = rng.choice([b"ab\x00", b"abc \x00\xc3\xa9\xe2\x82\xac", b"a\x00", bytes(range(0x20, 0x7F)) + b"\x00\x01\xff",
                            b"\xc3\xa9\xc3a\x80", b"a\x00b\x00\xd8\x00\xdc\x3d\xd8", b"\xe2\x82\xac\xf0\x9f\x98\x80a\x00"])
        return bytes(rng.choice(alpha) for _ in range(n))
    if kind == "text":
        s = ""
        words = _WORDS + _WORDS_MB if enc in (7, 8) else _WORDS
        while len(s) < n:
            s += rng.choice(words) + rng.choice([" ", " ", "\n", "\x00", "", "\t"])
        b = encode_for(enc, s)
        off = rng.randrange(0, 4)
        return (bytes(rng.getrandbits(8) for _ in range(off)) + b)[:n]
    if kind == "runs":
        out = bytearray()
        choices = [b"a", b"\xc3\xa9", b"\xe2\x82\xac", b"\xf0\x9f\x98\x80", b"a\x00", b"\x00a", b"\x00", b"\xff",
                   b"\xc0", b" ", b"\xe9"]
        if enc == 7:  # Big5 pairs: CJK, Greek, a two-code-point pair, an unmapped pair with an ASCII trail, a lone lead
            choices = choices + [b"\xa4\x40", b"\xa3\x44", b"\x88\x62", b"\x88\xa5a", b"\x81\x41", b"\xa1", b"\xfe\xfe"]
        if enc == 8:  # EUC-JP: kanji, half-width katakana, jis0212, broken three-byte sequences
            choices = choices + [b"\xc6\xfc", b"\x8e\xb1", b"\x8f\xaa\xa1", b"\x8f\xa1", b"\x8e", b"\xa1\x41", b"\x8f\x8f\xb0\xa1"]
        while len(out) < n:
            c = rng.choice(choices)
            out += c * rng.randrange(1, 300)
        return bytes(out[:n])
    out = bytearray()  # mixed
    while len(out) < n:
        out += gen(rng, rng.choice(["rand", "lowent", "text", "runs"]), rng.randrange(1, 400), enc)
    return bytes(out[:n])


def gen_blocks(rng: random.Random, n: int, enc: int) -> bytes:
    """Short runs that mix unicode blocks, separated by failing ASCII junk: what --same-unicode-block splits on, and
    where SplitStr's stale lead byte (helper.rs:221, :327-330) decides how a run across a window boundary is cut."""
    toks = ["Я", "ж", "Γ", "Δ", "é", "ü", "€", "日", "a", "b", "abc", "xy", " ", "\x01", "\x02", "\x01\x01", "Яa", "aΓ", "abЯ", "abcΓΔ", "éa€",
            "\x7f", "0123", "ЯЯ", "ΓΓΓ"]
    out = bytearray()
    junk = rng.choice([0.1, 0.3, 0.5])
    while len(out) < n:
        t = rng.choice(["\x01", "\x02", "\x01\x01\x01"]) if rng.random() < junk else rng.choice(toks)
        if rng.random() < 0.03:
            out += rng.choice([b"\xff", b"\xc0", b"\x80", b"\xe2\x82", b"\xd8"])
        b = encode_for(enc, t) if enc not in (0, 4) else t.encode("koi8-r" if enc == 4 else "latin-1", errors="ignore")
        if enc in (0, 4) and rng.random() < 0.3:
            b += bytes([rng.choice([0xC0, 0xE9, 0xA3, 0xB3, 0x80, 0x9F, 0xFF])])
        out += b
    return bytes(out[:n])


KINDS = ["rand", "lowent", "text", "mixed", "runs"]
LABELS = {0: "ascii", 1: "utf-8", 2: "utf-16le", 3: "utf-16be", 4: "koi8-r", 5: "utf-32le", 6: "utf-32be", 7: "big5", 8: "euc-jp"}


def random_mission(rng: random.Random, enc: int, M):
    q = rng.choice([6, 7, 10, 16, 32, 64, 64, 100])
    n = min(rng.choice([1, 2, 3, 4, 4, 6, 10, q]), q)
    af = rng.choice([M.AF_DEFAULT, M.AF_ALL, M.AF_DEFAULT | M.AF_WHITESPACE])
    ubf = rng.choice([M.UBF_COMMON, M.UBF_NONE, M.UBF_ALL_VALID, M.UBF_LATIN | M.UBF_ACCENTS, M.UBF_AFRICAN, M.UBF_ALL])
    label = LABELS[enc]
    if enc == 4:
        label = rng.choice(["koi8-r", "windows-1251", "windows-1252", "ibm866", "iso-8859-5"])
    if enc == 0 and rng.random() < 0.5:
        ubf = M.UBF_NONE
    return M.Mission.for_label(label, n, af, ubf, None, q, counter_offset=rng.choice([0, 10000]))


def random_general_mission(rng: random.Random, enc: int, M):
    """Missions with --grep-char and / or --same-unicode-block."""
    import dataclasses

    m = random_mission(rng, enc, M)
    kind = rng.choice(["grep", "same", "both"])
    grep = rng.choice([0x20, ord("a"), ord("e"), ord(":"), ord("?"), 0x00]) if kind in ("grep", "both") else None
    filt = dataclasses.replace(m.filter, grep_char=grep)
    return dataclasses.replace(m, filter=filt, require_same_unicode_block=kind in ("same", "both"))


# A window killed by its predecessor's leftover (helper.rs:389-415), found by the general-mission fuzz: ISO-8859-5,
# -n 3, q = 8 (16-byte windows), --grep-char ':', every ASCII char but NUL passes, no multi-byte block does.
# Window 1 ends in the 8 = q chars `s2W\x02i}|k` without the grep char: kept as a leftover ("again").  Window 2 starts
# with a failing byte: the leftover is evaluated there, holds q chars and no grep char, and the whole segment is dropped
# -- including its trailing `3:`, which is therefore NOT a leftover, so `3:O` across the boundary to window 3 does not
# print.  With a ':' in window 1's run both strings print.  (This is why such missions cannot use the prefilter as it is:
# window 2 has no run of >= n good bytes, yet its carry-out depends on its carry-in.)
KILLED_WINDOW_CASE = (
    b"!n\xb0\xbe&\x8c\\\x8aR\x07QB\xc0\x12\xcb\xd7"
    b"\xf3\xb8\xba\x94=\xa1\xd3\xfas2W\x02i}|k"
    b"\xd1\xc9\xe4}/\xe2\xd0\xb6\xed\xe0\xe5\x8c\xb8\xa63:"
    b"O\xd6T\xfd\x02\xb4\xaa{\x8e\x87\xfe\n\xd3O\x02\xb1"
)


def killed_window_inputs():
    """(mission arguments for Mission.for_label, [(input, expected (position, text) list)])"""
    pad = lambda c: c + b"\x00" * (4096 + 64 - len(c))
    return (("iso-8859-5", 3, (1 << 128) - 2, 0, ord(":"), 8),
            [(pad(KILLED_WINDOW_CASE), []),
             (pad(KILLED_WINDOW_CASE.replace(b"s2W", b"s2:")), [(16, b"s2:\x02i}|k"), (48, b"3:O")])])
