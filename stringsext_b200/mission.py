"""Mission / Utf8Filter: host-side mirror of the reference's scanner parameters.

Mirrors (names, meaning, defaults) of
  * `Utf8Filter`  /root/reference/src/mission.rs:308-349
  * `Mission`     /root/reference/src/mission.rs:382-421
  * filter constants `AF_*` mission.rs:225-253, `UBF_*` mission.rs:72-161
  * default filters mission.rs:32-50 and the `ascii` emulation mission.rs:623-679
  * option defaults /root/reference/src/options.rs:17-33

`Missions.new` (mission.rs:514-749) turns the reference's option strings -- `-e ENC,MIN,AF,UBF,GREP` items, filter
aliases, hexadecimal filters -- into resolved `Mission`s with the reference's defaulting and error rules, so that a
caller can hand over the flags it gave the reference; option parsing proper (clap, options.rs) stays out of scope.
"""
from __future__ import annotations

import re

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

# ---- ASCII filters (u128 bitmaps indexed by ASCII code), mission.rs:225-253
AF_ALL = 0xFFFF_FFFF_FFFF_FFFF_FFFF_FFFF_FFFF_FFFE
AF_NONE = 0
AF_CTRL = 0x8000_0000_0000_0000_0000_0000_FFFF_FFFF
AF_WHITESPACE = 0x0000_0000_0000_0000_0000_0001_0000_1E00
AF_DEFAULT = AF_ALL & ~AF_CTRL

# ---- Unicode block filters (u64 bitmaps indexed by utf8_lead_byte & 0x3f), mission.rs:72-161
UBF_ALL = 0xFFFF_FFFF_FFFF_FFFF
UBF_NONE = 0
UBF_INVALID = 0xFFE0_0000_0000_0003
UBF_ALL_VALID = UBF_ALL & ~UBF_INVALID
UBF_LATIN = 0x0000_0000_0000_01FC
UBF_ACCENTS = 0x0000_0000_0000_3000
UBF_GREEK = 0x0000_0000_0000_C000
UBF_IPA = 0x0000_0000_0000_0700
UBF_CYRILLIC = 0x0000_0000_001F_0000
UBF_ARMENIAN = 0x0000_0000_0020_0000
UBF_HEBREW = 0x0000_0000_00C0_0000
UBF_ARABIC = 0x0000_0000_2F00_0000
UBF_SYRIAC = 0x0000_0000_1000_0000
UBF_AFRICAN = 0x0000_0000_FFE0_0000
UBF_COMMON = 0x0000_0000_FFFF_FFFC
UBF_KANA = 0x0000_0008_0000_0000
UBF_CJK = 0x0000_03F0_0000_0000
UBF_HANGUL = 0x0000_3800_0000_0000
UBF_ASIAN = 0x0000_3FFC_0000_0000
UBF_PUA = 0x0010_4000_0000_0000
UBF_MISC = 0x0000_8006_0000_0000
UBF_UNCOMMON = 0x000F_0000_0000_0000

# options.rs:17-33
ENCODING_DEFAULT = "UTF-8"
CHARS_MIN_DEFAULT = 4
COUNTER_OFFSET_DEFAULT = 0
OUTPUT_LINE_CHAR_NB_MAX_DEFAULT = 64
OUTPUT_LINE_CHAR_NB_MIN = 6
ASCII_ENC_LABEL = "ascii"

# Resolved encoding ids shared with include/stringsext_b200.h (SX_ENC_*).
ENC_X_USER_DEFINED = 0
ENC_UTF_8 = 1
ENC_UTF_16LE = 2
ENC_UTF_16BE = 3
ENC_SINGLE_BYTE = 4
ENC_UTF_32LE = 5  # extension: the reference has no UTF-32 (mission.rs:681-688)
ENC_UTF_32BE = 6  # extension
ENC_BIG5 = 7      # WHATWG Big5, index table generated from CPython's big5hkscs (tools/gen_multibyte_tables.py)
ENC_EUC_JP = 8    # WHATWG EUC-JP, jis0208 / jis0212 generated from CPython's euc_jp

# label -> (encoding id, canonical name as printed by Encoding::name(), single-byte table key)
_LABELS = {
    "ascii": (ENC_X_USER_DEFINED, "x-user-defined", None),
    "x-user-defined": (ENC_X_USER_DEFINED, "x-user-defined", None),
    "utf-8": (ENC_UTF_8, "UTF-8", None),
    "utf8": (ENC_UTF_8, "UTF-8", None),
    "utf-16le": (ENC_UTF_16LE, "UTF-16LE", None),
    "utf-16": (ENC_UTF_16LE, "UTF-16LE", None),
    "utf-16be": (ENC_UTF_16BE, "UTF-16BE", None),
    "utf-32le": (ENC_UTF_32LE, "UTF-32LE", None),
    "utf-32be": (ENC_UTF_32BE, "UTF-32BE", None),
    "koi8-r": (ENC_SINGLE_BYTE, "KOI8-R", "koi8-r"),
    "ibm866": (ENC_SINGLE_BYTE, "IBM866", "ibm866"),
    "iso-8859-5": (ENC_SINGLE_BYTE, "ISO-8859-5", "iso-8859-5"),
    "windows-1251": (ENC_SINGLE_BYTE, "windows-1251", "windows-1251"),
    "windows-1252": (ENC_SINGLE_BYTE, "windows-1252", "windows-1252"),
    "big5": (ENC_BIG5, "Big5", None),
    "big5-hkscs": (ENC_BIG5, "Big5", None),
    "euc-jp": (ENC_EUC_JP, "EUC-JP", None),
}


@dataclass(frozen=True)
class Utf8Filter:
    """mission.rs:308-327."""

    af: int = AF_DEFAULT
    ubf: int = UBF_COMMON
    grep_char: Optional[int] = None

    def pass_af_filter(self, b: int) -> bool:  # mission.rs:333-337
        assert b & 0x80 == 0
        return (1 << b) & self.af != 0

    def pass_ubf_filter(self, b: int) -> bool:  # mission.rs:341-348
        assert b & 0x80 == 0x80
        return (1 << (b & 0x3F)) & self.ubf != 0


UTF8_FILTER_ASCII_MODE_DEFAULT = Utf8Filter(AF_ALL & ~AF_CTRL, UBF_NONE, None)  # mission.rs:32-36
UTF8_FILTER_NON_ASCII_MODE_DEFAULT = Utf8Filter(AF_ALL & ~AF_CTRL, UBF_COMMON, None)  # mission.rs:46-50
UTF8_FILTER_ALL_VALID = Utf8Filter(AF_ALL, UBF_ALL & ~UBF_INVALID, None)  # mission.rs:55-59 (cfg(test))
UTF8_FILTER_LATIN = Utf8Filter(AF_ALL & ~AF_CTRL | AF_WHITESPACE, UBF_LATIN | UBF_ACCENTS, None)  # :64-68


@dataclass(frozen=True)
class Mission:
    """mission.rs:382-421 with `encoding` resolved to an id (+ table for single-byte encodings)."""

    encoding_id: int = ENC_UTF_8
    encoding_name: str = "UTF-8"
    chars_min_nb: int = CHARS_MIN_DEFAULT
    require_same_unicode_block: bool = False
    filter: Utf8Filter = field(default_factory=lambda: UTF8_FILTER_NON_ASCII_MODE_DEFAULT)
    output_line_char_nb_max: int = OUTPUT_LINE_CHAR_NB_MAX_DEFAULT
    counter_offset: int = COUNTER_OFFSET_DEFAULT
    mission_id: int = 0
    print_encoding_as_ascii: bool = False
    sb_table: Optional[Tuple[int, ...]] = None  # 128 code points for bytes 0x80..0xFF, 0 = unmapped

    @staticmethod
    def for_label(
        label: str,
        chars_min_nb: Optional[int] = None,
        af: Optional[int] = None,
        ubf: Optional[int] = None,
        grep_char: Optional[int] = None,
        output_line_char_nb_max: Optional[int] = None,
        require_same_unicode_block: bool = False,
        counter_offset: int = COUNTER_OFFSET_DEFAULT,
        mission_id: int = 0,
    ) -> "Mission":
        """The defaulting rules of Missions::new (mission.rs:583-699) for one resolved `-e` item."""
        key = label.strip().lower()
        if key not in _LABELS:
            raise ValueError(f"invalid input encoding name `{label}`")  # mission.rs:681-688
        enc_id, name, table_key = _LABELS[key]
        is_ascii = key == ASCII_ENC_LABEL
        dflt = UTF8_FILTER_ASCII_MODE_DEFAULT if is_ascii else UTF8_FILTER_NON_ASCII_MODE_DEFAULT
        if grep_char is not None and grep_char > 127:
            raise ValueError("you can only grep for ASCII codes < 128")  # mission.rs:657-667
        q = OUTPUT_LINE_CHAR_NB_MAX_DEFAULT if output_line_char_nb_max is None else output_line_char_nb_max
        if q < OUTPUT_LINE_CHAR_NB_MIN:
            raise ValueError(f"minimum for `--output-line-len` is `{OUTPUT_LINE_CHAR_NB_MIN}`")  # :612-621
        table = None
        if table_key is not None:
            from .sb_tables import SINGLE_BYTE_TABLES

            table = SINGLE_BYTE_TABLES[table_key]
        return Mission(
            encoding_id=enc_id,
            encoding_name=name,
            chars_min_nb=CHARS_MIN_DEFAULT if chars_min_nb is None else chars_min_nb,
            require_same_unicode_block=require_same_unicode_block,
            filter=Utf8Filter(
                dflt.af if af is None else af, dflt.ubf if ubf is None else ubf, grep_char
            ),
            output_line_char_nb_max=q,
            counter_offset=counter_offset,
            mission_id=mission_id,
            print_encoding_as_ascii=is_ascii,
            sb_table=table,
        )

    @property
    def printed_encoding_name(self) -> str:  # finding.rs:144-148
        return ASCII_ENC_LABEL if self.print_encoding_as_ascii else self.encoding_name


def known_labels() -> Sequence[str]:
    return sorted(_LABELS)


# ---- filter aliases, mission.rs:163-212 / :245-268 (name, value); matched by PREFIX in this order (parse_filter_parameter!)
UNICODE_BLOCK_FILTER_ALIASSE = (
    ("African", UBF_AFRICAN), ("All-Asian", UBF_ALL & ~UBF_INVALID & ~UBF_ASIAN), ("All", UBF_ALL & ~UBF_INVALID),
    ("Arabic", UBF_ARABIC | UBF_SYRIAC), ("Armenian", UBF_ARMENIAN), ("Asian", UBF_ASIAN), ("Cjk", UBF_CJK), ("Common", UBF_COMMON),
    ("Cyrillic", UBF_CYRILLIC), ("Default", UBF_ALL & ~UBF_INVALID), ("Greek", UBF_GREEK), ("Hangul", UBF_HANGUL),
    ("Hebrew", UBF_HEBREW), ("Kana", UBF_KANA), ("Latin", UBF_LATIN | UBF_ACCENTS), ("None", UBF_NONE), ("Private", UBF_PUA),
    ("Uncommon", UBF_UNCOMMON | UBF_PUA),
)
ASCII_FILTER_ALIASSE = (
    ("All", AF_ALL), ("All-Ctrl", AF_ALL & ~AF_CTRL), ("All-Ctrl+Wsp", AF_ALL & ~AF_CTRL | AF_WHITESPACE), ("Default", AF_DEFAULT),
    ("None", AF_NONE), ("Wsp", AF_WHITESPACE),
)
_ALIAS_WIDTH = 12  # the reference pads every alias name to 12 bytes and compares the given string with that prefix


class MissionError(ValueError):
    """What Missions::new reports as an anyhow error."""


def _rust_uint(text: str, radix: int, bits: int, what: str) -> int:
    # {integer}::from_str / from_str_radix: optional '+', digits of the radix, nothing else; overflow is an error
    digits = "0-9" if radix == 10 else "0-9a-fA-F"
    if not re.fullmatch(rf"\+?[{digits}]+", text):
        raise MissionError(f"failed to parse {what}: `{text}`")
    v = int(text, radix)
    if v >> bits:
        raise MissionError(f"failed to parse {what}: `{text}` (number too large)")
    return v


def parse_integer(s: Optional[str], bits: int) -> Optional[int]:
    """mission.rs:440-457 (parse_integer!): None / empty -> None, `0x..` -> hexadecimal, else decimal."""
    if s is None or s == "":
        return None
    t = s.strip()
    if len(t) >= 2 and t[:2] == "0x":
        return _rust_uint(t[2:], 16, bits, "hexadecimal number")
    return _rust_uint(t, 10, bits, "number")


def parse_filter_parameter(s: Optional[str], bits: int, aliases) -> Optional[int]:
    """mission.rs:468-500 (parse_filter_parameter!): `0x..` -> hexadecimal bitmap, empty -> None, else the first alias
    the (trimmed) string is a prefix of."""
    if s is None:
        return None
    t = s.strip()
    if len(t) >= 2 and t[:2] == "0x":
        return _rust_uint(t[2:], 16, bits, "hexadecimal number")
    if s == "":
        return None
    for name, value in aliases:
        if len(t) <= _ALIAS_WIDTH and name.ljust(_ALIAS_WIDTH).startswith(t):
            return value
    raise MissionError(f"filter name `{t}` is not valid, try `--list-encodings`")


class Missions:
    """mission.rs:424-749: the missions of one run, built from the reference's option strings."""

    def __init__(self, v: Sequence[Mission]):
        self.v = list(v)

    def __len__(self) -> int:
        return len(self.v)

    @staticmethod
    def parse_enc_opt(enc_opt: str):
        """mission.rs:706-748: `ENC,MIN,AF,UBF,GREP` -> (enc_name, chars_min_nb, af, ubf, grep_char), None where absent."""
        items = enc_opt.split(",")
        if items and items[-1] == "":
            items.pop()  # str::split_terminator
        it = iter(items)
        first = next(it, None)
        enc_name = None if first in (None, "") else first.strip()
        chars_min_nb = parse_integer(next(it, None), 8)
        af = parse_filter_parameter(next(it, None), 128, ASCII_FILTER_ALIASSE)
        ubf = parse_filter_parameter(next(it, None), 64, UNICODE_BLOCK_FILTER_ALIASSE)
        grep_char = parse_integer(next(it, None), 8)
        if next(it, None) is not None:
            raise MissionError(f"Too many items in `{enc_opt}`.")
        return enc_name, chars_min_nb, af, ubf, grep_char

    @staticmethod
    def new(flag_counter_offset: Optional[str] = None, flag_encoding: Sequence[str] = (), flag_chars_min_nb: Optional[str] = None,
            flag_same_unicode_block: bool = False, flag_ascii_filter: Optional[str] = None, flag_unicode_block_filter: Optional[str] = None,
            flag_grep_char: Optional[str] = None, flag_output_line_len: Optional[str] = None) -> "Missions":
        """mission.rs:514-703.  Item values win over the global flags, those over the defaults (ASCII mode for `ascii`)."""
        counter_offset = parse_integer(flag_counter_offset, 64)
        chars_min = parse_integer(flag_chars_min_nb, 8)
        g_af = parse_filter_parameter(flag_ascii_filter, 128, ASCII_FILTER_ALIASSE)
        g_ubf = parse_filter_parameter(flag_unicode_block_filter, 64, UNICODE_BLOCK_FILTER_ALIASSE)
        g_grep = parse_integer(flag_grep_char, 8)
        if g_grep is not None and g_grep > 127:
            raise MissionError(f"you can only `--grep-char` for ASCII codes < 128, you tried: `{g_grep}`.")
        out_len = parse_integer(flag_output_line_len, 64)
        if out_len is not None and out_len < OUTPUT_LINE_CHAR_NB_MIN:
            raise MissionError(f"minimum for `--output-line-len` is `{OUTPUT_LINE_CHAR_NB_MIN}`, you tried: `{out_len}`.")
        v = []
        for mission_id, enc_opt in enumerate(list(flag_encoding) or [ENCODING_DEFAULT]):
            enc_name, n, af, ubf, grep = Missions.parse_enc_opt(enc_opt)
            scanner = chr(97 + mission_id)
            grep = grep if grep is not None else g_grep
            if grep is not None and grep > 127:
                raise MissionError(f"Scanner {scanner}: you can only grep for ASCII codes < 128, you tried: `{grep}`.")
            try:
                v.append(Mission.for_label(
                    enc_name or ENCODING_DEFAULT,
                    n if n is not None else chars_min,
                    af if af is not None else g_af,
                    ubf if ubf is not None else g_ubf,
                    grep,
                    out_len,
                    flag_same_unicode_block,
                    COUNTER_OFFSET_DEFAULT if counter_offset is None else counter_offset,
                    mission_id,
                ))
            except MissionError:
                raise
            except ValueError as e:
                raise MissionError(f"Scanner {scanner}: {e}, try flag `--list-encodings`.") from None
        return Missions(v)
