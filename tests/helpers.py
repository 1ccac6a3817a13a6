"""Shared test helpers: product Mission -> oracle mission, synthetic corpora, comparisons."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (tests are allowed to use the oracle)
from stringsext_b200 import mission as M  # noqa: E402


def to_oracle(m: M.Mission) -> O.OMission:
    return O.OMission(
        encoding_id=m.encoding_id,
        chars_min_nb=m.chars_min_nb,
        af=m.filter.af,
        ubf=m.filter.ubf,
        grep_char=m.filter.grep_char,
        output_line_char_nb_max=m.output_line_char_nb_max,
        require_same_unicode_block=m.require_same_unicode_block,
        counter_offset=m.counter_offset,
        mission_id=m.mission_id,
        print_encoding_as_ascii=m.print_encoding_as_ascii,
        sb_table=m.sb_table,
        encoding_name=m.printed_encoding_name,
    )


def oracle_state(m: M.Mission) -> O.OState:
    return O.OState(to_oracle(m))


# ---- the reference's test fixtures (scanner.rs:105-191) --------------------------------
def mission_all_utf8():
    return M.Mission(M.ENC_UTF_8, "UTF-8", 3, False, M.UTF8_FILTER_ALL_VALID, 10, 10_000)


def mission_latin_utf8():
    return M.Mission(M.ENC_UTF_8, "UTF-8", 3, False, M.UTF8_FILTER_LATIN, 10, 10_000)


def mission_latin_utf8_grep42():
    f = M.Utf8Filter(M.AF_ALL & ~M.AF_CTRL | M.AF_WHITESPACE, M.UBF_LATIN, 42)
    return M.Mission(M.ENC_UTF_8, "UTF-8", 3, False, f, 10, 10_000)


def mission_all_x_user_defined():
    return M.Mission(M.ENC_X_USER_DEFINED, "x-user-defined", 3, False, M.UTF8_FILTER_ALL_VALID, 10, 10_000)


def mission_ascii():
    f = M.Utf8Filter(M.AF_ALL & ~M.AF_CTRL | M.AF_WHITESPACE, M.UBF_NONE, None)
    return M.Mission(M.ENC_X_USER_DEFINED, "x-user-defined", 3, False, f, 10, 10_000)


def mission_real_data_scan():
    return M.Mission(M.ENC_UTF_8, "UTF-8", 4, False, M.UTF8_FILTER_LATIN, 60, 10_000)
