"""Missions.new / parse_enc_opt (stringsext_b200/mission.py) against the reference's own parser test
(/root/reference/src/mission.rs:776-863, transcribed) and against the missions main.rs:198-231 builds for its merger
test; the defaulting and error rules of mission.rs:514-703.  CPU only."""
import pytest

import reference_vectors as RV
from helpers import M, oracle_state

P = M.Missions.parse_enc_opt


def test_enc_opt_parser_reference_vectors():
    # mission.rs:777-862, in the reference's order
    assert P("ascii") == ("ascii", None, None, None, None)
    assert P("utf-8,10,0x89AB,0xCDEF,0x2f") == ("utf-8", 10, 0x89AB, 0xCDEF, 0x2F)
    assert P("utf-8,10,0x89AB,0xCDEF,211") == ("utf-8", 10, 0x89AB, 0xCDEF, 211)
    assert P(",,,,,") == (None, None, None, None, None)
    assert P("ascii,10,0x89AB") == ("ascii", 10, 0x89AB, None, None)
    for bad in ("ascii, 10n", "ascii,10,0x89,0x?B", "ascii,10,0x?9,0xAB", "ascii,1000000000000000000000,0x1,0x2",
                "ascii,10,0x1,0x2,0x3,0x4", "ascii,10,123", "ascii,10,,123"):
        with pytest.raises(M.MissionError):
            P(bad)
    assert P("ascii,10,Default") == ("ascii", 10, M.AF_DEFAULT, None, None)
    assert P("ascii,10,,Latin") == ("ascii", 10, None, M.UBF_LATIN | M.UBF_ACCENTS, None)
    for bad in ("ascii,10,my-no-encoding", "ascii,10,,my-no-encoding"):
        with pytest.raises(M.MissionError):
            P(bad)


def test_filter_aliases_match_by_prefix_in_list_order():
    # parse_filter_parameter! (mission.rs:481-487) compares the string with the padded alias name's prefix: the first
    # alias in the list that starts with it wins
    assert P("utf-8,,,Cyr")[3] == M.UBF_CYRILLIC
    assert P("utf-8,,,A")[3] == M.UBF_AFRICAN
    assert P("utf-8,,,All")[3] == M.UBF_ALL & ~M.UBF_INVALID & ~M.UBF_ASIAN  # "All-Asian" stands before "All"
    assert P("utf-8,,All")[2] == M.AF_ALL
    assert P("utf-8,,All-Ctrl+Wsp")[2] == M.AF_ALL & ~M.AF_CTRL | M.AF_WHITESPACE
    assert P("utf-8,,W")[2] == M.AF_WHITESPACE
    assert P("utf-8,, None ,None")[2:4] == (M.AF_NONE, M.UBF_NONE)
    with pytest.raises(M.MissionError):
        P("utf-8,,,Cyrillic-but-longer")
    assert P("utf-8,+7")[1] == 7 and P("utf-8, 0x10 ")[1] == 16
    for bad in ("utf-8,256", "utf-8,-1", "utf-8,1_0", "utf-8,,,0x1FFFFFFFFFFFFFFFF"):
        with pytest.raises(M.MissionError):
            P(bad)


def test_missions_new_builds_the_merger_tests_missions():
    # main.rs:198-231: -e ascii -e utf-8 -n 5 --same-unicode-block -q 30 -s 5000
    ms = M.Missions.new("5000", ["ascii", "utf-8"], "5", True, None, None, None, "30")
    assert ms.v == RV.merger_missions() and len(ms) == 2
    got = []
    for m in ms.v:
        fc = oracle_state(m).scan(RV.MERGER_INPUT, True, 0)
        got += [(f.s, f.position, f.precision, m.mission_id) for f in fc.v]
    assert got == [(s, p, int(pr), mid) for s, p, pr, mid in RV.MERGER_EXPECTED]


def test_missions_new_defaults_and_precedence():
    ms = M.Missions.new()  # no -e: one UTF-8 mission with the defaults (options.rs:17-33, mission.rs:46-50)
    assert ms.v == [M.Mission.for_label("utf-8")]
    assert ms.v[0].chars_min_nb == 4 and ms.v[0].output_line_char_nb_max == 64 and ms.v[0].filter == M.UTF8_FILTER_NON_ASCII_MODE_DEFAULT
    ms = M.Missions.new(None, ["ascii", "utf-8,10,,Latin", "koi8-r,,All-Ctrl+Wsp,Cyr,0x3a", ",,,,"], "6", False, "None", "Common", "0x20", "0x20")
    a, b, c, d = ms.v
    assert (a.encoding_name, a.print_encoding_as_ascii, a.printed_encoding_name) == ("x-user-defined", True, "ascii")
    # the global flags win over the ASCII-mode defaults, item values over the global flags
    assert (a.chars_min_nb, a.filter.af, a.filter.ubf, a.filter.grep_char) == (6, M.AF_NONE, M.UBF_COMMON, 0x20)
    assert (b.chars_min_nb, b.filter.af, b.filter.ubf, b.filter.grep_char) == (10, M.AF_NONE, M.UBF_LATIN | M.UBF_ACCENTS, 0x20)
    assert (c.chars_min_nb, c.filter.af, c.filter.ubf, c.filter.grep_char) == (6, M.AF_ALL & ~M.AF_CTRL | M.AF_WHITESPACE, M.UBF_CYRILLIC, 0x3A)
    assert (d.encoding_name, d.mission_id, d.output_line_char_nb_max, c.sb_table is not None) == ("UTF-8", 3, 32, True)
    # without flags `ascii` gets the ASCII-mode filter (mission.rs:32-36, :623-641)
    assert M.Missions.new(None, ["ascii"]).v[0].filter == M.UTF8_FILTER_ASCII_MODE_DEFAULT


@pytest.mark.parametrize("kwargs", [
    dict(flag_grep_char="128"), dict(flag_encoding=["utf-8,,,,200"]), dict(flag_output_line_len="5"), dict(flag_encoding=["no-such-encoding"]),
    dict(flag_encoding=["utf-8,1,2,3,4,5"]), dict(flag_chars_min_nb="300"), dict(flag_counter_offset="0x1g"), dict(flag_ascii_filter="Nope"),
])
def test_missions_new_errors(kwargs):
    with pytest.raises(M.MissionError):
        M.Missions.new(**kwargs)
