"""Golden vectors transcribed from the reference's own hot-path unit tests.

Each scenario is a sequence of `FindingCollection::from` calls on one `ScannerState`
with the assertions the reference makes (None = the reference does not assert it).
Sources: /root/reference/src/scanner.rs:193-559, src/finding_collection.rs:431-502,
src/main.rs:234-305.
"""
from helpers import (
    M,
    mission_all_utf8,
    mission_all_x_user_defined,
    mission_ascii,
    mission_latin_utf8,
    mission_latin_utf8_grep42,
    mission_real_data_scan,
)

B, E, A = 0, 1, 2  # Precision::{Before, Exact, After}  finding.rs:34-46


def F(pos, prec, s):
    return (pos, prec, s.encode() if isinstance(s, str) else s)


# (name, mission factory, [call, ...]); call = dict(inp, last, findings|None, n|None, first, consumed, cut, leftover)
SCENARIOS = [
    (
        "test_scan_input_buffer_chunks (scanner.rs:193-221)",
        mission_all_utf8,
        [
            dict(
                inp=b"a234567890b234567890c234", last=True,
                findings=[F(10000, E, "a234567890"), F(10000, A, "b234567890"), F(10020, E, "c234")],
                first=10000, consumed=10024, cut=False, leftover=None,
            )
        ],
    ),
    (
        "test_scan_store_in_scanner_state (scanner.rs:224-255)",
        mission_all_utf8,
        [
            dict(
                inp=b"a234567890b234567890c2", last=True,
                findings=[F(10000, E, "a234567890"), F(10000, A, "b234567890"), F(10020, E, "c2")],
                first=10000, consumed=10022, cut=False, leftover=None,
            )
        ],
    ),
    (
        "test_split_str_iterator_and_store_in_scanner_state (scanner.rs:258-304)",
        mission_all_utf8,
        [
            dict(inp=b"You\xC0\x82\xC0co", last=False, findings=[F(10000, E, "You")],
                 first=10000, consumed=10008, cut=None, leftover=b"co"),
            dict(inp=b"me\xC0\x82\xC0home.", last=True,
                 findings=[F(10008, B, "come"), F(10013, E, "home.")],
                 first=10008, consumed=10018, cut=None, leftover=b""),
        ],
    ),
    (
        "test_grep_in_scan (scanner.rs:307-350)",
        mission_latin_utf8_grep42,
        [
            dict(inp=b"You\xC0\x82\xC0co", last=False, findings=[], first=10000, consumed=10008, cut=None,
                 leftover=b"co"),
            dict(inp=b"me*\xC0\x82\xC0ho*me.\x82", last=True,
                 findings=[F(10008, B, "come*"), F(10014, E, "ho*me.")],
                 first=10008, consumed=10021, cut=None, leftover=b""),
        ],
    ),
    (
        "test_scan_buffer_split_multibyte (scanner.rs:355-412)",
        mission_all_utf8,
        [
            dict(inp=b"word\xe2\x82", last=False, findings=None, first=None, consumed=None, cut=None, leftover=None),
            dict(inp=b"\xacoh\xC0no no", last=False, findings=None, first=10006, consumed=10015, cut=None,
                 leftover=None, first_finding=F(10006, B, "word€oh")),
            dict(inp=b"\xe2\x82\xacStream end.", last=True,
                 findings=[F(10015, B, "no no€Stre"), F(10015, A, "am end.")],
                 first=10015, consumed=10029, cut=None, leftover=None),
        ],
    ),
    (
        "test_to_short1 (scanner.rs:415-470)",
        mission_all_utf8,
        [
            dict(inp=b"ii\xC0abc\xC0\xC1de\xC0fgh\xC0ijk", last=False,
                 findings=[F(10003, E, "abc"), F(10011, E, "fgh")],
                 first=10000, consumed=10018, cut=False, leftover=b"ijk"),
            dict(inp=b"b\xC0\x82c\xC0def", last=True,
                 findings=[F(10018, B, "ijkb"), F(10023, E, "def")],
                 first=10018, consumed=10026, cut=False, leftover=b""),
        ],
    ),
    (
        "test_to_short2 (scanner.rs:473-531)",
        mission_latin_utf8,
        [
            dict(inp="ii€ääà€€de€fgh€ijk".encode(), last=False,
                 findings=[F(10000, E, "ääà"), F(10020, B, "fgh")],
                 first=10000, consumed=10031, cut=False, leftover=b"ijk"),
            dict(inp=b"b\xC0\x82c\xC0def", last=True,
                 findings=[F(10031, B, "ijkb"), F(10036, E, "def")],
                 first=10031, consumed=10039, cut=False, leftover=b""),
        ],
    ),
    (
        "test_ascii_emulation part 1 (finding_collection.rs:431-465)",
        mission_all_x_user_defined,
        [
            dict(inp=b"abcdefg\x58\x59\x80\x82h\x83ijk\x89\x90", last=True,
                 findings=[F(10000, E, "abcdefgXY"), F(10000, A, "hijk")],
                 first=10000, consumed=10018, cut=False, leftover=b""),
        ],
    ),
    (
        "test_ascii_emulation part 2 (finding_collection.rs:467-502)",
        mission_ascii,
        [
            dict(inp=b"abcdefg\x58\x59\x80\x82h\x83ijk\x89\x90", last=False,
                 findings=[F(10000, E, "abcdefgXY"), F(10000, A, "ijk")],
                 first=10000, consumed=10018, cut=False, leftover=b""),
        ],
    ),
]

# scanner.rs:534-559: regression; the reference only asserts len != 1
FIELD_WITH_ZEROS = (mission_real_data_scan, b"\x00\x00\x00\x00\x40\x00\x38\x00\x0c\x00\x40\x00\x2c\x00\x2b\x00")


def merger_missions():
    """main.rs:198-231: -e ascii -e utf-8 -n 5 -r -q 30 -s 5000."""
    return [
        M.Mission.for_label("ascii", 5, output_line_char_nb_max=30, require_same_unicode_block=True,
                            counter_offset=5000, mission_id=0),
        M.Mission.for_label("utf-8", 5, output_line_char_nb_max=30, require_same_unicode_block=True,
                            counter_offset=5000, mission_id=1),
    ]


MERGER_INPUT = "abcdefgÜhijklmn€opÜqrstuvwÜxyz".encode()
# main.rs:254-304: merged order (s, position, precision, mission_id)
MERGER_EXPECTED = [
    (b"abcdefg", 5000, E, 0),
    (b"hijklmn", 5000, A, 0),
    (b"qrstuvw", 5000, A, 0),
    ("abcdefgÜhijklmn".encode(), 5000, E, 1),
    ("opÜqrstuvwÜxyz".encode(), 5000, A, 1),
]
