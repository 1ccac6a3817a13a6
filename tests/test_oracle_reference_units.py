"""The CPU oracle against every hot-path unit test the reference ships.

CPU-only: this pins the oracle; the CUDA path is compared with the oracle in test_gpu_*.py.
"""
import pytest

from helpers import M, O, oracle_state, to_oracle
import reference_vectors as RV


@pytest.mark.parametrize("name,mission_factory,calls", RV.SCENARIOS, ids=[s[0].split(" ")[0] + str(i) for i, s in enumerate(RV.SCENARIOS)])
def test_scanner_scenarios(name, mission_factory, calls):
    ss = oracle_state(mission_factory())
    for c in calls:
        fc = ss.scan(c["inp"], c["last"], 0)
        got = [(f.position, f.precision, f.s) for f in fc.v]
        if c.get("findings") is not None:
            assert got == c["findings"], name
        if c.get("first_finding") is not None:
            assert got[0] == c["first_finding"], name
        if c.get("first") is not None:
            assert fc.first_byte_position == c["first"]
        assert not fc.str_buf_overflow
        if c.get("consumed") is not None:
            assert ss.consumed_bytes == c["consumed"]
        if c.get("cut") is not None:
            assert ss.cut == c["cut"]
        if c.get("leftover") is not None:
            assert ss.leftover == c["leftover"]


def test_field_with_zeros():
    factory, inp = RV.FIELD_WITH_ZEROS
    fc = oracle_state(factory()).scan(inp, False, 0)
    assert len(fc.v) != 1  # scanner.rs:557


def test_merger():
    """main.rs:234-305."""
    ms = RV.merger_missions()
    res = []
    for m in ms:
        res.append(oracle_state(m).scan(RV.MERGER_INPUT, True, 0).v)
    assert [f.s for f in res[0]] == [b"abcdefg", b"hijklmn", b"qrstuvw"]
    assert [f.s for f in res[1]] == ["abcdefgÜhijklmn".encode(), "opÜqrstuvwÜxyz".encode()]
    merged = sorted(res[0] + res[1], key=lambda f: (f.position, f.mission_id))
    assert [(f.s, f.position, f.precision, f.mission_id) for f in merged] == RV.MERGER_EXPECTED


# ---- helper.rs:479-641 test_split_s ---------------------------------------------------------
LATIN = (M.AF_ALL, M.UBF_LATIN, None)


def split(b, n, same, cut, inv, f, q):
    s = b.encode()
    return O.split_str(s, n, same, cut, inv, f[0], f[1], f[2], len(s) if q is None else q)


def test_split_s():
    r = split("€abc€defg€hijk€lm€opq", 3, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"abc", b"defg", b"hijk", b"opq"]
    assert not r[0]["completes"]

    r = split("ab€€defg€hijk€lm€opq", 3, False, True, False, LATIN, None)
    assert [x["s"] for x in r] == [b"ab", b"defg", b"hijk", b"opq"]
    assert r[0]["completes"] and not r[0]["min_ok"] and not r[0]["again"]
    assert r[3]["maybe_cut"] and r[3]["min_ok"] and r[3]["again"]

    r = split("ab€€defg€hijk€lm€op", 3, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"defg", b"hijk", b"op"]
    assert not r[0]["completes"]
    assert r[2]["maybe_cut"] and not r[2]["min_ok"] and r[2]["again"]

    r = split("€abc€defg€hijk€lm", 4, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"defg", b"hijk", b"lm"]
    assert not r[1]["maybe_cut"]
    assert r[2]["maybe_cut"] and not r[2]["min_ok"] and r[2]["again"]

    r = split("€abc€defg€hijk€lmno€", 4, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"defg", b"hijk", b"lmno"]
    assert not r[2]["maybe_cut"] and r[2]["min_ok"] and not r[2]["again"]

    r = split("abc€defghiÜjklmnpqrs€", 4, False, False, False, LATIN, 7)
    assert [x["s"] for x in r] == ["defghiÜ".encode(), b"jklmnpq", b"rs"]
    assert (r[0]["completes"], r[0]["maybe_cut"], r[0]["again"], r[0]["min_ok"]) == (False, True, False, True)
    assert (r[1]["completes"], r[1]["maybe_cut"], r[1]["again"], r[1]["min_ok"]) == (True, True, False, True)
    assert (r[2]["completes"], r[2]["maybe_cut"], r[2]["again"], r[2]["min_ok"]) == (True, False, False, False)

    r = split("abcdefghijklm", 4, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"abcdefghijklm"]
    assert (r[0]["completes"], r[0]["maybe_cut"], r[0]["again"], r[0]["min_ok"]) == (False, True, False, True)

    r = split("abcdefghijklm€", 4, False, False, False, LATIN, None)
    assert [x["s"] for x in r] == [b"abcdefghijklm"]
    assert (r[0]["completes"], r[0]["maybe_cut"], r[0]["again"], r[0]["min_ok"]) == (False, False, False, True)

    r = split("öö€€ääää€üü€éééé€", 4, False, True, False, LATIN, None)
    assert [x["s"] for x in r] == ["öö".encode(), "ääää".encode(), "éééé".encode()]

    r = split("öö€€ääää€üü€éééé€", 4, False, True, False, (M.AF_ALL, M.UBF_NONE, None), None)
    assert r == []


def test_split_s_require_same_unicode_block():
    """helper.rs:644-677."""
    f = (M.AF_ALL, M.UBF_LATIN | M.UBF_GREEK, None)
    r = split("0α1βγöäü€α2βγöäüöαβγαg34αäβüäöüαβγöäü", 3, False, False, False, f, None)
    assert [x["s"].decode() for x in r] == ["0α1βγöäü", "α2βγöäüöαβγαg34αäβüäöüαβγöäü"]
    r = split("0α1βγöäü€α2βγöäüöαβγαg34αäβüäöü", 4, True, False, False, f, None)
    assert [x["s"].decode() for x in r] == ["0α1βγ", "α2βγ", "öäüö", "αβγαg34α", "üäöü"]


def test_split_s_grep_char():
    """helper.rs:680-809."""
    b = "ac€€xefg€xijk€xm€xp"
    r = split(b, 3, False, True, False, LATIN, None)
    assert [x["s"] for x in r] == [b"ac", b"xefg", b"xijk", b"xp"]
    assert (r[0]["completes"], r[0]["again"], r[0]["maybe_cut"]) == (True, False, False)
    assert (r[3]["completes"], r[3]["again"], r[3]["maybe_cut"]) == (False, True, True)

    r = O.split_str(b.encode(), 2, False, True, False, M.AF_ALL, M.UBF_LATIN, ord("b"), 3)
    assert [x["s"] for x in r] == [b"ac"]
    assert (r[0]["completes"], r[0]["again"], r[0]["maybe_cut"]) == (True, False, False)

    r = O.split_str(b.encode(), 2, False, True, False, M.AF_ALL, M.UBF_LATIN, ord("x"), 3)
    assert [x["s"] for x in r] == [b"ac", b"xef", b"g", b"xij", b"k", b"xm", b"xp"]
    flags = [(x["completes"], x["again"], x["maybe_cut"], x["grep_ok"]) for x in r]
    assert flags == [
        (True, False, False, False),
        (False, False, True, True),
        (True, False, False, False),
        (False, False, True, True),
        (True, False, False, False),
        (False, False, False, True),
        (False, True, True, True),
    ]

    s = "öä€€äüöä€äüöö€üö€üü".encode()
    r = O.split_str(s, 3, False, False, False, M.AF_ALL, M.UBF_LATIN, ord("y"), len(s))
    assert [x["s"] for x in r] == ["üü".encode()]
    assert (r[0]["completes"], r[0]["again"], r[0]["maybe_cut"]) == (False, True, True)


def test_char_count_and_multibyte():
    """helper.rs:811-831."""
    assert O.char_count(b"hello") == 5
    assert O.char_count("abcö".encode()) == 4
    assert O.char_count("abc\U0010FFFFdef".encode()) == 7


def test_pass_filter():
    """mission.rs:758-774."""
    f = M.Utf8Filter(M.AF_ALL, M.UBF_LATIN, None)
    assert f.pass_af_filter("A".encode()[0])
    assert not f.pass_ubf_filter("€".encode()[0])
    assert f.pass_ubf_filter("©".encode()[0])
