"""The oracle's Big5 / EUC-JP decoders (WHATWG decoders over the library's generated index tables): table layout and
pointer arithmetic against CPython's codecs (the tables' source), malformed-sequence conventions of SURVEY.md App. A.4.
Parity for these two encodings is self-consistent (oracle == kernels on the same table), not reference-pinned."""
import random

import pytest

from helpers import M, O, oracle_state


@pytest.mark.parametrize("enc,codec,sample", [
    (O.ENC_BIG5, "big5hkscs", "hello 中文字 ΑΒΓ абв 日本語 end"),
    (O.ENC_EUC_JP, "euc_jp", "hello 日本語 かな ｱｲｳ ΑΒΓ абв Āā end"),
])
def test_valid_text_decodes_like_the_codec(enc, codec, sample):
    b = sample.encode(codec)
    res, rd, out = O.decode(enc, [b])[-1]
    assert out.decode("utf-8") == sample and rd == len(b)


def test_every_mapped_pair_round_trips():
    rng = random.Random(1)
    for enc, codec, leads, trails in ((O.ENC_BIG5, "big5hkscs", range(0x81, 0xFF), list(range(0x40, 0x7F)) + list(range(0xA1, 0xFF))),
                                      (O.ENC_EUC_JP, "euc_jp", range(0xA1, 0xFF), range(0xA1, 0xFF))):
        for _ in range(3000):
            pair = bytes([rng.choice(list(leads)), rng.choice(list(trails))])
            try:
                exp = pair.decode(codec)
            except UnicodeDecodeError:
                exp = None
            res, rd, out = O.decode(enc, [pair + b"!"])[0]
            if exp is not None:
                assert out.decode("utf-8") == exp + "!", pair.hex()
            else:  # unmapped: malformed; an ASCII trail is not consumed
                assert res == 2 and rd == (1 if pair[1] < 0x80 else 2), pair.hex()


def test_big5_two_code_point_pairs_and_conventions():
    for pair, exp in ((b"\x88\x62", "Ê̄"), (b"\x88\x64", "Ê̌"), (b"\x88\xa3", "ê̄"), (b"\x88\xa5", "ê̌")):
        assert O.decode(O.ENC_BIG5, [pair])[0][2].decode("utf-8") == exp
    # lead + invalid non-ASCII trail: both consumed; 0x80 / 0xFF alone: malformed, one byte; a lead stays pending
    assert O.decode(O.ENC_BIG5, [b"\xa4\x80x"])[0][:2] == (2, 2)
    assert O.decode(O.ENC_BIG5, [b"\xffx"])[0][:2] == (2, 1)
    r = O.decode(O.ENC_BIG5, [b"ab\xa4", b"\x40c"], last_on_final=False)
    assert r[0][2] == b"ab" and r[-1][2].decode("utf-8") == "一c"


def test_eucjp_conventions():
    assert O.decode(O.ENC_EUC_JP, [b"\x8e\xb1"])[0][2].decode("utf-8") == "ｱ"
    assert O.decode(O.ENC_EUC_JP, [b"\x8f\xaa\xa1"])[0][2].decode("utf-8") == b"\x8f\xaa\xa1".decode("euc_jp")
    assert O.decode(O.ENC_EUC_JP, [b"\x8e\x41"])[0][:2] == (2, 1)      # ASCII after 8E: not consumed
    assert O.decode(O.ENC_EUC_JP, [b"\x8f\x8e\xb1"])[0][:2] == (2, 2)  # 8E after 8F: consumed with it
    assert O.decode(O.ENC_EUC_JP, [b"\x80"])[0][:2] == (2, 1)


@pytest.mark.parametrize("label", ["big5", "euc-jp"])
def test_scanner_finds_planted_text(label):
    m = M.Mission.for_label(label, 4, ubf=M.UBF_ALL_VALID)
    text = "find me 日本語 ΑΒΓ here"
    data = b"\x00\x01\x02" + text.encode("big5hkscs" if label == "big5" else "euc_jp") + b"\x00\xff\x00"
    fc = oracle_state(m).scan_stream(data, False, 4096)
    assert [f.s.decode("utf-8") for f in fc.v] == [text]
