"""CPU differential test: the product's per-window automaton and window decomposition
(stringsext_b200/csrc/sx_core.cuh, compiled for the host by tests/emul/) against the oracle.

This validates the *algorithm* the CUDA kernels run (look-back decoder state, transfer-function
classification, emit rule, record/text ranges, ScannerState hand-off) without a GPU.  The
kernels themselves are compared with the oracle in test_gpu_parity.py.
"""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))

import corpus
import emul
import reference_vectors as RV
from helpers import M, oracle_state


def _cmp(es, os_, f, o):
    o_v = [(x.position, x.precision, x.s, x.completes) for x in o]
    assert f == o_v
    assert es.leftover == os_.leftover
    assert es.cut == os_.cut
    assert es.consumed == os_.consumed_bytes


@pytest.mark.parametrize("name,mission_factory,calls", [s for s in RV.SCENARIOS if "grep" not in s[0]],
                         ids=lambda v: v.split(" ")[0] if isinstance(v, str) else None)
def test_reference_vectors(name, mission_factory, calls):
    m = mission_factory()
    es, os_ = emul.EmulState(m), oracle_state(m)
    for c in calls:
        f, first = es.scan_stream(c["inp"], c["last"], 4096)
        o = os_.scan(c["inp"], c["last"], 0)
        assert first == o.first_byte_position
        _cmp(es, os_, f, o.v)
    assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz(enc, seed):
    rng = random.Random(seed * 100 + enc)
    for _ in range(60):
        m = corpus.random_mission(rng, enc, M)
        slice_len = rng.choice([4096, 4096, 4096, 1024, 256, 100, 33, 17, 8192])
        es, os_ = emul.EmulState(m), oracle_state(m)
        ncalls = rng.choice([1, 1, 2, 3])
        for c in range(ncalls):
            ln = rng.choice([0, 1, 2, 3, 5, 50, 500, 5000]) if rng.random() < 0.5 else rng.randrange(1, 9000)
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), ln, enc)
            last = (c == ncalls - 1) and rng.random() < 0.3
            f, _ = es.scan_stream(buf, last, slice_len)
            o = os_.scan_stream(buf, last, slice_len).v if ln else []
            _cmp(es, os_, f, o)
        # classification self-checks of the harness: replay mismatches / missed emits / count shortcuts / text lengths
        assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_fuzz_general_missions(enc):
    """--grep-char / --same-unicode-block: general automaton + dual-simulation classification (classify_general)."""
    rng = random.Random(31337 + enc)
    for _ in range(50):
        m = corpus.random_general_mission(rng, enc, M)
        slice_len = rng.choice([4096, 4096, 1024, 256, 100, 33, 8192])
        es, os_ = emul.EmulState(m), oracle_state(m)
        ncalls = rng.choice([1, 1, 2, 3])
        for c in range(ncalls):
            ln = rng.choice([0, 1, 2, 3, 5, 50, 500, 5000]) if rng.random() < 0.5 else rng.randrange(1, 9000)
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), ln, enc)
            last = (c == ncalls - 1) and rng.random() < 0.3
            f, _ = es.scan_stream(buf, last, slice_len)
            o = os_.scan_stream(buf, last, slice_len).v if ln else []
            _cmp(es, os_, f, o)
        assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0


def test_general_missions_guard_rules():
    """WT_GUARD (sx_core.cuh): a general-mission window whose first run ends inside it keeps the carry-out seen under
    the null carry unless the carried leftover fills that run up to q chars (guard_benign), and behind an adjacent
    guard / constant window that holds for every possible carry-in (guard_known_behind, what ends the block kernel's
    warm-up).  The harness checks both claims against the replay under the real carry on every window (stats[3]);
    small q and n make the "killer" leftovers frequent.  --grep-char and --same-unicode-block missions use the
    prefilter, each option with the extra listing rule the fuzz's counterexamples asked for;
    chars_min_nb > q does not."""
    import ctypes as C
    import dataclasses

    L = emul.lib()
    L.sx_emul_guard_behind.restype = C.c_uint64
    L.sx_emul_guard_behind_hard.restype = C.c_uint64
    ok0, k0 = C.c_uint64(), C.c_uint64()
    L.sx_emul_guard_counts(C.byref(ok0), C.byref(k0))
    b0, h0 = L.sx_emul_guard_behind(), L.sx_emul_guard_behind_hard()
    for enc in (0, 1, 4, 2):
        rng = random.Random(31000 + enc)
        for _ in range(40):
            m = corpus.random_general_mission(rng, enc, M)
            q = rng.choice([8, 8, 16, 32, 64])
            m = dataclasses.replace(m, output_line_char_nb_max=q, chars_min_nb=rng.choice([1, 2, 3, 4, 6, 8, 10, 12, 20]))
            es, os_ = emul.EmulState(m), oracle_state(m)
            for c in range(rng.choice([1, 2])):
                buf = corpus.gen(rng, rng.choice(corpus.KINDS), rng.randrange(1, 30000), enc)
                f, _ = es.scan_stream(buf, False, 4096)
                _cmp(es, os_, f, os_.scan_stream(buf, False, 4096).v)
                # prefilter for --grep-char (PrefCfg::kill_trail) and --same-unicode-block (PrefCfg::sb_rule), not for n > q
                assert es.stats[7] == (1 if m.chars_min_nb <= q else 0)
            assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0
    ok1, k1 = C.c_uint64(), C.c_uint64()
    L.sx_emul_guard_counts(C.byref(ok1), C.byref(k1))
    assert ok1.value - ok0.value > 5000 and k1.value - k0.value > 100
    assert L.sx_emul_guard_behind() - b0 > 5000 and L.sx_emul_guard_behind_hard() - h0 > 20


@pytest.mark.parametrize("enc", [0, 1, 2, 4, 5])
def test_fuzz_grep_missions_with_prefilter(enc):
    """--grep-char alone keeps the prefilter (PrefCfg::kill_trail: the window behind >= q good chars is listed).  Small
    q makes q-char leftovers -- the ones that kill or "maybe cut" the next window -- frequent; the first version of the
    rule (without the pending-bytes clause) failed 6 of 12 000 such missions, UTF-8 and UTF-16."""
    import dataclasses

    rng = random.Random(7100 + enc)
    used = 0
    for _ in range(60):
        m = corpus.random_mission(rng, enc, M)
        grep = rng.choice([0x20, ord("a"), ord("e"), ord(":"), ord("?"), 0x00])
        q = rng.choice([64, 32, 16, 8, 8, 8])
        m = dataclasses.replace(m, filter=dataclasses.replace(m.filter, grep_char=grep), require_same_unicode_block=False,
                                output_line_char_nb_max=q, chars_min_nb=min(rng.choice([1, 2, 3, 3, 4, 6, 8, 10]), q))
        es, os_ = emul.EmulState(m), oracle_state(m)
        for c in range(rng.choice([1, 1, 2, 3])):
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), rng.randrange(1, 40000), enc)
            sl = rng.choice([4096, 4096, 1024, 8192])
            f, _ = es.scan_stream(buf, False, sl)
            _cmp(es, os_, f, os_.scan_stream(buf, False, sl).v)
            used += es.stats[7]
        assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0
    assert used > 60  # the prefilter was on for most calls


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_fuzz_same_block_missions_with_prefilter(enc):
    """--same-unicode-block (here: alone and, every third mission, together with --grep-char) keeps the prefilter (PrefCfg::sb_rule: a window whose trailing good run and the one
    of the window before it may both hold a multi-byte char is listed; the pre-roll of a head is the whole window before
    it).  Small q makes window boundaries frequent; corpus.gen_blocks mixes blocks inside runs behind ASCII junk."""
    import dataclasses

    rng = random.Random(9300 + enc)
    used = 0
    for _ in range(60):
        m = corpus.random_mission(rng, enc, M)
        q = rng.choice([64, 32, 16, 8, 8, 8])
        ubf = rng.choice([M.UBF_ALL, M.UBF_ALL_VALID, M.UBF_ALL, m.filter.ubf])
        grep = rng.choice([None, None, ord("a"), ord("e"), 0x20, ord("1")])
        m = dataclasses.replace(m, filter=dataclasses.replace(m.filter, grep_char=grep, ubf=ubf), require_same_unicode_block=True,
                                output_line_char_nb_max=q, chars_min_nb=min(rng.choice([1, 2, 3, 3, 4, 6, 8, 10]), q))
        es, os_ = emul.EmulState(m), oracle_state(m)
        for c in range(rng.choice([1, 1, 2, 3])):
            ln = rng.randrange(1, 40000)
            buf = corpus.gen_blocks(rng, ln, enc) if rng.random() < 0.6 else corpus.gen(rng, rng.choice(corpus.KINDS), ln, enc)
            sl = rng.choice([4096, 4096, 1024, 8192])
            f, _ = es.scan_stream(buf, False, sl)
            _cmp(es, os_, f, os_.scan_stream(buf, False, sl).v)
            used += es.stats[7]
        assert es.stats[3] == 0 and es.stats[4] == 0 and es.stats[5] == 0 and es.stats[6] == 0
    assert used > 60  # the prefilter was on for most calls


def test_killed_window_case():
    """corpus.KILLED_WINDOW_CASE: the oracle drops the segment behind a q-char leftover without the grep char
    (helper.rs:389-415), the harness follows."""
    args, cases = corpus.killed_window_inputs()
    m = M.Mission.for_label(*args)
    for data, expected in cases:
        es, os_ = emul.EmulState(m), oracle_state(m)
        o = os_.scan_stream(data, False, 4096).v
        assert [(x.position, x.s) for x in o] == expected
        f, _ = es.scan_stream(data, False, 4096)
        _cmp(es, os_, f, o)
        assert es.stats[7] == 1 and 2 in es.last_list  # --grep-char alone: prefilter on, the killed window is listed


def test_planted_corpus_utf16():
    """Random bytes + planted UTF-16 strings (random alone yields nothing)."""
    for enc, label in ((2, "utf-16le"), (3, "utf-16be")):
        m = M.Mission.for_label(label, 10, ubf=M.UBF_AFRICAN)
        buf = corpus.sx_mix_bytes(3, 0, 1 << 18)
        corpus.plant(buf, 3, enc, 10, 64, density=1 << 12)
        es, os_ = emul.EmulState(m), oracle_state(m)
        f, _ = es.scan_stream(buf.tobytes(), False, 4096)
        o = os_.scan_stream(buf, False, 4096).v
        assert len(o) > 5
        _cmp(es, os_, f, o)


@pytest.mark.parametrize("enc", [1, 0, 4])
def test_mask_engine_cross_checks(enc):
    """The bit-parallel window engines (sx_mask_utf8.cuh: UTF-8 and the single-byte family) run on the CPU by the
    harness: every call is cross-checked against the byte-wise engine (results and records), the one-pass head against
    pre-roll + pass, the carry-independence claim and the closed form (eval_caseb) against a replay under the real
    carry; the findings against the oracle."""
    import ctypes as C
    import dataclasses

    L = emul.lib()
    L.sx_emul_mask_mismatches.restype = C.c_uint64
    L.sx_emul_head_ok.restype = C.c_uint64
    ok0, dec0 = C.c_uint64(), C.c_uint64()
    L.sx_emul_mask_counts(C.byref(ok0), C.byref(dec0))
    heads0, mism0 = L.sx_emul_head_ok(), L.sx_emul_mask_mismatches()
    rng = random.Random(2024 + enc)
    for it in range(24):
        m = corpus.random_mission(rng, enc, M)
        q = rng.choice([64, 64, 64, 32, 16, 8])
        m = dataclasses.replace(m, output_line_char_nb_max=q, chars_min_nb=min(rng.choice([1, 2, 3, 4, 6, 10, 16]), q))
        kind = rng.choice(["rand", "rand", "mixed", "lowent", "runs", "text"])
        size = rng.choice([1 << 16, (1 << 17) + 77, 200000])
        if kind == "rand":
            buf = corpus.sx_mix_bytes(it + 500, 0, size)
            corpus.plant(buf, it, enc, m.chars_min_nb, q, density=1 << 12)
            buf = buf.tobytes()
        else:
            buf = corpus.gen(rng, kind, size, enc)
        es, os_ = emul.EmulState(m), oracle_state(m)
        cuts = [0] + sorted(rng.sample(range(1, len(buf)), rng.choice([0, 1, 2]))) + [len(buf)]
        for a, b in zip(cuts, cuts[1:]):
            f, _ = es.scan_stream(buf[a:b], False, rng.choice([4096, 4096, 512]))
            # (the oracle must see the same slicing)
        es2, os2 = emul.EmulState(m), oracle_state(m)
        for a, b in zip(cuts, cuts[1:]):
            f, _ = es2.scan_stream(buf[a:b], False, 4096)
            o = os2.scan_stream(buf[a:b], False, 4096).v
            _cmp(es2, os2, f, o)
        assert es.stats[3] == 0 and es2.stats[3] == 0 and es2.stats[4] == 0 and es2.stats[5] == 0 and es2.stats[6] == 0
    ok1, dec1 = C.c_uint64(), C.c_uint64()
    L.sx_emul_mask_counts(C.byref(ok1), C.byref(dec1))
    assert L.sx_emul_mask_mismatches() == mism0          # mask engine == byte-wise engine on every call it accepted
    assert ok1.value - ok0.value > 10000                 # ... and it accepted most of them
    assert ok1.value - ok0.value > 3 * (dec1.value - dec0.value)
    assert L.sx_emul_head_ok() > heads0                  # the one-pass head path was exercised
