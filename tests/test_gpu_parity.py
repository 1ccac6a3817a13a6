"""GPU parity tests: the CUDA scanner, called through the C ABI, against the CPU oracle.

Bit-exact bar: identical (position, precision, text, completes) lists in identical order, and an
identical ScannerState (leftover, cut flag, consumed bytes) after every call.
"""
import dataclasses
import os
import random

import numpy as np
import pytest

import corpus
import reference_vectors as RV
import stringsext_b200 as sx
from helpers import M, O, oracle_state, to_oracle

pytestmark = pytest.mark.gpu


def gpu_findings(fc):
    return [(f.position, int(f.position_precision), f.s, f.s_completes_previous_s) for f in fc.v]


def oracle_findings(oc):
    return [(f.position, f.precision, f.s, f.completes) for f in oc.v]


def check_state(gs, os_):
    assert gs.last_scan_run_leftover == os_.leftover
    assert gs.last_run_str_was_printed_and_is_maybe_cut_str == os_.cut
    assert gs.consumed_bytes == os_.consumed_bytes


def test_library_is_the_cuda_one():
    assert sx.device_count() >= 1
    st = sx.ScannerState(sx.Mission.for_label("utf-8"))
    st.scan_stream(b"hello world, this is a test\x00" * 100)
    assert st.last_stats.kernel_launches >= 1


@pytest.mark.parametrize("name,mission_factory,calls", RV.SCENARIOS,
                         ids=lambda v: v.split(" ")[0] if isinstance(v, str) else None)
def test_reference_unit_vectors(name, mission_factory, calls):
    """scanner.rs:193-531, finding_collection.rs:431-502 through FindingCollection::from on the GPU."""
    m = mission_factory()
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    for c in calls:
        fc = gs.scan(c["inp"], c["last"], 0)
        got = gpu_findings(fc)
        if c.get("findings") is not None:
            assert [(p, pr, s) for p, pr, s, _ in got] == c["findings"]
        if c.get("first") is not None:
            assert fc.first_byte_position == c["first"]
        assert not fc.str_buf_overflow
        oc = os_.scan(c["inp"], c["last"], 0)
        assert got == oracle_findings(oc)
        check_state(gs, os_)
        if c.get("consumed") is not None:
            assert gs.consumed_bytes == c["consumed"]
        if c.get("leftover") is not None:
            assert gs.last_scan_run_leftover == c["leftover"]


def test_unsupported_missions_fail_loudly():
    import dataclasses as dc

    with pytest.raises(sx.ScannerError) as e:
        sx.ScannerState(dc.replace(RV.SCENARIOS[0][1](), output_line_char_nb_max=5))  # options.rs:33: q >= 6
    assert e.value.code == 3


@pytest.mark.parametrize("n,q,grep,files,expected", [
    (None, 16, 63, ["input1"], "expected_output1"),
    (10, 32, 58, ["input1", "input2"], "expected_output2"),
], ids=["golden1", "golden2"])
def test_cli_golden_1_2(golden_dir, n, q, grep, files, expected):
    """run-tests:11-30 (--grep-char, three encodings, multi-file): scanned on the GPU, merged and printed like
    main.rs:103-141 / finding.rs:112-155 -> the reference's golden files byte for byte."""
    missions = [M.Mission.for_label(lbl, n, M.AF_ALL & ~M.AF_CTRL, M.UBF_COMMON, grep, q, mission_id=i)
                for i, lbl in enumerate(["UTF-8", "utf-16le", "utf-16be"])]
    states = [sx.ScannerState(m) for m in missions]
    out = bytearray(b"\xef\xbb\xbf")
    for fid, name in enumerate(files, start=1):
        data = open(os.path.join(golden_dir, name), "rb").read()
        # the merge is per 4096-byte slice batch in the reference (main.rs:118-136); positions are monotone within a
        # mission, so merging the whole file's collections gives the same order
        fcs = [s.scan_stream(data, False, 4096, fid) for s in states]
        for f in sx.merge(fcs):
            out += f.print(len(files), 3, "x")
    out += b"\n"
    assert bytes(out) == open(os.path.join(golden_dir, expected), "rb").read()


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_differential_fuzz_general_missions(enc):
    """--grep-char / --same-unicode-block on the GPU vs the oracle."""
    rng = random.Random(5150 + enc)
    for _ in range(25):
        m = corpus.random_general_mission(rng, enc, M)
        slice_len = rng.choice([4096, 4096, 1024, 256, 100, 8192])
        gs, os_ = sx.ScannerState(m), oracle_state(m)
        ncalls = rng.choice([1, 2, 3])
        for c in range(ncalls):
            ln = rng.choice([0, 1, 3, 50, 5000]) if rng.random() < 0.4 else rng.randrange(1, 70000)
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), ln, enc)
            last = (c == ncalls - 1) and rng.random() < 0.3
            got = gpu_findings(gs.scan_stream(buf, last, slice_len))
            exp = oracle_findings(os_.scan_stream(buf, last, slice_len)) if ln else []
            assert got == exp, (enc, m, slice_len, ln, last)
            check_state(gs, os_)


@pytest.mark.parametrize("label,n,q,grep,kind", [
    ("utf-8", 6, 64, ":", "rand"), ("utf-8", 10, 64, "e", "rand"), ("ascii", 4, 8, "e", "rand"), ("koi8-r", 3, 8, ":", "runs"),
    ("iso-8859-5", 3, 8, ":", "rand"), ("utf-16le", 4, 16, "e", "rand"), ("utf-8", 3, 8, "a", "mixed"), ("ascii", 8, 16, " ", "text"),
])
def test_prefilter_window_list_matches_spec_grep(label, n, q, grep, kind):
    """--grep-char alone keeps the prefilter with one more rule (a window behind >= q good chars is listed,
    PrefCfg::kill_trail): the kernel lists exactly the windows of the byte-wise specification and the findings are the
    oracle's."""
    import emul

    m = M.Mission.for_label(label, n, (1 << 128) - 2 if q == 8 else None, None, ord(grep), q)
    rng = random.Random(123)
    size = (1 << 20) + 4096
    if kind == "rand":
        buf = corpus.sx_mix_bytes(17, 0, size)
        corpus.plant(buf, 17, m.encoding_id, n, q, density=1 << 12)
        buf = buf.tobytes()
    else:
        buf = corpus.gen(rng, kind, size, m.encoding_id)
    gs, es, os_ = sx.ScannerState(m), emul.EmulState(m, True), oracle_state(m)
    got = gpu_findings(gs.scan_stream(buf, False, 4096))
    exp, _ = es.scan_stream(buf, False, 4096)
    assert gs.last_stats.prefilter_used == 1 and es.stats[7] == 1
    assert gs.last_window_list() == es.last_list
    assert got == exp == oracle_findings(os_.scan_stream(buf, False, 4096))
    check_state(gs, os_)


@pytest.mark.parametrize("label,n,q,ubf,kind,grep", [
    ("utf-8", 6, 64, None, "rand", None), ("utf-8", 10, 64, M.UBF_ALL, "rand", None), ("utf-8", 8, 8, M.UBF_ALL, "blocks", None),
    ("utf-8", 4, 16, M.UBF_ALL, "blocks", None), ("koi8-r", 6, 16, M.UBF_ALL, "blocks", None), ("koi8-r", 6, 64, None, "rand", None),
    ("ascii", 4, 8, None, "rand", None), ("utf-16le", 4, 16, M.UBF_ALL, "blocks", None), ("utf-16be", 10, 64, None, "rand", None),
    ("big5", 6, 16, M.UBF_ALL, "blocks", None), ("euc-jp", 6, 64, None, "rand", None), ("utf-8", 3, 8, None, "mixed", None),
    ("utf-32le", 4, 16, M.UBF_ALL, "blocks", None), ("utf-8", 6, 32, M.UBF_ALL_VALID, "text", None),
    ("utf-8", 4, 8, M.UBF_ALL, "blocks", "a"), ("utf-8", 6, 64, None, "rand", "e"), ("koi8-r", 3, 8, M.UBF_ALL, "blocks", "b"),
    ("utf-16le", 4, 16, M.UBF_ALL, "blocks", "1"), ("big5", 4, 8, M.UBF_ALL, "blocks", "a"),
])
def test_prefilter_window_list_matches_spec_same_block(label, n, q, ubf, kind, grep):
    """--same-unicode-block (alone or with --grep-char) keeps the prefilter with one more rule (PrefCfg::sb_rule: a window whose trailing good
    run may hold a multi-byte char is listed when the trailing run of the window before it or its own leading run may
    hold one too; the pre-roll of a head is the whole window before it): the kernel lists exactly the windows of the
    byte-wise specification and the findings are the oracle's."""
    import dataclasses

    import emul

    m = dataclasses.replace(M.Mission.for_label(label, n, None, ubf, ord(grep) if grep else None, q), require_same_unicode_block=True)
    rng = random.Random(321)
    size = (1 << 20) + 4096
    if kind == "rand":
        buf = corpus.sx_mix_bytes(19, 0, size)
        corpus.plant(buf, 19, m.encoding_id, n, q, density=1 << 12)
        buf = buf.tobytes()
    elif kind == "blocks":
        buf = corpus.gen_blocks(rng, size, m.encoding_id)
    else:
        buf = corpus.gen(rng, kind, size, m.encoding_id)
    gs, es, os_ = sx.ScannerState(m), emul.EmulState(m, True), oracle_state(m)
    got = gpu_findings(gs.scan_stream(buf, False, 4096))
    exp, _ = es.scan_stream(buf, False, 4096)
    assert gs.last_stats.prefilter_used == 1 and es.stats[7] == 1
    assert gs.last_window_list() == es.last_list
    assert got == exp == oracle_findings(os_.scan_stream(buf, False, 4096))
    check_state(gs, os_)


def test_killed_window_case():
    """corpus.KILLED_WINDOW_CASE (a window dropped because of its predecessor's leftover) on the GPU, also embedded in
    a larger stream so that the windows sit in the middle of a block."""
    args, cases = corpus.killed_window_inputs()
    m = M.Mission.for_label(*args)
    for data, expected in cases:
        got = sx.ScannerState(m).scan_stream(data, False, 4096)
        assert [(f.position, f.s) for f in got.v] == expected
        big = b"\x00" * 8192 + data + b"\x00" * 8192
        gs, os_ = sx.ScannerState(m), oracle_state(m)
        assert gpu_findings(gs.scan_stream(big, False, 4096)) == oracle_findings(os_.scan_stream(big, False, 4096))
        assert [(f[0], f[2]) for f in gpu_findings(sx.ScannerState(m).scan_stream(big, False, 4096))] == [(p + 8192, t) for p, t in expected]
        check_state(gs, os_)


def test_field_with_zeros():
    factory, inp = RV.FIELD_WITH_ZEROS
    assert len(sx.ScannerState(factory()).scan(inp, False, 0).v) != 1


def test_cli_golden3(golden_dir):
    """run-tests:33-41: -a None -u None over input1 input2 -> only BOM + newline."""
    missions = [M.Mission.for_label(lbl, None, M.AF_NONE, M.UBF_NONE, None, 32, mission_id=i)
                for i, lbl in enumerate(["UTF-8", "utf-16le", "utf-16be"])]
    states = [sx.ScannerState(m) for m in missions]
    out = bytearray(b"\xef\xbb\xbf")
    for fid, name in enumerate(["input1", "input2"], start=1):
        data = open(os.path.join(golden_dir, name), "rb").read()
        fcs = [s.scan_stream(data, False, 4096, fid) for s in states]
        for f in sx.merge(fcs):
            out += f.print(2, 3, "x")
    out += b"\n"
    assert bytes(out) == open(os.path.join(golden_dir, "expected_output3"), "rb").read()


@pytest.mark.parametrize("label,n,q,ubf", [
    ("UTF-8", 4, 16, M.UBF_COMMON), ("utf-16le", 4, 16, M.UBF_COMMON), ("utf-16be", 10, 32, M.UBF_COMMON),
    ("ascii", 4, 64, None), ("utf-8", 10, 64, M.UBF_ALL_VALID), ("utf-16le", 6, 64, M.UBF_ALL_VALID),
    ("koi8-r", 6, 64, None),
])
def test_reference_fixture_files(golden_dir, label, n, q, ubf):
    """The reference's functional-test inputs (text + EFI binary), multi-file stream with carry across files."""
    m = M.Mission.for_label(label, n, None, ubf, None, q)
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    for fid, name in enumerate(["input1", "input2"], start=1):
        data = open(os.path.join(golden_dir, name), "rb").read()
        got = gpu_findings(gs.scan_stream(data, False, 4096, fid))
        exp = oracle_findings(os_.scan_stream(data, False, 4096, fid))
        assert got == exp
        check_state(gs, os_)


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_differential_fuzz(enc):
    rng = random.Random(4242 + enc)
    for _ in range(40):
        m = corpus.random_mission(rng, enc, M)
        slice_len = rng.choice([4096, 4096, 4096, 1024, 256, 100, 33, 17, 8192])
        gs, os_ = sx.ScannerState(m), oracle_state(m)
        ncalls = rng.choice([1, 1, 2, 3])
        for c in range(ncalls):
            ln = rng.choice([0, 1, 2, 3, 5, 50, 500, 5000]) if rng.random() < 0.5 else rng.randrange(1, 70000)
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), ln, enc)
            last = (c == ncalls - 1) and rng.random() < 0.3
            got = gpu_findings(gs.scan_stream(buf, last, slice_len))
            exp = oracle_findings(os_.scan_stream(buf, last, slice_len)) if ln else []
            assert got == exp, (enc, m, slice_len, ln, last)
            check_state(gs, os_)


@pytest.mark.parametrize("label,n,size_mib,seed", [
    ("ascii", 4, 64, 1),      # BASELINE config 1
    ("utf-8", 10, 64, 2),     # config 2 at an oracle-friendly size
    ("utf-16le", 10, 16, 3),  # config 3 (ubf African), planted corpus carries the parity
    ("utf-16be", 10, 16, 3),
    ("koi8-r", 6, 16, 5),     # config 5 member with heavy output
    ("utf-32le", 6, 16, 5),   # extension
    ("big5", 8, 16, 4),       # config 4 member (table provenance: DESIGN.md section 2)
    ("euc-jp", 6, 16, 5),     # config 5 member
])
def test_baseline_configs_vs_oracle(label, n, size_mib, seed):
    ubf = M.UBF_AFRICAN if label.startswith("utf-16") else None
    m = M.Mission.for_label(label, n, ubf=ubf)
    size = size_mib << 20
    buf = corpus.sx_mix_bytes(seed, 0, size)
    corpus.plant(buf, seed, m.encoding_id, n, 64)
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    raw = gs.scan_stream(buf, False, 4096, raw=True)
    got = raw.all()
    exp = oracle_findings(os_.scan_stream(buf, False, 4096))
    assert len(got) == len(exp)
    assert got == exp
    check_state(gs, os_)
    assert len(exp) > 0


def test_chained_calls_equal_one_call():
    """State hand-off: one sx_scan_stream call == k chained calls cut at slice multiples (and at odd places)."""
    m = M.Mission.for_label("utf-8", 4, ubf=M.UBF_ALL_VALID)
    buf = corpus.sx_mix_bytes(9, 0, 1 << 20)
    corpus.plant(buf, 9, 1, 4, 64, density=1 << 12)
    one = gpu_findings(sx.ScannerState(m).scan_stream(buf, False, 4096))
    gs = sx.ScannerState(m)
    parts = []
    cuts = [0, 4096 * 3, 4096 * 64, 4096 * 65, 4096 * 200, len(buf)]
    for a, b in zip(cuts, cuts[1:]):
        parts += gpu_findings(gs.scan_stream(buf[a:b], False, 4096))
    assert parts == one
    # cut at non-slice multiples: compare with the oracle run the same way (the slice grid restarts per call)
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    cuts = [0, 1000, 1001, 5000, 70001, len(buf)]
    for a, b in zip(cuts, cuts[1:]):
        got = gpu_findings(gs.scan_stream(buf[a:b], False, 4096))
        exp = oracle_findings(os_.scan_stream(buf[a:b], False, 4096))
        assert got == exp
        check_state(gs, os_)


def test_device_resident_input_and_fill():
    """Device-pointer entry + sx_fill_random reproduces corpus.sx_mix_bytes."""
    import torch

    n = (8 << 20) + 123
    t = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    L = sx.load_library()
    assert L.sx_fill_random(t.data_ptr(), n, 77, 0, 0, None) == 0
    torch.cuda.synchronize()
    host = t.cpu().numpy()
    assert np.array_equal(host, corpus.sx_mix_bytes(77, 0, n))
    m = M.Mission.for_label("utf-8", 6)
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    got = gpu_findings(gs.scan_stream(None, False, 4096, device_ptr=t.data_ptr(), length=n))
    exp = oracle_findings(os_.scan_stream(host, False, 4096))
    assert got == exp and len(exp) > 100
    check_state(gs, os_)


# ---- prefilter ---------------------------------------------------------------------------------
import sys as _sys

_sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))


@pytest.mark.parametrize("label,n,q,ubf,kind", [
    ("utf-8", 10, 64, None, "rand"), ("utf-8", 6, 64, None, "rand"), ("utf-8", 10, 64, M.UBF_ALL_VALID, "mixed"),
    ("ascii", 6, 64, None, "rand"), ("utf-16le", 10, 64, M.UBF_AFRICAN, "rand"), ("utf-16be", 6, 32, None, "text"),
    ("utf-16le", 4, 64, M.UBF_ALL_VALID, "mixed"), ("utf-32le", 6, 64, None, "rand"), ("utf-32be", 4, 16, None, "text"),
    ("koi8-r", 10, 64, M.UBF_NONE, "rand"), ("windows-1252", 8, 64, None, "mixed"), ("utf-8", 4, 8, None, "mixed"),
    ("big5", 8, 64, None, "rand"), ("big5", 4, 64, M.UBF_ALL_VALID, "text"), ("euc-jp", 6, 64, None, "rand"), ("euc-jp", 4, 32, M.UBF_ALL, "mixed"),
])
def test_prefilter_window_list_matches_spec(label, n, q, ubf, kind):
    """The SWAR prefilter must list exactly the windows its byte-wise specification
    (sx_core.cuh pref_*_ref, run on the CPU by tests/emul) lists."""
    import emul

    m = M.Mission.for_label(label, n, ubf=ubf, output_line_char_nb_max=q)
    rng = random.Random(99)
    size = (1 << 20) + 77
    if kind == "rand":
        buf = corpus.sx_mix_bytes(5, 0, size)
        corpus.plant(buf, 5, m.encoding_id, n, q, density=1 << 13)
        buf = buf.tobytes()
    else:
        buf = corpus.gen(rng, kind, size, m.encoding_id)
    for pend_prefix in (b"", b"\xe2" if label == "utf-8" else b"\xa4" if label in ("big5", "euc-jp") else b"\x41"):
        gs = sx.ScannerState(m)
        es = emul.EmulState(m, True)
        if pend_prefix:  # shift the unit grid / leave a pending sequence from a previous call
            gs.scan_stream(pend_prefix, False, 4096)
            es.scan_stream(pend_prefix, False, 4096)
        got = gpu_findings(gs.scan_stream(buf, False, 4096))
        exp, _ = es.scan_stream(buf, False, 4096)
        assert gs.last_stats.prefilter_used == 1 and es.stats[7] == 1
        assert gs.last_stats.tma_used == 1  # tiles staged with cp.async.bulk.tensor
        assert gs.last_window_list() == es.last_list
        assert got == exp
        # the plain-load staging path must list the same windows
        g2 = sx.ScannerState(m)
        g2.set_tma(False)
        if pend_prefix:
            g2.scan_stream(pend_prefix, False, 4096)
        got2 = gpu_findings(g2.scan_stream(buf, False, 4096))
        assert g2.last_stats.tma_used == 0 and g2.last_window_list() == es.last_list and got2 == exp


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_prefilter_on_off_identical(enc):
    rng = random.Random(777 + enc)
    for _ in range(12):
        m = corpus.random_mission(rng, enc, M)
        import dataclasses

        q = rng.choice([8, 16, 32, 64, 64])
        m = dataclasses.replace(m, output_line_char_nb_max=q, chars_min_nb=min(m.chars_min_nb, q))
        buf = corpus.gen(rng, rng.choice(corpus.KINDS), rng.randrange(1, 300000), enc)
        a, b = sx.ScannerState(m), sx.ScannerState(m)
        b.set_prefilter(False)
        ra = gpu_findings(a.scan_stream(buf, False, 4096))
        rb = gpu_findings(b.scan_stream(buf, False, 4096))
        assert a.last_stats.prefilter_used == 1 and b.last_stats.prefilter_used == 0
        assert ra == rb
        assert a.last_scan_run_leftover == b.last_scan_run_leftover
        assert a.last_run_str_was_printed_and_is_maybe_cut_str == b.last_run_str_was_printed_and_is_maybe_cut_str


@pytest.mark.parametrize("enc", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_general_missions_large_buffers_vs_oracle(enc):
    """--grep-char / --same-unicode-block / n > q on buffers of many 128-entry blocks (the block kernel's warm-up finds
    a known carry through the WT_GUARD rules, sx_core.cuh guard_benign / guard_known_behind): the oracle's findings,
    on sparse input (random bytes + planted strings) and dense input, with and without the prefilter (which --grep-char
    and --same-unicode-block missions use, PrefCfg::kill_trail / sb_rule, and n > q does not)."""
    import dataclasses

    rng = random.Random(4242 + enc)
    for it in range(8):
        m = corpus.random_general_mission(rng, enc, M)
        q = rng.choice([8, 16, 32, 64, 64])
        n = min(m.chars_min_nb, q) if rng.random() < 0.8 else q + rng.choice([1, 3, 10])
        m = dataclasses.replace(m, output_line_char_nb_max=q, chars_min_nb=n)
        if it < 3:
            arr = corpus.sx_mix_bytes(900 + it, 0, (1 << 20) + 4096 * it)
            corpus.plant(arr, it, enc, max(1, min(n, q)), q, density=1 << 11)
            buf = arr.tobytes()
        else:
            buf = corpus.gen(rng, rng.choice(corpus.KINDS), rng.randrange(1, 300000), enc)
        a, b, os_ = sx.ScannerState(m), sx.ScannerState(m), oracle_state(m)
        b.set_prefilter(False)
        cut = rng.randrange(1, len(buf)) if len(buf) > 1 and rng.random() < 0.5 else len(buf)
        for part in (buf[:cut], buf[cut:]):
            if not part:
                continue
            ra = gpu_findings(a.scan_stream(part, False, 4096))
            rb = gpu_findings(b.scan_stream(part, False, 4096))
            assert a.last_stats.prefilter_used == (1 if n <= q else 0) and b.last_stats.prefilter_used == 0
            exp = oracle_findings(os_.scan_stream(part, False, 4096))
            assert ra == exp, (enc, m, len(part))
            assert rb == exp, (enc, m, len(part))
            check_state(a, os_)
            check_state(b, os_)


@pytest.mark.parametrize("n,q,ubf,kind", [(10, 64, None, "rand"), (6, 64, None, "rand"), (4, 64, M.UBF_ALL_VALID, "rand"),
                                         (10, 64, M.UBF_ALL_VALID, "mixed"), (3, 16, None, "rand"), (8, 8, M.UBF_AFRICAN, "text"),
                                         (2, 64, None, "rand"), (16, 32, M.UBF_ALL, "runs")])
def test_sparse_pipeline_matches_block_kernel_and_oracle(n, q, ubf, kind):
    """UTF-8: the barrier-free sparse-list pipeline (mask engine) == the block kernel == the oracle, incl. chained calls."""
    m = M.Mission.for_label("utf-8", n, ubf=ubf, output_line_char_nb_max=q)
    rng = random.Random(4242 + n + q)
    size = (3 << 20) + 4096 * 3 + 17
    if kind == "rand":
        buf = corpus.sx_mix_bytes(11, 0, size)
        corpus.plant(buf, 11, 1, n, q, density=1 << 13)
        buf = buf.tobytes()
    else:
        # mostly binary with embedded text, so the list stays sparse
        parts, total = [], 0
        while total < size:
            ln = rng.randrange(2000, 60000)
            parts.append(corpus.sx_mix_bytes(rng.randrange(1 << 30), 0, ln).tobytes())
            parts.append(corpus.gen(rng, kind, rng.randrange(1, 600), 1))
            total += ln + len(parts[-1])
        buf = b"".join(parts)[:size]
    a, b, os_ = sx.ScannerState(m), sx.ScannerState(m), oracle_state(m)
    b.set_sparse(False)
    cuts = [0, 4096 * 100 + 5, 4096 * 300 + 5, len(buf)]
    used = 0
    for lo, hi in zip(cuts, cuts[1:]):
        last = hi == len(buf) and n == 6
        ra = gpu_findings(a.scan_stream(buf[lo:hi], last, 4096))
        rb = gpu_findings(b.scan_stream(buf[lo:hi], last, 4096))
        exp = oracle_findings(os_.scan_stream(buf[lo:hi], last, 4096))
        used += a.last_stats.sparse_used
        assert b.last_stats.sparse_used == 0
        assert ra == exp and rb == exp
        check_state(a, os_)
        check_state(b, os_)
    assert used >= 1 or kind != "rand" or n <= 6  # short minimum lengths list too many windows for the sparse path


@pytest.mark.parametrize("enc", [0, 1, 2, 4, 5])
def test_direct_host_output_matches_record_download(enc):
    """Findings written by the GPU into the collection's pinned set (both exact-stage variants) == records
    downloaded and converted on the host == oracle; collections stay valid after later scans (own their set)."""
    rng = random.Random(900 + enc)
    kept = []
    for it in range(6):
        m = corpus.random_mission(rng, enc, M)
        m = dataclasses.replace(m, output_line_char_nb_max=64, chars_min_nb=min(m.chars_min_nb, 64))
        kind = rng.choice(["rand", "mixed", "text"])
        size = rng.randrange(1, 600000)
        buf = corpus.gen(rng, kind, size, enc) if kind != "rand" else corpus.sx_mix_bytes(it, 0, size).tobytes()
        a, b, os_ = sx.ScannerState(m), sx.ScannerState(m), oracle_state(m)
        b.set_direct_output(False)
        cut = rng.randrange(0, len(buf) + 1)
        for part in (buf[:cut], buf[cut:]):
            fa = a.scan_stream(part, False, 4096)
            ra, rb = gpu_findings(fa), gpu_findings(b.scan_stream(part, False, 4096))
            exp = oracle_findings(os_.scan_stream(part, False, 4096)) if len(part) else []
            assert ra == exp and rb == exp
            check_state(a, os_)
            check_state(b, os_)
            kept.append((fa, exp))
    for fa, exp in kept:  # earlier collections are untouched by the scans that followed
        assert gpu_findings(fa) == exp


@pytest.mark.parametrize("n,q,ubf,kind", [(10, 64, None, "text"), (4, 64, M.UBF_ALL_VALID, "text"), (6, 64, M.UBF_ALL, "mixed"),
                                         (3, 64, None, "runs"), (8, 64, M.UBF_ALL_VALID, "lowent"), (10, 64, M.UBF_AFRICAN, "text"),
                                         (6, 32, M.UBF_ALL_VALID, "text"), (2, 64, None, "mixed")])
def test_sparse_pipeline_forced_on_dense_input(n, q, ubf, kind):
    """UTF-8 text-like input (every window listed, long runs of adjacent windows): the sparse-list pipeline forced on
    (carry-independence rule, closed-form walks) == the block kernel == the oracle, incl. chained calls and is_last."""
    m = M.Mission.for_label("utf-8", n, ubf=ubf, output_line_char_nb_max=q)
    rng = random.Random(777 + n + q)
    buf = corpus.gen(rng, kind, 700000 + rng.randrange(5000), 1)
    a, b, os_ = sx.ScannerState(m), sx.ScannerState(m), oracle_state(m)
    a.set_sparse(2)
    b.set_sparse(0)
    cuts = [0, 4096 * 50 + 3, 4096 * 120, len(buf)]
    for lo, hi in zip(cuts, cuts[1:]):
        last = hi == len(buf)
        ra = gpu_findings(a.scan_stream(buf[lo:hi], last, 4096))
        rb = gpu_findings(b.scan_stream(buf[lo:hi], last, 4096))
        exp = oracle_findings(os_.scan_stream(buf[lo:hi], last, 4096))
        assert a.last_stats.sparse_used == 1 and b.last_stats.sparse_used == 0
        assert ra == exp and rb == exp
        check_state(a, os_)
        check_state(b, os_)


@pytest.mark.parametrize("label,n,q,kind", [("ascii", 4, 64, "rand"), ("ascii", 6, 64, "mixed"), ("koi8-r", 6, 64, "rand"),
                                           ("windows-1252", 8, 64, "text"), ("koi8-r", 3, 32, "runs"), ("ascii", 10, 64, "text"),
                                           ("ibm866", 2, 64, "lowent"), ("iso-8859-5", 16, 16, "mixed")])
def test_sparse_pipeline_single_byte_encodings(label, n, q, kind):
    """x-user-defined / table-driven single-byte missions: the per-stage pipeline with the single-byte mask engine
    == the block kernel == the oracle, incl. chained calls and is_last."""
    m = M.Mission.for_label(label, n, output_line_char_nb_max=q)
    rng = random.Random(1234 + n + q)
    size = 900000 + rng.randrange(5000)
    if kind == "rand":
        buf = corpus.sx_mix_bytes(21, 0, size)
        corpus.plant(buf, 21, m.encoding_id, n, q, density=1 << 13)
        buf = buf.tobytes()
    else:
        buf = corpus.gen(rng, kind, size, m.encoding_id)
    a, b, os_ = sx.ScannerState(m), sx.ScannerState(m), oracle_state(m)
    b.set_sparse(0)
    cuts = [0, 4096 * 60 + 1, 4096 * 130, len(buf)]
    used = 0
    for lo, hi in zip(cuts, cuts[1:]):
        last = hi == len(buf)
        ra = gpu_findings(a.scan_stream(buf[lo:hi], last, 4096))
        rb = gpu_findings(b.scan_stream(buf[lo:hi], last, 4096))
        exp = oracle_findings(os_.scan_stream(buf[lo:hi], last, 4096))
        used += a.last_stats.sparse_used
        assert b.last_stats.sparse_used == 0
        assert ra == exp and rb == exp
        check_state(a, os_)
        check_state(b, os_)
    assert used >= 1 or q != 64
