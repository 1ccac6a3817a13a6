"""GPU tests of the pieces / ranges machinery: a call cut into pipelined pieces, and one mission's
stream scanned range by range (sx_scan_range), must give exactly the findings of the plain fold -- checked against the
CPU oracle -- whatever the cut points are."""
import dataclasses
import random

import pytest

import corpus
import stringsext_b200 as sx
from helpers import M, oracle_state

pytestmark = pytest.mark.gpu


def gpu_findings(fc):
    return [(f.position, int(f.position_precision), f.s, f.s_completes_previous_s) for f in fc.v]


def oracle_findings(oc):
    return [(f.position, f.precision, f.s, f.completes) for f in oc.v]


def check_state(gs, os_):
    assert gs.last_scan_run_leftover == os_.leftover
    assert gs.last_run_str_was_printed_and_is_maybe_cut_str == os_.cut
    assert gs.consumed_bytes == os_.consumed_bytes


def make_buf(kind, size, enc, n, q, seed):
    rng = random.Random(seed)
    if kind == "rand":
        buf = corpus.sx_mix_bytes(seed, 0, size)
        corpus.plant(buf, seed, enc, n, q, density=1 << 12)
        return buf.tobytes()
    return corpus.gen(rng, kind, size, enc)


CASES = [("utf-8", 10, 64, None, "rand"), ("utf-8", 4, 64, M.UBF_ALL_VALID, "text"), ("utf-8", 6, 32, M.UBF_ALL, "mixed"),
         ("ascii", 4, 64, None, "rand"), ("koi8-r", 6, 64, None, "rand"), ("koi8-r", 3, 64, None, "text"),
         ("windows-1252", 8, 16, None, "runs"), ("utf-8", 2, 64, None, "lowent")]


@pytest.mark.parametrize("label,n,q,ubf,kind", CASES)
def test_pieces_equal_the_fold(label, n, q, ubf, kind):
    """The sparse pipeline cut into K pieces (prefilter of piece k+1 beside the exact stage of piece k, carry into each
    piece from sx_range_carry_kernel): identical findings and ScannerState for every K, chained calls included."""
    m = M.Mission.for_label(label, n, ubf=ubf, output_line_char_nb_max=q)
    size = (3 << 20) + 4096 * 5 + 321
    buf = make_buf(kind, size, m.encoding_id, n, q, 31 + n + q)
    cuts = [0, 4096 * 311 + 7, len(buf)]
    for pieces in (1, 3, 7, 32):
        gs, os_ = sx.ScannerState(m), oracle_state(m)
        gs.set_pieces(pieces)
        for lo, hi in zip(cuts, cuts[1:]):
            last = hi == len(buf) and pieces == 3
            got = gpu_findings(gs.scan_stream(buf[lo:hi], last, 4096))
            exp = oracle_findings(os_.scan_stream(buf[lo:hi], last, 4096))
            assert gs.last_stats.sparse_used == 1
            assert gs.last_stats.pieces == min(pieces, gs.last_stats.pieces)
            assert got == exp, (label, pieces, lo, hi)
            check_state(gs, os_)


RANGE_CASES = CASES + [("utf-16le", 4, 64, M.UBF_ALL_VALID, "mixed"), ("utf-16be", 6, 32, None, "text"), ("utf-32le", 6, 64, None, "rand"),
                       ("utf-16le", 10, 64, M.UBF_AFRICAN, "rand"), ("big5", 8, 64, None, "rand"), ("big5", 4, 64, M.UBF_ALL_VALID, "text"),
                       ("euc-jp", 6, 64, None, "rand"), ("euc-jp", 3, 32, M.UBF_ALL, "mixed")]


@pytest.mark.parametrize("label,n,q,ubf,kind", RANGE_CASES)
def test_ranges_concatenate_to_the_whole_scan(label, n, q, ubf, kind):
    """concat(sx_scan_range pieces) == sx_scan_stream == oracle, cut at random slice multiples; the ScannerState moves on
    only with the range that reaches the end of the buffer, and then exactly like the fold."""
    m = M.Mission.for_label(label, n, ubf=ubf, output_line_char_nb_max=q)
    size = (2 << 20) + 4096 * 3 + 1234
    buf = make_buf(kind, size, m.encoding_id, n, q, 77 + n + q)
    os_ = oracle_state(m)
    exp = oracle_findings(os_.scan_stream(buf, False, 4096))
    rng = random.Random(5 + n)
    for trial in range(3):
        nslices = len(buf) // 4096
        cuts = sorted({0, len(buf)} | {4096 * rng.randrange(1, nslices) for _ in range(rng.choice([1, 2, 5]))})
        if trial == 0:
            cuts = sorted(set(cuts) | {4096, 8192})  # tiny first ranges
        gs = sx.ScannerState(m)
        got = []
        for lo, hi in zip(cuts, cuts[1:]):
            before = gs.consumed_bytes
            fc = gs.scan_stream(buf, False, 4096, lo=lo, hi=hi)
            assert fc.first_byte_position == lo
            got += gpu_findings(fc)
            assert gs.consumed_bytes == (before if hi != len(buf) else before + len(buf))
        assert got == exp, (label, cuts)
        check_state(gs, os_)


@pytest.mark.parametrize("label,n,q,ubf,kind", RANGE_CASES[:6] + RANGE_CASES[8:10] + RANGE_CASES[12:])
def test_range_with_unknown_prefix(label, n, q, ubf, kind):
    """A rank of a range-sharded job only holds [range start - halo, range end): the carry into the range is found by the
    walk back inside the halo (SX_RANGE_PREFIX_UNKNOWN), positions come from the state's counter_offset."""
    m = M.Mission.for_label(label, n, ubf=ubf, output_line_char_nb_max=q)
    size = (2 << 20) + 4096
    buf = make_buf(kind, size, m.encoding_id, n, q, 99 + n + q)
    lo, hi = 4096 * 200, 4096 * 390
    whole = sx.ScannerState(m)
    exp = gpu_findings(whole.scan_stream(buf, False, 4096, lo=lo, hi=hi))
    for halo in (4096 * 64, 4096 * 8):
        m2 = dataclasses.replace(m, counter_offset=lo - halo)
        gs = sx.ScannerState(m2)
        part = buf[lo - halo:hi]
        got = gpu_findings(gs.scan_stream(part, False, 4096, lo=halo, hi=len(part), prefix_unknown=True))
        assert got == exp, (label, halo)


def test_range_argument_checks():
    m = M.Mission.for_label("utf-8", 6)
    gs = sx.ScannerState(m)
    buf = bytes(20000)
    for lo, hi in ((100, 4096), (0, 5000), (8192, 4096)):
        with pytest.raises(sx.ScannerError):
            gs.scan_stream(buf, False, 4096, lo=lo, hi=hi)
    with pytest.raises(sx.ScannerError):
        gs.scan_stream(buf, False, 4096, lo=0, hi=4096, prefix_unknown=True)
    assert len(gs.scan_stream(buf, False, 4096, lo=4096, hi=4096).v) == 0


def test_offsets_beyond_4gib_vs_oracle():
    """One device-resident call of 5 GiB + 12 KiB: stream offsets, window indices and text offsets beyond 2^32
    (finding_collection.rs:260: `position` is u64).  Ranges below, across and above the 4 GiB line are regenerated on
    the host and compared with the oracle finding by finding; the ScannerState counts every byte."""
    import numpy as np
    import torch

    size = (5 << 30) + 4096 * 3
    if torch.cuda.mem_get_info(0)[0] < size + (4 << 30):
        pytest.skip("needs ~10 GB of free device memory")
    seed, n = 12, 8
    m = M.Mission.for_label("utf-8", n)
    t = torch.empty(size, dtype=torch.uint8, device="cuda:0")
    L = sx.load_library()
    assert L.sx_fill_random(t.data_ptr(), size, seed, 0, 0, None) == 0
    planted = {}
    for off in ((1 << 32) - 20, (1 << 32) + 4096 * 7 - 3, (9 << 29) + 5, size - 60):  # straddling 2^32, a slice boundary, the tail
        s = b"\x00planted across the line: \xc3\xa4\xc3\xb6\xc3\xbc 0123456789\x00"
        planted[off] = s
        t[off:off + len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    gs = sx.ScannerState(m)
    raw = gs.scan_stream(None, False, 4096, device_ptr=t.data_ptr(), length=size, raw=True)
    assert gs.consumed_bytes == size and gs.last_stats.windows_total == (size + 127) // 128
    halo = 1 << 20
    for a, b in ((0, 8 << 20), ((1 << 32) - (8 << 20), (1 << 32) + (8 << 20)), (9 << 29, (9 << 29) + (8 << 20)), (size - (8 << 20) - 4096 * 3, size)):
        data = corpus.sx_mix_bytes(seed, a - min(a, halo), b - a + min(a, halo))
        for off, s in planted.items():
            lo, hi = max(off, a - min(a, halo)), min(off + len(s), b)
            if lo < hi:
                data[lo - (a - min(a, halo)):hi - (a - min(a, halo))] = np.frombuffer(s[lo - off:hi - off], dtype=np.uint8)
        mo = dataclasses.replace(m, counter_offset=a - min(a, halo))
        exp = [f for f in oracle_findings(oracle_state(mo).scan_stream(data, False, 4096)) if a <= f[0] < b]
        got = raw.select(a, b)
        assert got == exp, (a, b)
        assert len(exp) > 1000
    assert any(p >= (1 << 32) and b"planted across" in s for p, _, s, _ in raw.select(1 << 32, (1 << 32) + (1 << 20)))
    raw.close()


def test_capacity_overflow_reruns_with_counted_sizes():
    """A first call of a state has no history: entry arrays sized for 1/16 of the windows and a pinned set for one million
    findings.  koi8-r over 40 MiB of random bytes lists every window (327 680) and prints 1.1 million findings: both
    capacities overflow, the device reports the counted sizes, the call reruns -- same findings as the oracle; the second
    call of the state is sized from the first and needs no rerun."""
    m = M.Mission.for_label("koi8-r", 6)
    buf = corpus.sx_mix_bytes(44, 0, 40 << 20)
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    raw = gs.scan_stream(buf, False, 4096, raw=True)
    assert gs.last_stats.relaunches >= 1 and gs.last_stats.windows_listed > 300000
    exp = oracle_findings(os_.scan_stream(buf, False, 4096))
    assert len(raw) == len(exp) > 1000000
    assert raw.all() == exp
    check_state(gs, os_)
    raw.close()
    raw = gs.scan_stream(buf, False, 4096, raw=True)
    assert gs.last_stats.relaunches == 0
    exp2 = oracle_findings(os_.scan_stream(buf, False, 4096))
    assert raw.all() == exp2
    check_state(gs, os_)
