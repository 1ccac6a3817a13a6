"""the reference's input geometry (input.rs:33-168) and the streaming driver on top of it."""
import io
import os
import random

import pytest

import corpus
import stringsext_b200 as sx
from helpers import M, O, to_oracle


def _write(tmp_path, name, data):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(data)
    return p


def test_slicer_geometry(tmp_path):
    """Per-file 4096-byte grid, an empty piece with the NEXT file's label at every file switch, 1-based labels, stdin
    unlabelled, is_last never true (input.rs:118-167)."""
    a = _write(tmp_path, "a", bytes(range(256)) * 40)   # 10240 = 2 * 4096 + 2048
    b = _write(tmp_path, "b", b"x" * 4096)
    got = [(len(s), fid, last) for s, fid, last in sx.Slicer([a, b])]
    assert got == [(4096, 1, False), (4096, 1, False), (2048, 1, False), (0, 2, False), (4096, 2, False)]
    got = [(len(s), fid, last) for s, fid, last in sx.Slicer(["-"], stdin=io.BytesIO(b"y" * 5000))]
    assert got == [(4096, None, False), (904, None, False)]
    missing = os.path.join(str(tmp_path), "missing")
    assert [(len(s), fid) for s, fid, _ in sx.Slicer([missing, b])] == [(0, 2), (4096, 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [4096 * 3, 1 << 20])
def test_scan_files_matches_slice_by_slice_driver(tmp_path, golden_dir, chunk):
    """scan_files (large pieces, double-buffered upload, three missions) == the reference driver fed slice by slice
    (oracle cli_scan): same findings in the same merged order; golden 2's inputs plus a random binary."""
    rng = random.Random(5)
    extra = corpus.sx_mix_bytes(77, 0, 300000 + 123)
    corpus.plant(extra, 77, 1, 10, 32, density=1 << 13)
    files = [os.path.join(golden_dir, "input1"), os.path.join(golden_dir, "input2"), _write(tmp_path, "rnd.bin", extra.tobytes())]
    missions = [M.Mission.for_label(lbl, 10, M.AF_ALL & ~M.AF_CTRL, M.UBF_COMMON, 58, 32, mission_id=i)
                for i, lbl in enumerate(["UTF-8", "utf-16le", "utf-16be"])]
    states = [sx.ScannerState(m) for m in missions]
    got = []
    for fid, fcs in sx.scan_files(states, files, chunk_bytes=chunk):
        got += [(f.position, f.mission.mission_id, int(f.position_precision), f.s, f.s_completes_previous_s, f.input_file_id)
                for f in sx.merge(fcs)]
    data = [open(p, "rb").read() for p in files]
    exp = [(f.position, f.mission_id, f.precision, f.s, f.completes, f.input_file_id)
           for f in O.cli_scan([to_oracle(m) for m in missions], data)]
    assert got == exp and len(exp) > 100


@pytest.mark.gpu
def test_streaming_driver_from_a_pipe_like_reader():
    """sx_scan_reader with a reader callback that returns short reads (a pipe): the library normalises the grid to full
    4096-byte slices, i.e. the result equals the oracle fed with 4096-byte slices of the same bytes."""
    data = corpus.sx_mix_bytes(3, 0, 200000 + 77)
    corpus.plant(data, 3, 1, 6, 64, density=1 << 12)
    data = data.tobytes()

    class Dribble(io.RawIOBase):
        def __init__(self, b):
            self.b, self.o, self.rng = b, 0, random.Random(1)

        def read(self, n=-1):
            k = min(n, self.rng.choice([1, 7, 100, 4096, 5000]), len(self.b) - self.o)
            out = self.b[self.o:self.o + k]
            self.o += k
            return out

    m = M.Mission.for_label("utf-8", 6)
    st = sx.ScannerState(m)
    got = []
    for fid, fcs in sx.scan_files([st], ["-"], chunk_bytes=4096 * 8, stdin=Dribble(data)):
        assert fid is None
        got += [(f.position, int(f.position_precision), f.s, f.s_completes_previous_s) for f in fcs[0].v]
    os_ = O.OState(to_oracle(m))
    exp = [(f.position, f.precision, f.s, f.completes) for f in os_.scan_stream(data, False, 4096).v]
    assert got == exp and len(exp) > 10


@pytest.mark.gpu
def test_async_calls_overlap_missions_and_keep_call_order():
    """sx_scan_stream_async: two missions issued before either is waited for; two chained calls on one state issued back
    to back (the worker runs them in call order, the second sees the ScannerState the first left)."""
    buf = corpus.sx_mix_bytes(21, 0, (1 << 20) + 4096 * 3 + 5)
    corpus.plant(buf, 21, 1, 4, 64, density=1 << 12)
    corpus.plant(buf, 21, 2, 4, 64, density=1 << 12)
    buf = buf.tobytes()
    cut = 4096 * 100 + 3
    ms = [M.Mission.for_label("utf-8", 4, ubf=M.UBF_ALL_VALID, mission_id=0), M.Mission.for_label("utf-16le", 4, ubf=M.UBF_ALL_VALID, mission_id=1)]
    states = [sx.ScannerState(m) for m in ms]
    pend = [[s.scan_stream_async(buf[:cut], False, 4096), s.scan_stream_async(buf[cut:], False, 4096)] for s in states]
    for s, m, (p1, p2) in zip(states, ms, pend):
        os_ = O.OState(to_oracle(m))
        for p, part in ((p1, buf[:cut]), (p2, buf[cut:])):
            got = [(f.position, int(f.position_precision), f.s, f.s_completes_previous_s) for f in p.wait().v]
            exp = [(f.position, f.precision, f.s, f.completes) for f in os_.scan_stream(part, False, 4096).v]
            assert got == exp
        assert s.consumed_bytes == os_.consumed_bytes and s.last_scan_run_leftover == os_.leftover
    # errors travel with the handle
    with pytest.raises(sx.ScannerError):
        states[0].scan_stream_async(buf, False, 4096, lo=100, hi=4096).wait()


@pytest.mark.gpu
def test_two_devices_one_process(tmp_path, golden_dir):
    """Two states on two devices driven from one process (how a Rust main.rs would bind): the streaming
    driver uploads every piece to both devices, the scans run side by side, sx_merge gives the reference's order."""
    if sx.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    files = [os.path.join(golden_dir, "input1"), os.path.join(golden_dir, "input2")]
    missions = [M.Mission.for_label(lbl, 10, M.AF_ALL & ~M.AF_CTRL, M.UBF_COMMON, None, 32, mission_id=i)
                for i, lbl in enumerate(["UTF-8", "utf-16le"])]
    states = [sx.ScannerState(m, device=i) for i, m in enumerate(missions)]
    got = []
    for fid, fcs in sx.scan_files(states, files, chunk_bytes=4096 * 16):
        got += [(f.position, f.mission.mission_id, int(f.position_precision), f.s, f.s_completes_previous_s, f.input_file_id)
                for f in sx.merge(fcs)]
    data = [open(p, "rb").read() for p in files]
    exp = [(f.position, f.mission_id, f.precision, f.s, f.completes, f.input_file_id)
           for f in O.cli_scan([to_oracle(m) for m in missions], data)]
    assert got == exp and len(exp) > 100
