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
