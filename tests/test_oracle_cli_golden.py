"""The oracle, driven like main.rs::run(), must reproduce the reference's 3 CLI golden files
byte for byte (tests/functional/run-tests:11-41 -> tests/golden/expected_output{1,2,3})."""
import os

import pytest

from helpers import M, O, to_oracle

LABELS = ["UTF-8", "utf-16le", "utf-16be"]
CASES = [
    # (n, q, grep, af, ubf, files, expected)
    (None, 16, 63, M.AF_ALL & ~M.AF_CTRL, M.UBF_COMMON, ["input1"], "expected_output1"),
    (10, 32, 58, M.AF_ALL & ~M.AF_CTRL, M.UBF_COMMON, ["input1", "input2"], "expected_output2"),
    (None, 32, None, M.AF_NONE, M.UBF_NONE, ["input1", "input2"], "expected_output3"),
]


@pytest.mark.parametrize("n,q,grep,af,ubf,files,expected", CASES, ids=["golden1", "golden2", "golden3"])
def test_cli_golden(golden_dir, n, q, grep, af, ubf, files, expected):
    missions = [
        M.Mission.for_label(lbl, n, af, ubf, grep, q, mission_id=i) for i, lbl in enumerate(LABELS)
    ]
    data = [open(os.path.join(golden_dir, f), "rb").read() for f in files]
    findings = O.cli_scan([to_oracle(m) for m in missions], data)
    out = O.print_findings(findings, [m.printed_encoding_name for m in missions], len(files), "x")
    assert out == open(os.path.join(golden_dir, expected), "rb").read()
