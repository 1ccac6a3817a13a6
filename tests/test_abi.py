"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/stringsext_b200.h declares, and refuses loudly to scan without a CUDA device."""
import ctypes
import os
import re

import pytest

import stringsext_b200 as sx
from stringsext_b200 import build as sxbuild
from stringsext_b200 import scanner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(scanner.LIB_PATH):
        sxbuild.build()
    return scanner.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "stringsext_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, not comments
    declared = set(re.findall(r"\b(sx_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(scanner.exported_symbols())


def test_no_silent_cpu_fallback(lib):
    if sx.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sx.ScannerError) as ei:
        sx.ScannerState(sx.Mission.for_label("utf-8"))
    assert ei.value.code == 1  # SX_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "stringsext_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"(import|from|include|CDLL).*oracle", txt), f"{f} uses the oracle"
