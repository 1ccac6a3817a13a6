"""stringsext_b200: B200-native (sm_100a CUDA) implementation of the stringsext scanner hot path
(`FindingCollection::from` over `ScannerState`, /root/reference/src/finding_collection.rs:84-342).

Layout: csrc/ holds the CUDA kernels and the C ABI (include/stringsext_b200.h); mission.py and
scanner.py mirror the reference's Mission / ScannerState / FindingCollection / Finding interface
on top of that ABI.  There is no CPU scanning path.
"""
from .mission import *  # noqa: F401,F403
from .mission import Mission, MissionError, Missions, Utf8Filter  # noqa: F401
from .input import INPUT_BUF_LEN, Slicer, scan_files, scan_inputs  # noqa: F401
from .scanner import (  # noqa: F401
    Finding,
    FindingCollection,
    PendingScan,
    Precision,
    ScannerError,
    ScannerState,
    device_count,
    load_library,
    merge,
)
