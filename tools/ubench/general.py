"""Throughput of general missions (--grep-char / --same-unicode-block) on random bytes, input resident on the device.
Usage: python tools/ubench/general.py [size_mib]"""
import dataclasses
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stringsext_b200 as sx  # noqa: E402

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
g = torch.Generator(device="cuda").manual_seed(1)
buf = torch.randint(0, 256, (size,), dtype=torch.uint8, device="cuda", generator=g)
stream = torch.cuda.current_stream()
# (label, grep, same block, prefilter, bytes scanned).  Of the general missions --grep-char alone and --same-unicode-block
# alone may use the prefilter (PrefCfg::kill_trail / sb_rule); for the others the flag has no effect.
small = min(size, 32 << 20)
CASES = (("ascii", ord("e"), False, True, small), ("koi8-r", ord("e"), False, True, small), ("utf-8", ord("e"), False, True, size),
         ("utf-8", None, True, True, size), ("ascii", ord("e"), False, True, size), ("koi8-r", None, True, True, size),
         ("utf-16le", ord("e"), True, True, size), ("utf-16le", None, True, True, size), ("ascii", None, True, True, size),
         ("utf-8", None, True, False, small), ("utf-8", None, False, True, size))
for label, grep, same, pref, nbytes in CASES:
    m = sx.Mission.for_label(label, 10)
    m = dataclasses.replace(m, filter=dataclasses.replace(m.filter, grep_char=grep), require_same_unicode_block=same)
    if True:
        size = nbytes
        ss = sx.ScannerState(m, 0)
        ss.set_prefilter(pref)
        ms, n = [], 0
        for it in range(3):
            ss.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fc = ss.scan_stream(None, False, 4096, device_ptr=buf.data_ptr(), length=size, cuda_stream=stream.cuda_stream, raw=True)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            n = len(fc)
            del fc
        print(f"{nbytes >> 20} MiB {label} -n 10 grep={grep} same_block={same} prefilter={pref}: {min(ms):.2f} ms = {size / 2**30 / (min(ms) / 1e3):.1f} GiB/s, "
              f"{n} findings, listed {ss.last_stats.windows_listed} of {ss.last_stats.windows_total}")
