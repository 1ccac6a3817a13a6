#!/usr/bin/env python3
"""Ad-hoc differential stress: GPU scanner vs the CPU oracle on many random missions / corpora / call splits.
(Test infrastructure; the fixed-seed subset lives in tests/test_gpu_parity.py.)  usage: gpu_vs_oracle.py SEED ITERS"""
import dataclasses
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus  # noqa: E402
import stringsext_b200 as sx  # noqa: E402
from helpers import M, oracle_state  # noqa: E402

seed, iters = int(sys.argv[1]), int(sys.argv[2])
rng = random.Random(seed)
t0 = time.time()
nf = 0
for it in range(iters):
    if it and it % 200 == 0:
        print("...", it, "iterations ok", flush=True)
    enc = rng.choice([0, 1, 1, 1, 2, 3, 4, 4, 5, 6])
    general = rng.random() < float(os.environ.get("SX_STRESS_GENERAL", "0.15"))  # --grep-char / --same-unicode-block share
    m = corpus.random_general_mission(rng, enc, M) if general else corpus.random_mission(rng, enc, M)
    if rng.random() < 0.7:
        q = rng.choice([64, 64, 64, 32, 16, 8])
        n = min(m.chars_min_nb, q) if not general or rng.random() < 0.8 else q + rng.choice([1, 4, 20])  # n > q: general too
        m = dataclasses.replace(m, output_line_char_nb_max=q, chars_min_nb=n)
    size = rng.choice([0, 1, 777, 4096, 70000, 300000, 1 << 20, (2 << 20) + 13])
    kind = rng.choice(["rand", "rand", "mixed", "lowent", "runs", "text", "planted"])
    if kind == "planted":
        buf = corpus.sx_mix_bytes(it + seed * 1000, 0, max(size, 64))
        corpus.plant(buf, it, m.encoding_id, m.chars_min_nb, m.output_line_char_nb_max, density=1 << 12)
        buf = buf.tobytes()
    else:
        buf = corpus.gen(rng, kind, size, enc)
    slice_len = rng.choice([4096, 4096, 4096, 8192, 1024])
    gs, os_ = sx.ScannerState(m), oracle_state(m)
    gs.set_sparse(rng.choice([1, 1, 0]))
    gs.set_direct_output(rng.choice([True, True, False]))
    k = rng.choice([1, 1, 2, 3])
    cuts = [0] + sorted(rng.randrange(0, len(buf) + 1) for _ in range(k - 1)) + [len(buf)]
    for c, (a, b) in enumerate(zip(cuts, cuts[1:])):
        last = c == k - 1 and rng.random() < 0.3
        got = [(f.position, int(f.position_precision), f.s, f.s_completes_previous_s) for f in gs.scan_stream(buf[a:b], last, slice_len).v]
        exp = [(f.position, f.precision, f.s, f.completes) for f in os_.scan_stream(buf[a:b], last, slice_len).v] if b > a else []
        if got != exp or gs.last_scan_run_leftover != os_.leftover or gs.last_run_str_was_printed_and_is_maybe_cut_str != os_.cut:
            d = [(i, x, y) for i, (x, y) in enumerate(zip(got, exp)) if x != y][:2]
            print("MISMATCH seed", seed, "it", it, "enc", enc, m, "kind", kind, "size", size, "slice", slice_len, "cuts", cuts, "call", c,
                  "last", last, d, len(got), len(exp))
            sys.exit(1)
        nf += len(exp)
print("ok seed", seed, "iters", iters, "findings", nf, "in", round(time.time() - t0, 1), "s")
