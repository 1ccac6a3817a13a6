#!/usr/bin/env python3
"""bench.py -- scanned GiB/s of the stringsext scanner hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (sx_scan_range / sx_scan_stream: prefilter, exact stage, findings and their text
on the host) over the configured synthetic stream for every mission of the configuration.  N=1 runs BASELINE.json
configs[1] (`-e utf-8 -n 10` over a 4 GiB random buffer); N>1 runs configs[2..4] as written (N encodings over one
stream).  One process per GPU, no collective on the data path.

Sharding:
  --shard range   (default) the stream is cut into N slice-aligned ranges; rank r holds range r (+ a 1 MiB halo in
                  front) and scans it for EVERY mission with sx_scan_range.  Same total work as one encoding per GPU,
                  but balanced: an output-heavy mission (koi8-r on random bytes) leaves over N PCIe links.
  --shard mission one encoding per GPU over the whole stream, the reference's decomposition (main.rs:151-167).

  value     whole-job GiB/s = missions x stream bytes / step time, input resident in HBM (max over ranks)
  e2e       the same metric with the rank's bytes in pinned HOST memory: H2D copy, scans, findings on the host
  roofline  the dominant kernel of rank 0's slowest mission: algorithmic bytes / summed launch duration (CUDA events
            recorded by the library on the launching stream) vs MEASURED_PEAKS.json hbm_gbs
  parity    after the timed loops: sample ranges of the benchmarked stream are regenerated on the host, scanned by the
            CPU oracle and compared finding by finding with what the GPU returned for the same ranges
  cpu_baseline  the CPU oracle (a port of the reference algorithm; the Rust reference cannot be built here) on a
            bounded sample of the same stream, one scanning thread per mission exactly like the reference
            (main.rs:151-167); cpu_baseline_all_cores: the same port range-sharded over every host core

`--impl reference` times that CPU port as the reference arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = float(1 << 30)
MIB = 1 << 20
SLICE = 4096
HALO = 1 << 20

# BASELINE.json configs -> missions (label, ubf name), stream size, chars_min_nb, seed
CONFIGS = {
    1: dict(name="-e utf-8 -n 10 over 4 GiB random buffer, 1xB200", size=4 << 30, n=10, seed=2,
            missions=[("utf-8", None)]),
    2: dict(name="-e utf-16le -e utf-16be -n 10 -u African over 4 GiB, 2xB200 (one encoding per GPU)", size=4 << 30,
            n=10, seed=3, missions=[("utf-16le", "African"), ("utf-16be", "African")], balanced=True),
    4: dict(name="-e utf-8 -e utf-16le -e utf-16be -e big5 -n 8 over 16 GiB, 4xB200", size=16 << 30, n=8, seed=4,
            missions=[("utf-8", None), ("utf-16le", None), ("utf-16be", None), ("big5", None)]),
    8: dict(name="8 encodings (ascii, utf-8, utf-16le, utf-16be, utf-32le, utf-32be, euc-jp, koi8-r) -n 6 over 32 GiB, 8xB200",
            size=32 << 30, n=6, seed=5,
            missions=[("ascii", None), ("utf-8", None), ("utf-16le", None), ("utf-16be", None), ("utf-32le", None),
                      ("utf-32be", None), ("euc-jp", None), ("koi8-r", None)]),
}


def config_for(n_gpus: int, config_id: int = 0):
    if config_id:
        return CONFIGS[config_id]
    if n_gpus in CONFIGS:
        return CONFIGS[n_gpus]
    return CONFIGS[8]


def make_mission(sx, label, ubf_name, n, mission_id=0, counter_offset=0):
    ubf = sx.UBF_AFRICAN if ubf_name == "African" else None
    return sx.Mission.for_label(label, n, ubf=ubf, mission_id=mission_id, counter_offset=counter_offset)


def plant_patches(seed, enc_id, n, q, length, per_mib=1):
    """(offset, bytes) patches = the planted corpus (tests/corpus.py) of one encoding, applied in order."""
    import random

    import corpus

    rng = random.Random(seed * 7919 + enc_id)
    strings = corpus.planted_strings(rng, enc_id, n, q)
    count = max(12, int(length // (1 << 20)) * per_mib)
    out = []
    W, slice_len = 2 * q, SLICE
    for i in range(count):
        s = strings[i % len(strings)]
        base = rng.randrange(0, max(1, length - len(s) - 8))
        kind = i % 4
        if kind == 1:
            base = (base // W) * W + W - rng.randrange(1, max(2, min(len(s), W)))
        elif kind == 2:
            base = (base // slice_len) * slice_len + slice_len - rng.randrange(1, max(2, min(len(s), slice_len)))
        elif kind == 3:
            base |= 1
        base = max(0, min(base, length - len(s)))
        out.append((base, s))
    return out


def all_patches(cfg, M, size):
    """The planted corpus of the stream: strings in every mission's encoding (the stream is the same for all ranks)."""
    out, seen = [], set()
    for label, ubf in cfg["missions"]:
        enc = make_mission(M, label, ubf, cfg["n"]).encoding_id
        if enc in seen:
            continue
        seen.add(enc)
        out += plant_patches(cfg["seed"], enc, cfg["n"], 64, size)
    return out


class PatchIndex:
    """Patches sorted by offset, for range queries."""

    def __init__(self, patches):
        import bisect

        self.bisect = bisect
        self.order = sorted(range(len(patches)), key=lambda i: patches[i][0])
        self.offs = [patches[i][0] for i in self.order]
        self.patches = patches
        self.maxlen = max((len(s) for _, s in patches), default=0)

    def overlapping(self, start, length):
        """Patches that touch [start, start + length), in APPLICATION order (later patches overwrite earlier ones)."""
        a = self.bisect.bisect_left(self.offs, start - self.maxlen)
        b = self.bisect.bisect_left(self.offs, start + length)
        idx = sorted(self.order[a:b])
        return [self.patches[i] for i in idx if self.patches[i][0] + len(self.patches[i][1]) > start]


def host_range(seed, pidx, start, length):
    """Regenerate bytes [start, start+length) of the benchmark stream on the host."""
    import corpus
    import numpy as np

    buf = corpus.sx_mix_bytes(seed, start, length)
    for off, s in pidx.overlapping(start, length):
        a, b = max(off, start), min(off + len(s), start + length)
        if a < b:
            buf[a - start : b - start] = np.frombuffer(s[a - off : b - off], dtype=np.uint8)
    return buf


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU oracle legs (the only places bench.py touches oracle/) -----------------------------------------------------
def oracle_scan(mission_o, sample, keep=False):
    """One oracle pass of one mission over `sample` (numpy uint8); returns (n findings, findings or None)."""
    import ctypes as C

    from oracle import oracle as O

    ss = O.OState(mission_o)
    if keep:
        fc = ss.scan_stream(sample, False, SLICE)
        return len(fc.v), [(f.position, f.precision, f.s, f.completes) for f in fc.v]
    L = O.lib()
    fc = L.sxo_scan_stream(ss._h, -1, C.c_void_p(sample.ctypes.data), sample.size, SLICE, 0)
    n = L.sxo_fc_len(fc)
    L.sxo_fc_free(fc)
    return n, None


def cpu_ref_shape(missions_o, sample):
    """The reference's decomposition: one scanning thread per mission over the same bytes (main.rs:151-167)."""
    t0 = time.perf_counter()
    if len(missions_o) == 1:
        oracle_scan(missions_o[0], sample)
    else:
        ths = [threading.Thread(target=oracle_scan, args=(m, sample)) for m in missions_o]
        [t.start() for t in ths]
        [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return len(missions_o) * sample.size / GIB / dt, dt


def cpu_all_cores(missions_o, sample, cores):
    """The same port, range-sharded: every (mission, chunk) is an independent scan that starts 64 KiB early (the carry
    dies within a few windows on this input); threads = host cores.  A throughput figure, findings are not collected."""
    pre = 64 << 10
    nchunks = max(1, cores // max(1, len(missions_o)))
    per = max(SLICE, (sample.size // nchunks // SLICE) * SLICE)
    jobs = []
    for m in missions_o:
        for k in range(nchunks):
            lo = k * per
            hi = sample.size if k == nchunks - 1 else lo + per
            if lo >= sample.size:
                break
            s0 = max(0, lo - pre)
            jobs.append((dataclasses.replace(m, counter_offset=s0), sample[s0:hi]))
    t0 = time.perf_counter()
    ths = [threading.Thread(target=oracle_scan, args=j) for j in jobs]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return len(missions_o) * sample.size / GIB / dt, dt, len(jobs)


def run_reference_arm(args):
    """`--impl reference`: the CPU port of the reference algorithm, run the way the reference runs (one scanning thread
    per mission), on a bounded sample of the same workload per step; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import stringsext_b200.mission as M
    from helpers import to_oracle

    cfg = config_for(args.gpus, args.config)
    size = cfg["size"] if not args.size_mib else args.size_mib << 20
    missions = [make_mission(M, lbl, u, cfg["n"], i) for i, (lbl, u) in enumerate(cfg["missions"])]
    mo = [to_oracle(m) for m in missions]
    pidx = PatchIndex(all_patches(cfg, M, size))
    # size the per-step sample so that warm-up + steps finish in about --ref-seconds: calibrate on 8 MiB
    cal = host_range(cfg["seed"], pidx, 0, 8 * MIB)
    rate, _ = cpu_ref_shape(mo, cal)  # GiB/s over all missions
    passes = args.warmup + args.steps
    want = rate / len(mo) * GIB * args.ref_seconds / max(1, passes)
    sample_bytes = int(max(8 * MIB, min(size, want)) // MIB) * MIB
    sample = host_range(cfg["seed"], pidx, 0, sample_bytes)
    times = []
    for i in range(passes):
        _, dt = cpu_ref_shape(mo, sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = len(mo) * sample.size / GIB / (ms / 1e3)
    full = sample_bytes == size
    line = {
        "impl": "reference", "metric": "scanned GiB/s (whole job, all encodings)", "value": val, "unit": "GiB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["name"], "missions": [l for l, _ in cfg["missions"]], "sample_bytes_per_step": sample.size,
                   "stream_bytes": size, "full_stream_per_step": full,
                   "note": "CPU port of the reference algorithm (oracle/); the Rust reference cannot be built in this image. "
                           + ("Every step scans the whole configured stream." if full else
                              f"Every step scans the first {sample_bytes >> 20} MiB of the configured stream (a rate: the data is "
                              f"uniform); the full stream would take {size / GIB / (val / len(mo)):.0f} s per pass, "
                              "the declared warm-up and steps are the ones that were run.")},
        "cpu_baseline": {"value": val, "unit": "GiB/s", "cores": len(mo), "kind": "port",
                         "sample": f"first {sample_bytes >> 20} MiB of the workload stream per step, one scanning thread per mission "
                                   f"(the reference's decomposition, main.rs:151-167), {os.cpu_count()} host cores available"},
        "e2e": {"value": val, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def parity_ranges(lo, hi, n_gpus, n_missions, first_mib):
    """Sample ranges [a, b) of this rank's part of the stream that the oracle re-scans: the start (which at N=1 is the
    range the cpu_baseline pass covers anyway), points inside and the tail."""
    span = hi - lo
    if n_gpus == 1 and n_missions == 1:
        sizes = [first_mib * MIB, 64 * MIB, 64 * MIB, 64 * MIB]
        starts = [lo, lo + span // 4, lo + span * 5 // 8, hi]
    else:
        sizes = [16 * MIB, 16 * MIB, 16 * MIB]
        starts = [lo, lo + span // 2, hi]
    out = []
    for a, s in zip(starts, sizes):
        s = min(span, s) // SLICE * SLICE
        a = max(lo, min(a // SLICE * SLICE, hi - s))
        if s > 0 and (a, a + s) not in out:
            out.append((a, a + s))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--shard", default="auto", choices=["auto", "range", "mission"],
                    help="auto: by range, except for a config whose missions cost the same (configs[2]: utf-16le / utf-16be), where one "
                         "encoding per GPU avoids paying the exact stage's fixed latency twice per GPU")
    ap.add_argument("--config", type=int, default=0, help="BASELINE config to run (1, 2, 4, 8) instead of the one --gpus names; "
                                                          "e.g. --config 1 --gpus 8 = one mission range-sharded over 8 GPUs")
    ap.add_argument("--size-mib", type=int, default=0, help="override the stream size (debug only; invalidates the number)")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--ref-seconds", type=float, default=120.0, help="--impl reference: wall-clock budget of the whole run")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sparse-mode", type=int, default=1, help="debug: 0 block kernel only, 1 default")
    ap.add_argument("--as-rank", type=int, default=-1, help="debug: a single process doing the work of rank R of a --gpus N job")
    ap.add_argument("--mission-threads", type=int, default=1, help="1: one host thread + stream per mission of a rank (default); 0: the rank's missions one after the other")
    ap.add_argument("--only", default="", help="debug: comma-separated mission indices of the config to run (others are skipped)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import stringsext_b200 as sx
    from helpers import to_oracle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    lay_world, lay_rank = world, rank  # the job layout this process takes its part from
    if args.as_rank >= 0:
        assert world == 1
        lay_world, lay_rank = args.gpus, args.as_rank
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = config_for(args.gpus, args.config)
    size = cfg["size"] if not args.size_mib else args.size_mib << 20
    n_miss = len(cfg["missions"])
    L = sx.load_library()

    # ---- this rank's part of the job -------------------------------------------------------------------
    shard = args.shard if lay_world > 1 else "range"
    if shard == "auto":
        shard = "mission" if (cfg.get("balanced") and lay_world == n_miss) else "range"
    if shard == "mission" and lay_world != n_miss:
        shard = "range"
    if shard == "range":
        from stringsext_b200.multi import range_plan

        base, lo, hi = range_plan(size, lay_world, SLICE, HALO)[lay_rank]  # base: stream offset of the first byte this rank holds
        my_missions = list(range(n_miss))
    else:
        base, lo, hi = 0, 0, size
        my_missions = [lay_rank]
    if args.only:
        my_missions = [i for i in my_missions if i in {int(x) for x in args.only.split(",")}]
        if not my_missions:
            sys.exit("--only selects none of this rank's missions (mission sharding gives rank r mission r)")
    blen = hi - base
    missions = [make_mission(sx, cfg["missions"][i][0], cfg["missions"][i][1], cfg["n"], i, counter_offset=base) for i in my_missions]
    labels = [cfg["missions"][i][0] for i in my_missions]
    pidx = PatchIndex(all_patches(cfg, sx, size))

    # ---- synthetic input, resident in HBM ----------------------------------------------------------------
    dbuf = torch.empty(blen, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    assert L.sx_fill_random(dbuf.data_ptr(), blen, cfg["seed"], base, local, None) == 0
    for off, s in pidx.overlapping(base, blen):
        a, b = max(off, base), min(off + len(s), hi)
        if a < b:
            dbuf[a - base : b - base] = torch.frombuffer(bytearray(s[a - off : b - off]), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()

    states = [sx.ScannerState(m, local) for m in missions]
    for st in states:
        st.set_sparse(args.sparse_mode)
    last_fc = [None] * len(states)

    # The missions of a rank are independent (main.rs:155-166 gives each its own thread and ScannerState): one host
    # thread and one CUDA stream per mission, so the latency-bound exact stage of one mission runs beside the streaming
    # prefilter of another.
    from concurrent.futures import ThreadPoolExecutor

    mstreams = [torch.cuda.Stream() for _ in states]
    pool = ThreadPoolExecutor(max_workers=len(states)) if len(states) > 1 else None

    def scan_one(k, ptr):
        st = states[k]
        st.reset()  # ScannerState::new semantics per pass (same work every step), device buffers are reused
        if last_fc[k] is not None:
            last_fc[k].close()
        last_fc[k] = st.scan_stream(None, False, SLICE, device_ptr=ptr, length=blen, cuda_stream=mstreams[k].cuda_stream, raw=True,
                                    lo=lo - base, hi=blen, prefix_unknown=base > 0)
        return st.last_stats

    def scan_all(ptr):
        """One step: every mission of this rank over its range of the stream (device pointer `ptr`)."""
        for ms in mstreams:
            ms.wait_stream(stream)  # e.g. the H2D copy of the e2e leg
        if pool is None or args.mission_threads == 0:
            return [scan_one(k, ptr) for k in range(len(states))]
        return list(pool.map(lambda k: scan_one(k, ptr), range(len(states))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        scan_all(dbuf.data_ptr())
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_mission = [dict(pre=[], lst=[], ex=[], mat=[], scan=[], sp=[], host=[], call=[]) for _ in states]
    launches, d2h_step = 0, 0
    e0.record(stream)
    for _ in range(args.steps):
        stats = scan_all(dbuf.data_ptr())
        d2h_step = 0
        for k, st in enumerate(stats):
            pm = per_mission[k]
            pm["pre"].append(st.prefilter_kernel_ms)
            pm["lst"].append(st.list_kernels_ms)
            pm["ex"].append(st.exact_kernel_ms)
            pm["mat"].append(st.materialize_kernel_ms)
            pm["scan"].append(st.scan_kernel_ms)
            pm["sp"].append([float(x) for x in st.sparse_stage_ms])
            pm["host"].append([float(x) for x in st.host_phase_ms])
            pm["call"].append(st.host_total_ms)
            pm["sparse_used"], pm["pieces"] = int(st.sparse_used), int(st.pieces)
            pm["win_total"], pm["win_listed"] = int(st.windows_total), int(st.windows_listed)
            launches += st.kernel_launches
            d2h_step += st.d2h_bytes
    e1.record(stream)
    barrier()
    clock_note = None
    if len(sampler.lines) < 3:
        # the timed region is shorter than a few of nvidia-smi's 20 ms sampling periods: keep the same load up (untimed)
        # until the sampler has seen it
        t_probe = time.perf_counter()
        while len(sampler.lines) < 4 and time.perf_counter() - t_probe < 0.5:
            scan_all(dbuf.data_ptr())
        torch.cuda.synchronize()
        clock_note = "timed region shorter than nvidia-smi's 20 ms sampling period: sampled over identical untimed steps run right after it"
    clocks = sampler.stop()
    if clock_note:
        clocks["note"] = clock_note
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    my_ms = total_ms / args.steps
    ms_per_step = float(t.item()) / args.steps
    value = n_miss * size / GIB / (ms_per_step / 1e3)
    if args.as_rank >= 0:  # debug: this process did 1 / lay_world of the job
        value = len(states) * (hi - lo) / GIB / (ms_per_step / 1e3)
    nfind = [len(fc) for fc in last_fc]

    # ---- parity: the oracle over sample ranges of the benchmarked stream ------------------------------------
    parity = None
    cpu_first = None  # (seconds, bytes) of the oracle pass over the first range at N=1 (doubles as cpu_baseline)
    if not args.no_parity:
        ranges = parity_ranges(lo, hi, lay_world, n_miss, args.cpu_sample_mib)
        jobs = []
        for k in range(len(missions)):
            for (a, b) in ranges:
                pre = min(a, HALO) // SLICE * SLICE
                jobs.append((k, a, b, pre))
        results = [None] * len(jobs)

        def work(j):
            k, a, b, pre = jobs[j]
            end = min(size, b + SLICE)
            data = host_range(cfg["seed"], pidx, a - pre, end - (a - pre))
            mo = dataclasses.replace(to_oracle(missions[k]), counter_offset=a - pre)
            t0 = time.perf_counter()
            _, fo = oracle_scan(mo, data, keep=True)
            dt = time.perf_counter() - t0
            exp = [f for f in fo if a <= f[0] < b]
            got = last_fc[k].select(a, b)
            bad = None
            if got != exp:
                for i in range(max(len(got), len(exp))):
                    g = got[i] if i < len(got) else None
                    e = exp[i] if i < len(exp) else None
                    if g != e:
                        bad = {"mission": labels[k], "range": [a, b], "index": i, "gpu": repr(g)[:200], "oracle": repr(e)[:200]}
                        break
            results[j] = (len(exp), 0 if got == exp else 1, bad, dt, end - (a - pre))

        nthreads = max(1, min(len(jobs), (os.cpu_count() or 4) // max(1, min(world, 8)), 8))
        if world == 1 and n_miss == 1:
            # the first range doubles as the single-thread cpu_baseline measurement: run it alone, the others afterwards
            work(0)
            cpu_first = (results[0][3], results[0][4])
            pending = list(range(1, len(jobs)))
        else:
            pending = list(range(len(jobs)))
        it = iter(pending)
        lock = threading.Lock()

        def runner():
            while True:
                with lock:
                    j = next(it, None)
                if j is None:
                    return
                work(j)

        ths = [threading.Thread(target=runner) for _ in range(nthreads)]
        [t_.start() for t_ in ths]
        [t_.join() for t_ in ths]
        par = {"ranges": [[a, b] for a, b in ranges], "missions": labels, "findings_compared": sum(r[0] for r in results),
               "mismatches": sum(r[1] for r in results), "examples": [r[2] for r in results if r[2]][:3]}
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, par)
            parity = {"ranges_per_rank": [p["ranges"] for p in allp], "missions": labels,
                      "findings_compared": sum(p["findings_compared"] for p in allp),
                      "mismatches": sum(p["mismatches"] for p in allp), "examples": sum((p["examples"] for p in allp), [])[:3]}
        else:
            parity = par
        parity["how"] = ("the host regenerates each range (+1 MiB in front for the carry), the CPU oracle scans it, the findings with "
                         "position inside the range are compared with the GPU's from the last timed step: position, precision, "
                         "text, completes flag")

    # ---- e2e: the rank's bytes in pinned host memory -> H2D -> every mission -> findings on the host -------------
    e2e = None
    if not args.no_e2e:
        hbuf = torch.empty(blen, dtype=torch.uint8, pin_memory=True)
        hbuf.copy_(dbuf)
        torch.cuda.synchronize()

        def step_host():
            dbuf.copy_(hbuf, non_blocking=True)  # H2D inside the timed region
            return scan_all(dbuf.data_ptr())     # the library's streams wait for the caller's stream (the copy)

        step_host()
        barrier()
        e2e_steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            st_e = step_host()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": n_miss * size / GIB / float(t.item()), "unit": "GiB/s", "h2d_bytes_per_step": int(blen),
               "d2h_bytes_per_step": int(sum(s.d2h_bytes for s in st_e)), "bytes_per_gpu": int(blen), "ms_per_step_rank0": dt * 1e3,
               "findings_rank0": [len(fc) for fc in last_fc],
               "config": "each rank uploads its whole part of the stream once per step (pinned host -> device) and scans it for "
                         "every mission; findings + text land in pinned host memory"}

    gathered = None
    if world > 1:
        info = {"rank": rank, "ms": my_ms, "range": [lo, hi], "missions": labels, "findings": nfind,
                "mission_ms": [sum(pm["call"]) / len(pm["call"]) for pm in per_mission]}
        gathered = [None] * world
        dist.all_gather_object(gathered, info)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        avg = lambda v: sum(v) / len(v)
        # the mission that takes longest on this rank, and its kernels
        slow = max(range(len(states)), key=lambda k: avg(per_mission[k]["call"]))
        pm = per_mission[slow]
        kernels = {"sx_prefilter_kernel": avg(pm["pre"]), "sx_materialize_kernel": avg(pm["mat"])}
        if pm["sparse_used"]:
            names = ("sx_list_compact_kernel", "sx_sp_heads_kernel", "sx_sp_members_kernel", "sx_sp_fix+late_kernel", "sx_sp_ext_kernel",
                     "sx_sp_scan+gather_kernels")
            for i, nm in enumerate(names):
                kernels[nm] = avg([v[i] for v in pm["sp"]])
            kernels["exact_stage_total"] = avg(pm["ex"])
            candidates = ("sx_prefilter_kernel",) + names
        else:
            kernels["sx_list_offsets_kernel"] = avg(pm["lst"])
            kernels["sx_exact_kernel"] = avg(pm["ex"])
            candidates = ("sx_prefilter_kernel", "sx_exact_kernel")
        dominant = max(candidates, key=lambda k: kernels[k])
        k_ms = kernels[dominant]
        # algorithmic bytes per step of that kernel: the prefilter reads every input byte of the range
        # once (summed over its launches when the call is cut into pieces); the exact stage reads the listed windows
        alg_bytes = (hi - lo) if dominant == "sx_prefilter_kernel" else pm["win_listed"] * 2 * 64
        achieved = alg_bytes / 1e9 / (k_ms / 1e3)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if tr.get("workload") == cfg["name"] and not args.size_mib and world == 1 and dominant in tr:
                traffic = tr[dominant]["dram_bytes_read"] + tr[dominant]["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": "scanned GiB/s (whole job, all encodings)", "value": value, "unit": "GiB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["name"], "stream_bytes": size, "missions": [l for l, _ in cfg["missions"]],
                       "sharding": ("range: rank r scans slice-aligned range r of the stream (+1 MiB halo) for every mission, sx_scan_range"
                                    if shard == "range" else "mission: one encoding per GPU over the whole stream"),
                       "bytes_resident_per_gpu": blen, "chars_min_nb": cfg["n"], "slice_len": SLICE, "output_line_char_nb_max": 64,
                       "l2": "per-GPU input (>= 2 GiB) is larger than L2; no flush needed",
                       "input": "splitmix64 counter-based random bytes + 1 planted string per MiB and encoding",
                       "findings_rank0": dict(zip(labels, nfind)), "substituted": {}, "size_overridden": bool(args.size_mib)},
            "clocks": clocks,
            "gpu_launches": launches,
            "gpu_launches_note": "library kernels launched by rank 0 inside the timed region",
            "roofline": {"bound": "hbm", "kernel": dominant, "mission": labels[slow], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "peak_note": "the measured peak is a device copy (read + write); a read-only pass can slightly exceed it",
                         "algorithmic_bytes_per_step": alg_bytes, "kernel_ms": k_ms, "kernels_ms": kernels, "pieces": pm["pieces"],
                         "pipeline_ms": avg(pm["scan"]), "pipeline_gbs": (hi - lo) / 1e9 / (avg(pm["scan"]) / 1e3),
                         "whole_step_frac_of_peak": (hi - lo) * len(states) / 1e9 / (my_ms / 1e3) / peak,
                         "windows_total": pm["win_total"], "windows_listed": pm["win_listed"],
                         "host_call_ms": avg(pm["call"]), "host_phase_ms": [avg([h[k] for h in pm["host"]]) for k in range(4)],
                         "mission_ms_rank0": {labels[k]: avg(per_mission[k]["call"]) for k in range(len(states))}},
        }
        if e2e:
            line["e2e"] = e2e
        if parity:
            line["parity"] = parity
        if gathered:
            line["ranks"] = gathered
        if dominant == "sx_sp_scan+gather_kernels":
            line["roofline"]["note"] = ("dominant kernels stage the findings for the host: the step is bound by the PCIe download of "
                                        f"{d2h_step} bytes per step, the HBM fraction is not its ceiling")
        if not args.no_cpu:
            sample_bytes = min(hi - lo, args.cpu_sample_mib << 20) // SLICE * SLICE
            sample = host_range(cfg["seed"], pidx, lo, sample_bytes)
            mo = [dataclasses.replace(to_oracle(m), counter_offset=lo) for m in missions]
            if cpu_first and cpu_first[1] >= sample_bytes:
                v, dtc = cpu_first[1] / GIB / cpu_first[0], cpu_first[0]
            else:
                v, dtc = cpu_ref_shape(mo, sample)
            line["cpu_baseline"] = {"value": v, "unit": "GiB/s", "cores": len(mo), "kind": "port",
                                    "sample": f"first {sample_bytes >> 20} MiB of rank 0's range, {dtc:.1f} s, one scanning thread per "
                                              "mission (the reference runs one thread per encoding, main.rs:151-167)",
                                    "host_cores_available": os.cpu_count()}
            cores = os.cpu_count() or 1
            v2, dt2, njobs = cpu_all_cores(mo, sample, cores)
            line["cpu_baseline_all_cores"] = {"value": v2, "unit": "GiB/s", "cores": min(cores, njobs), "kind": "port",
                                              "sample": f"the same {sample_bytes >> 20} MiB, range-sharded into {njobs} independent scans "
                                                        f"(64 KiB pre-roll each), {dt2:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity and parity["mismatches"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
