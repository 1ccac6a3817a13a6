#!/usr/bin/env python3
"""bench.py -- scanned GiB/s of the stringsext scanner hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (sx_scan_stream: scan kernel + text materialisation + result
download) over one synthetic buffer.  N=1 runs BASELINE.json configs[1] (`-e utf-8 -n 10` over a
4 GiB random buffer); N>1 runs one encoding per GPU (configs[2..4]), one process per GPU, no
collective on the data path (weak scaling: every GPU scans the whole buffer for its encoding).

  value     whole-job GiB/s with the input resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the C ABI with a pinned HOST buffer (H2D inside the timed region)
  roofline  the dominant kernel (sx_prefilter_kernel, the one pass over every input byte): input bytes / its
            CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs; `pipeline_*` covers all kernels of a step
  cpu_baseline  the CPU oracle (a port of the reference algorithm; the Rust reference cannot be
            built here) on a bounded sample of the same buffer, one scanning thread per mission
            exactly like the reference (main.rs:151-167)

`--impl reference` times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
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

# BASELINE.json configs -> (label, n, ubf name) per rank, buffer size.  Encodings the kernels do not
# implement yet (big5, euc-jp) are replaced and the replacement is named in config.substituted.
CONFIGS = {
    1: dict(name="-e utf-8 -n 10 over 4 GiB random buffer, 1xB200", size=4 << 30, n=10, seed=2,
            missions=[("utf-8", None)]),
    2: dict(name="-e utf-16le -e utf-16be -n 10 -u African over 4 GiB, 2xB200 (one encoding per GPU)", size=4 << 30,
            n=10, seed=3, missions=[("utf-16le", "African"), ("utf-16be", "African")]),
    4: dict(name="-e utf-8 -e utf-16le -e utf-16be -e big5 -n 8 over 16 GiB, 4xB200", size=16 << 30, n=8, seed=4,
            missions=[("utf-8", None), ("utf-16le", None), ("utf-16be", None), ("ascii", None)],
            substituted={"big5": "ascii: big5 is not implemented (needs the WHATWG index); on random bytes with the default "
                                 "filters big5 -n 8 prints ASCII runs almost only (SURVEY.md 8a: ~181 findings/MB), the "
                                 "x-user-defined `ascii` mission has the closest finding density (~240/MB)"}),
    8: dict(name="8 encodings (ascii, utf-8, utf-16le, utf-16be, utf-32le, utf-32be, euc-jp, koi8-r) -n 6 over 32 GiB, 8xB200",
            size=32 << 30, n=6, seed=5,
            missions=[("ascii", None), ("utf-8", None), ("utf-16le", None), ("utf-16be", None), ("utf-32le", None),
                      ("utf-32be", None), ("ascii", None), ("koi8-r", None)],
            substituted={"euc-jp": "ascii: euc-jp is not implemented (needs the WHATWG jis0208/jis0212 indexes); with the "
                                   "default filters its kanji (3-byte UTF-8 leads) do not pass, so on random bytes it prints "
                                   "ASCII runs: 1717 findings/MB vs 1670/MB for `ascii -n 6` (SURVEY.md 8a)"}),
}


def config_for(n_gpus: int):
    if n_gpus in CONFIGS:
        return CONFIGS[n_gpus]
    c = dict(CONFIGS[8])
    c["missions"] = c["missions"][:n_gpus]
    return c


def make_mission(sx, label, ubf_name, n, mission_id=0):
    ubf = sx.UBF_AFRICAN if ubf_name == "African" else None
    return sx.Mission.for_label(label, n, ubf=ubf, mission_id=mission_id)


def plant_patches(seed, enc_id, n, q, length, per_mib=1):
    """(offset, bytes) patches = the planted corpus (tests/corpus.py), applied in order."""
    import random

    import corpus

    rng = random.Random(seed * 7919 + enc_id)
    strings = corpus.planted_strings(rng, enc_id, n, q)
    count = max(12, int(length // (1 << 20)) * per_mib)
    out = []
    W, slice_len = 2 * q, 4096
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


def host_range(seed, patches, start, length):
    """Regenerate bytes [start, start+length) of the benchmark buffer on the host."""
    import corpus
    import numpy as np

    buf = corpus.sx_mix_bytes(seed, start, length)
    for off, s in patches:
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


def cpu_port_throughput(missions_o, sample, threads):
    """Reference-shaped CPU run of the oracle: one scanning thread per mission over the same sample."""
    from oracle import oracle as O

    states = [O.OState(m) for m in missions_o]
    L = O.lib()
    import ctypes as C

    def work(ss):
        fc = L.sxo_scan_stream(ss._h, -1, C.c_void_p(sample.ctypes.data), sample.size, 4096, 0)
        n = L.sxo_fc_len(fc)
        L.sxo_fc_free(fc)
        return n

    t0 = time.perf_counter()
    if len(states) == 1:
        work(states[0])
    else:
        ths = [threading.Thread(target=work, args=(s,)) for s in states]
        [t.start() for t in ths]
        [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return len(states) * sample.size / GIB / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import stringsext_b200.mission as M
    from helpers import to_oracle

    cfg = config_for(args.gpus)
    missions = [make_mission(M, lbl, u, cfg["n"], i) for i, (lbl, u) in enumerate(cfg["missions"])][: args.gpus]
    sample_bytes = args.cpu_sample_mib << 20
    patches = plant_patches(cfg["seed"], missions[0].encoding_id, cfg["n"], 64, cfg["size"])
    sample = host_range(cfg["seed"], patches, 0, sample_bytes)
    mo = [to_oracle(m) for m in missions]
    times = []
    # bounded: the CPU port runs at tens of MiB/s per thread, so cap the number of timed passes
    n_warm, n_steps = min(args.warmup, 1), min(args.steps, 3)
    for i in range(n_warm + n_steps):
        v, dt = cpu_port_throughput(mo, sample, len(mo))
        if i >= n_warm:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = len(mo) * sample.size / GIB / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "scanned GiB/s (whole job, all encodings)", "value": val, "unit": "GiB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample_bytes_per_step": sample.size, "timed_passes": n_steps,
                   "note": "CPU port of the reference algorithm (oracle/); the Rust reference cannot be built in this image"},
        "cpu_baseline": {"value": val, "unit": "GiB/s", "cores": len(mo), "kind": "port",
                         "sample": f"first {args.cpu_sample_mib} MiB of the workload buffer per step, one thread per mission"},
        "e2e": {"value": val, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size-mib", type=int, default=0, help="override the buffer size (debug only; invalidates the number)")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--e2e-max-mib", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--as-rank", type=int, default=-1, help="debug: single process scanning rank R's encoding of the --gpus config")
    ap.add_argument("--sparse-mode", type=int, default=1, help="debug: 0 block kernel only, 1 default, 2 sparse pipeline whenever possible")
    ap.add_argument("--corpus", default="random", help="debug: 'text' = UTF-8 text (repeated 4 MiB block) instead of random bytes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import stringsext_b200 as sx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = config_for(args.gpus)
    size = cfg["size"] if not args.size_mib else args.size_mib << 20
    label, ubf_name = cfg["missions"][(args.as_rank if args.as_rank >= 0 else rank) % len(cfg["missions"])]
    mission = make_mission(sx, label, ubf_name, cfg["n"], rank)
    L = sx.load_library()

    # ---- synthetic input, resident in HBM --------------------------------------------------------
    dbuf = torch.empty(size, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    assert L.sx_fill_random(dbuf.data_ptr(), size, cfg["seed"], 0, local, None) == 0
    patches = plant_patches(cfg["seed"], mission.encoding_id, cfg["n"], 64, size)
    for off, s in patches:
        dbuf[off : off + len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()

    if args.corpus == "text":
        import random as _r

        import corpus as _c

        blk = _c.gen(_r.Random(7), "text", 4 << 20, mission.encoding_id)
        tblk = torch.frombuffer(bytearray(blk), dtype=torch.uint8).cuda()
        for off in range(0, size, len(blk)):
            n = min(len(blk), size - off)
            dbuf[off : off + n] = tblk[:n]
        torch.cuda.synchronize()
    ss_dev = sx.ScannerState(mission, local)
    ss_dev.set_sparse(args.sparse_mode)

    def step_device():
        ss_dev.reset()  # ScannerState::new semantics per pass (same work every step), device buffers are reused
        fc = ss_dev.scan_stream(None, False, 4096, device_ptr=dbuf.data_ptr(), length=size, cuda_stream=stream.cuda_stream, raw=True)
        return len(fc), ss_dev.last_stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, mat_ms, pre_ms, lst_ms, ex_ms, launches, nfind, d2h = [], [], [], [], [], 0, 0, 0
    sp_ms, sparse_used = [], 0
    host_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        nfind, st = step_device()
        scan_ms.append(st.scan_kernel_ms)
        mat_ms.append(st.materialize_kernel_ms)
        pre_ms.append(st.prefilter_kernel_ms)
        lst_ms.append(st.list_kernels_ms)
        ex_ms.append(st.exact_kernel_ms)
        sp_ms.append([float(x) for x in st.sparse_stage_ms])
        sparse_used = int(st.sparse_used)
        win_total, win_listed = st.windows_total, st.windows_listed
        host_ms.append((st.host_total_ms, st.host_post_ms, *[float(x) for x in st.host_phase_ms]))
        launches += st.kernel_launches
        d2h = st.d2h_bytes
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * size / GIB / (ms_per_step / 1e3)

    # ---- e2e: pinned host buffer through the C ABI -----------------------------------------------
    e2e_size = min(size, args.e2e_max_mib << 20)
    hbuf = torch.empty(e2e_size, dtype=torch.uint8, pin_memory=True)
    hbuf.copy_(dbuf[:e2e_size])
    torch.cuda.synchronize()
    harr = hbuf.numpy()
    # the device-timed state and its buffer go first: at the 32 GiB configs they hold most of the HBM
    ss_dev.close()
    del dbuf
    torch.cuda.empty_cache()

    ss_host = sx.ScannerState(mission, local)

    def step_host():
        ss_host.reset()
        fc = ss_host.scan_stream(harr, False, 4096, cuda_stream=stream.cuda_stream, raw=True)
        return len(fc), ss_host.last_stats

    step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        n_e2e, st_e = step_host()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_size / GIB / float(t.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        avg = lambda v: sum(v) / len(v)
        kernels = {"sx_prefilter_kernel": avg(pre_ms), "sx_list_offsets_kernel": avg(lst_ms), "sx_materialize_kernel": avg(mat_ms)}
        if sparse_used:
            # exact stage = the sparse-list pipeline (sx_sparse_utf8.cuh): one entry per kernel (group), CUDA events between them
            names = ("sx_list_compact_kernel", "sx_sp_heads_kernel", "sx_sp_members_kernel", "sx_sp_fix_kernel", "sx_sp_ext_kernel",
                     "sx_sp_scan+gather")
            for i, nm in enumerate(names):
                kernels[nm] = avg([v[i] for v in sp_ms])
            kernels["exact_stage_total(incl. list-length round trip)"] = avg(ex_ms)
            candidates = ("sx_prefilter_kernel",) + names
        else:
            kernels["sx_exact_kernel"] = avg(ex_ms)
            candidates = ("sx_prefilter_kernel", "sx_exact_kernel")
        dominant = max(candidates, key=lambda k: kernels[k])
        k_ms = kernels[dominant]
        # algorithmic bytes per launch: the prefilter reads every input byte once; the exact
        # kernel reads the listed windows (128 B each at the default geometry)
        alg_bytes = size if dominant == "sx_prefilter_kernel" else int(win_listed) * 2 * 64
        achieved = alg_bytes / 1e9 / (k_ms / 1e3)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tr.get("workload") == cfg["name"] and not args.size_mib and dominant in tr:
                traffic = tr[dominant]["dram_bytes_read"] + tr[dominant]["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": "scanned GiB/s (whole job, all encodings)", "value": value, "unit": "GiB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["name"], "bytes_per_gpu": size, "encoding_rank0": label,
                       "chars_min_nb": cfg["n"], "slice_len": 4096, "output_line_char_nb_max": 64,
                       "l2": "input (>= 4 GiB) is larger than L2; no flush needed",
                       "input": "splitmix64 counter-based random bytes + 1 planted string per MiB",
                       "findings_rank0": nfind, "substituted": cfg.get("substituted", {}),
                       "size_overridden": bool(args.size_mib)},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "GiB/s", "h2d_bytes_per_step": int(st_e.h2d_bytes),
                    "d2h_bytes_per_step": int(st_e.d2h_bytes), "bytes_scanned_per_gpu": e2e_size, "findings_rank0": n_e2e,
                    "ms_per_step_rank0": dt * 1e3, "host_phase_ms_rank0": [float(x) for x in st_e.host_phase_ms],
                    "exact_stage_ms_rank0": float(st_e.exact_kernel_ms), "sparse_used": int(st_e.sparse_used)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "peak_note": "the measured peak is a device copy (read + write); a read-only pass can slightly exceed it",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "kernels_ms": kernels,
                         "pipeline_ms": avg(scan_ms), "pipeline_gbs": size / 1e9 / (avg(scan_ms) / 1e3),
                         "windows_total": int(win_total), "windows_listed": int(win_listed),
                         "host_call_ms": avg([h[0] for h in host_ms]), "host_post_ms": avg([h[1] for h in host_ms]),
                         "host_phase_ms": [avg([h[2 + k] for h in host_ms]) for k in range(4)]},
        }
        if dominant == "sx_sp_scan+gather":
            # output-heavy workloads: the dominant kernel stores the findings (48 B each) and their text into pinned host
            # memory, so its ceiling is the PCIe link (measured D2H copy on this pool: ~55 GB/s, tools/ubench/d2h.py)
            line["roofline"]["note"] = "dominant kernel writes the findings to host memory: PCIe-bound, the HBM fraction is not its ceiling"
            line["roofline"]["pcie"] = {"bytes_per_launch": int(d2h), "achieved_gbs": d2h / 1e9 / (k_ms / 1e3), "peak_gbs": 55.0,
                                        "frac": d2h / 1e9 / (k_ms / 1e3) / 55.0}
        if not args.no_cpu:
            from helpers import to_oracle

            sample_bytes = min(size, args.cpu_sample_mib << 20)
            sample = host_range(cfg["seed"], patches, 0, sample_bytes)
            v, dtc = cpu_port_throughput([to_oracle(mission)], sample, 1)
            line["cpu_baseline"] = {"value": v, "unit": "GiB/s", "cores": 1, "kind": "port",
                                    "sample": f"first {sample_bytes >> 20} MiB of the same buffer, {dtc:.1f} s, one scanning thread "
                                              "(the reference runs one thread per encoding, main.rs:151-167)",
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
