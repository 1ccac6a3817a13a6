"""world_size-2 gloo tests of the N>1 host logic (mission sharding and range sharding, gather, merge order) on the CPU.
The per-mission scan is injected: on the GPU box it is the CUDA scanner (test_gpu_parity.py), here it is
the oracle, because only the plumbing is under test."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import corpus
    from helpers import M, oracle_state
    from stringsext_b200 import multi

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    missions = [M.Mission.for_label(lbl, 6, mission_id=i) for i, lbl in enumerate(["ascii", "utf-8", "utf-16le"])]
    buf = corpus.sx_mix_bytes(11, 0, 1 << 18)
    corpus.plant(buf, 11, 2, 6, 64, density=1 << 12)

    def scan_fn(m, data):
        return [(f.position, m.mission_id, f.precision, f.s, f.completes) for f in oracle_state(m).scan_stream(data).v]

    merged = multi.scan_sharded(missions, buf, scan_fn, dist)
    dist.barrier()
    if rank == 0:
        single = multi.scan_sharded(missions, buf, scan_fn, None)
        ok = merged == single and len(merged) > 10
        # merged order: non-decreasing position, ties by mission id
        keys = [(f[0], f[1]) for f in merged]
        ok = ok and keys == sorted(keys)
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_ranks_shard_missions_and_merge(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_missions_for_rank():
    sys.path.insert(0, ROOT)
    from stringsext_b200 import multi

    ms = list(range(5))
    assert multi.missions_for_rank(ms, 0, 2) == [0, 2, 4]
    assert multi.missions_for_rank(ms, 1, 2) == [1, 3]


def _range_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dataclasses

    import torch.distributed as dist

    import corpus
    from helpers import M, oracle_state
    from stringsext_b200 import multi

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    missions = [M.Mission.for_label(lbl, 6, mission_id=i) for i, lbl in enumerate(["koi8-r", "utf-8", "big5"])]
    size = (1 << 19) + 4096 * 3 + 100
    buf = corpus.sx_mix_bytes(13, 0, size)
    corpus.plant(buf, 13, 1, 6, 64, density=1 << 12)
    halo = 1 << 16

    def scan_fn(m, base, lo, hi):
        # what sx_scan_range gives on the GPU box, emulated with the oracle: scan [base, hi) with the state's byte counter
        # at `base`, keep the findings of [lo, hi) (the carry at `lo` is exact once the halo has been scanned)
        mo = dataclasses.replace(m, counter_offset=base)
        end = min(size, hi + 4096)
        fs = oracle_state(mo).scan_stream(buf[base:end]).v
        return [(f.position, m.mission_id, f.precision, f.s, f.completes) for f in fs if lo <= f.position < hi]

    merged = multi.scan_range_sharded(missions, size, scan_fn, dist, halo=halo)
    dist.barrier()
    if rank == 0:
        whole = multi.merge_findings([[(f.position, m.mission_id, f.precision, f.s, f.completes)
                                       for f in oracle_state(m).scan_stream(buf).v] for m in missions])
        ok = merged == whole and len(merged) > 100
        open(out_path, "w").write("ok" if ok else f"mismatch {len(merged)} {len(whole)}")
    dist.destroy_process_group()


def test_two_ranks_shard_ranges_and_merge(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_range_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_range_plan():
    sys.path.insert(0, ROOT)
    from stringsext_b200 import multi

    plan = multi.range_plan((4 << 30) + 5, 8)
    assert plan[0][:2] == (0, 0) and plan[-1][2] == (4 << 30) + 5
    assert all(lo % 4096 == 0 and base % 4096 == 0 and base == max(0, lo - (1 << 20)) for base, lo, hi in plan)
    assert all(a[2] == b[1] for a, b in zip(plan, plan[1:]))
    assert multi.range_plan(100, 4)[1:] == [(0, 100, 100)] * 3  # more ranks than slices: empty ranges
