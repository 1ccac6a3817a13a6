"""world_size-2 gloo test of the N>1 host logic (mission sharding, gather, merge order) on the CPU.
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
