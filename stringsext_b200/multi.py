"""Sharding a scan job over several GPUs, one process per GPU: by mission or by stream range.

The reference runs one scanner thread per `--encoding` mission and merges the per-slice results in a
separate thread (/root/reference/src/main.rs:97-167).  Missions are independent, so here rank r of a
`torch.distributed` job scans the whole stream for the missions assigned to it on its own GPU; there is no
collective on the data path.  Only the (sparse) findings travel: they are gathered on rank 0 and merged in
the order of `impl PartialOrd for Finding` (finding.rs:92-109).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

from .mission import Mission


def missions_for_rank(missions: Sequence[Mission], rank: int, world_size: int) -> List[Mission]:
    """Round-robin assignment: mission i runs on rank i % world_size."""
    return [m for i, m in enumerate(missions) if i % world_size == rank]


def merge_findings(per_mission: Sequence[Sequence[tuple]]) -> List[tuple]:
    """k-way merge of per-mission finding lists (each position-monotone) by (position, mission_id).
    Items are tuples whose first two fields are (position, mission_id)."""
    allf = [f for lst in per_mission for f in lst]
    allf.sort(key=lambda f: (f[0], f[1]))  # stable: keeps each mission's emission order
    return allf


def scan_sharded(missions: Sequence[Mission], data, scan_fn: Callable[[Mission, object], List[tuple]], dist=None):
    """Run `scan_fn(mission, data)` for this rank's missions, gather everything on rank 0 and merge.
    `dist`: the torch.distributed module (initialised) or None for a single process.
    Returns the merged list on rank 0, None elsewhere."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    mine = [scan_fn(m, data) for m in missions_for_rank(missions, rank, world)]
    if dist is None or world == 1:
        return merge_findings(mine)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    return merge_findings([lst for per_rank in gathered for lst in per_rank])


# ---- range sharding ----------------------------------------------------------------------------
# A mission's stream is cut at slice boundaries; rank r holds bytes [base_r, hi_r) = its range plus a halo in front and
# scans [lo_r, hi_r) with `sx_scan_range(..., SX_RANGE_PREFIX_UNKNOWN)`: the carry into the range is derived on the
# device inside the halo.  No data flows between ranks; the per-range collections concatenate in rank order.
def range_plan(size: int, world_size: int, slice_len: int = 4096, halo: int = 1 << 20) -> List[Tuple[int, int, int]]:
    """[(base, lo, hi)] per rank: slice-aligned ranges [lo, hi) covering [0, size), base = start of the halo in front."""
    per = -(-size // world_size)
    per = -(-per // slice_len) * slice_len
    out = []
    for r in range(world_size):
        lo, hi = min(size, r * per), min(size, (r + 1) * per)
        out.append((max(0, lo - (halo // slice_len) * slice_len), lo, hi))
    return out


def scan_range_sharded(missions: Sequence[Mission], size: int, scan_fn: Callable[[Mission, int, int, int], List[tuple]], dist=None,
                       slice_len: int = 4096, halo: int = 1 << 20):
    """Every rank runs `scan_fn(mission, base, lo, hi)` -- the findings emitted for stream bytes [lo, hi), scanned from a
    buffer that starts at `base` -- for EVERY mission over its own range; rank 0 concatenates each mission's lists in
    rank order (they are position-monotone and disjoint) and merges the missions.  Returns the merged list on rank 0."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    base, lo, hi = range_plan(size, world, slice_len, halo)[rank]
    mine = [scan_fn(m, base, lo, hi) if hi > lo else [] for m in missions]
    if dist is None or world == 1:
        return merge_findings(mine)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    per_mission = [[f for per_rank in gathered for f in per_rank[i]] for i in range(len(missions))]
    return merge_findings(per_mission)
