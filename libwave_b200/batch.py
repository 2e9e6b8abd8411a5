"""Batch-of-scans sharding (BASELINE.json config 5): independent scan pairs are dealt
round-robin to the ranks (one process per GPU), matched with no data-path communication, and the
per-scan result records are gathered once at the end.  The reference's only parallelism is the same
thing with host threads (MultiMatcher, impl/multi_matcher_impl.hpp:45-48).

Works over any torch.distributed backend: NCCL on the GPU box (bench.py), gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

RECORD = 16 + 2  # 4x4 transform (row major) + converged flag + iteration count


def shard_scan_ids(n_scans: int, rank: int, world: int) -> list[int]:
    """Scan k goes to rank k mod world."""
    return list(range(rank, n_scans, world))


def pack_record(T, converged: bool, iterations: int) -> np.ndarray:
    r = np.empty(RECORD, dtype=np.float64)
    r[:16] = np.asarray(T, dtype=np.float64).reshape(16)
    r[16], r[17] = float(converged), float(iterations)
    return r


def gather_records(local: dict[int, np.ndarray], n_scans: int, device=None):
    """All ranks receive every scan's record, ordered by scan id: returns (n_scans, RECORD) float64.
    One all_gather of world * ceil(n_scans / world) records (256 scans: ~37 KB per rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    per = (n_scans + world - 1) // world
    host = np.full((per, RECORD + 1), -1.0, dtype=np.float64)
    for slot, k in enumerate(shard_scan_ids(n_scans, rank, world)):
        host[slot, 0] = float(k)
        host[slot, 1:] = local[k]
    buf = torch.from_numpy(host).to(device) if device is not None else torch.from_numpy(host)  # one copy
    if world > 1:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        allb = torch.cat(parts, dim=0)
    else:
        allb = buf
    allb = allb.cpu().numpy()
    out = np.full((n_scans, RECORD), np.nan)
    for row in allb:
        if row[0] >= 0:
            out[int(row[0])] = row[1:]
    return out


def reduce_timing(total_ms: float, units: float, device=None):
    """(max over ranks of the time, sum over ranks of the units) - the contract's aggregation."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(total_ms), float(units)
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
