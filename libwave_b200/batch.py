"""Batch-of-scans sharding (BASELINE.json config 5): independent scan pairs are dealt
round-robin to the ranks (one process per GPU), matched with no data-path communication, and the
per-scan result records are gathered once at the end.  The reference's only parallelism is the same
thing with host threads (MultiMatcher, impl/multi_matcher_impl.hpp:45-48).

Works over any torch.distributed backend: NCCL on the GPU box (bench.py), gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

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


class ScanBatch:
    """Python mirror of the C batch API (include/wavecu.h, wavecu_batch_*): scans spread over the GPUs of
    this process with several concurrent matches per GPU, an optional shared map indexed once per GPU, and
    - for multi-process jobs - NCCL inside the library: map broadcast and ONE all-gather of the records."""

    def __init__(self, params=None, devices=None, workers_per_device: int = 8):
        self._L = capi.lib()
        self._h = C.c_void_p()
        prm = params.to_c() if params is not None else None
        dev = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
        capi.check(self._L.wavecu_batch_create(C.byref(prm) if prm is not None else None,
                                               dev.ctypes.data_as(C.POINTER(C.c_int)), len(dev), workers_per_device,
                                               C.byref(self._h)))
        self._keep = []
        self.rank, self.world = 0, 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.wavecu_batch_destroy(h)
            self._h = None

    @staticmethod
    def _xyzw(a):
        a = np.asarray(a, dtype=np.float32)
        if a.ndim == 2 and a.shape[1] == 4 and a.flags.c_contiguous:
            return a
        out = np.ones((a.shape[0], 4), dtype=np.float32)
        out[:, :3] = a[:, :3]
        return out

    def set_map(self, cloud):
        a = self._xyzw(cloud)
        self._keep = [a]
        capi.check(self._L.wavecu_batch_set_map(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0]))

    def init_comm(self, rank: int, world: int, exchange):
        """exchange(bytes_or_None) -> bytes: hands rank 0's 128-byte id to every rank (the caller's transport:
        torch.distributed.broadcast_object_list, a file, MPI ...)."""
        buf = (C.c_char * 128)()
        if rank == 0:
            capi.check(self._L.wavecu_batch_unique_id(buf))
        raw = exchange(bytes(buf) if rank == 0 else None)
        idbuf = (C.c_char * 128).from_buffer_copy(raw)
        capi.check(self._L.wavecu_batch_init_comm(self._h, idbuf, rank, world))
        self.rank, self.world = rank, world

    def broadcast_map(self, cloud_on_root, n: int, root: int = 0):
        ptr = None
        if self.rank == root:
            a = self._xyzw(cloud_on_root)
            self._keep = [a]
            ptr = a.ctypes.data_as(C.POINTER(C.c_float))
        capi.check(self._L.wavecu_batch_broadcast_map(self._h, ptr, n, root))

    def match(self, scans, scan_ids=None, targets=None, with_info: bool = False):
        """scans: list of (n,4) fp32 C-contiguous arrays (page-locked or not).  Returns the record array."""
        n = len(scans)
        arrs = [self._xyzw(s) for s in scans]
        ptrs = (C.POINTER(C.c_float) * max(n, 1))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        sizes = (C.c_size_t * max(n, 1))(*[a.shape[0] for a in arrs])
        ids = None
        if scan_ids is not None:
            ids = (C.c_int * max(n, 1))(*[int(i) for i in scan_ids])
        tptrs, tsizes, tarrs = None, None, []
        if targets is not None:
            tarrs = [self._xyzw(t) for t in targets]
            tptrs = (C.POINTER(C.c_float) * max(n, 1))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in tarrs])
            tsizes = (C.c_size_t * max(n, 1))(*[a.shape[0] for a in tarrs])
        out = (capi.BatchRecordC * max(n, 1))()
        capi.check(self._L.wavecu_batch_match(self._h, ptrs, sizes, ids, n, tptrs, tsizes, int(with_info), out))
        return out

    def allgather(self, local, n_local: int):
        """local: BatchRecordC array with n_local used slots (same n_local on every rank)."""
        allr = (capi.BatchRecordC * max(1, n_local * self.world))()
        capi.check(self._L.wavecu_batch_allgather(self._h, local, n_local, allr))
        return allr


def records_to_table(records, n_scans: int) -> np.ndarray:
    """(n_scans, RECORD) float64 table ordered by scan id from an array of BatchRecordC."""
    out = np.full((n_scans, RECORD), np.nan)
    for r in records:
        if 0 <= r.scan_id < n_scans:
            out[r.scan_id, :16] = np.frombuffer(r.T, dtype=np.float64, count=16)
            out[r.scan_id, 16], out[r.scan_id, 17] = float(r.converged), float(r.iterations)
    return out
