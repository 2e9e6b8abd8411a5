"""CPU test (-m "not gpu") of the N > 1 host path: two gloo ranks shard a batch of scans, gather the
result records and aggregate the timing exactly as bench.py does over NCCL."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_scans, out):
    import torch.distributed as dist

    from libwave_b200 import batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = batch.shard_scan_ids(n_scans, rank, world)
    local = {}
    for k in mine:  # a fake "match": the record encodes the scan id so the gather can be checked
        T = np.eye(4)
        T[0, 3] = 0.1 * k
        local[k] = batch.pack_record(T, k % 3 != 0, 10 + k)
    table = batch.gather_records(local, n_scans)
    t, u = batch.reduce_timing(100.0 + 50.0 * rank, float(len(mine)))
    if rank == 0:
        out.put((mine, table, t, u))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gather_and_timing():
    world, n_scans = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scans, q)) for r in range(world)]
    for p in procs:
        p.start()
    mine, table, t, u = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert mine == [0, 2, 4, 6]
    assert table.shape == (n_scans, 18) and not np.isnan(table).any()
    for k in range(n_scans):
        assert np.isclose(table[k, 3], 0.1 * k) and table[k, 16] == float(k % 3 != 0) and table[k, 17] == 10 + k
    assert t == 150.0 and u == float(n_scans)   # max over ranks of the time, sum of the units


def test_shards_are_a_partition():
    from libwave_b200 import batch
    for world in (1, 2, 4, 8):
        ids = sorted(k for r in range(world) for k in batch.shard_scan_ids(256, r, world))
        assert ids == list(range(256))
        assert max(len(batch.shard_scan_ids(256, r, world)) for r in range(world)) == 256 // world
