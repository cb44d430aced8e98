"""N>1 host logic on CPU: world_size-2 gloo processes exercise the table/sample
bucketing and the blob rendezvous used by the table-sharded mode (the CUDA side
is covered by tests/test_multigpu_gpu.py on a 2-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from deeprecsys_b200 import shard


def test_bucketing_is_a_partition():
    for world in (1, 2, 3, 4, 8):
        for T in (1, 4, 8, 10, 27):
            owned = [shard.owned_tables(T, r, world) for r in range(world)]
            flat = sorted(t for o in owned for t in o)
            assert flat == list(range(T))
            for r, o in enumerate(owned):
                assert all(shard.table_owner(t, world) == r for t in o)
        for B in (1, 2, 7, 1024, 1025, 16384):
            covered = []
            for r in range(world):
                row0, rows = shard.sample_slice(B, r, world)
                covered += list(range(row0, row0 + rows))
                for b in range(row0, row0 + rows):
                    assert shard.sample_owner(b, B, world) == r
            assert covered == list(range(B))
    # RMC1 at 8 GPUs: one table per rank; RMC3 (10 tables): 2,2,1,1,1,1,1,1
    assert [len(shard.owned_tables(8, r, 8)) for r in range(8)] == [1] * 8
    assert [len(shard.owned_tables(10, r, 8)) for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]


class _FakeEngine(object):
    def __init__(self, rank):
        self.rank = rank
        self.connected = None

    def shard_export(self):
        return bytes([self.rank]) * 192

    def shard_connect(self, rank, world, blobs):
        self.connected = (rank, world, blobs)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _FakeEngine(rank)
        r, w = shard.connect(eng, dist)
        assert (r, w) == (rank, world)
        _, _, blobs = eng.connected
        assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 192 for b in blobs)
        B, out_dim = 11, 2
        row0, rows = shard.sample_slice(B, rank, world)
        mine = np.arange(B * out_dim, dtype=np.float32).reshape(B, out_dim)[row0:row0 + rows]
        full = shard.assemble_outputs(dist, mine, B, out_dim)
        assert np.array_equal(full, np.arange(B * out_dim, dtype=np.float32).reshape(B, out_dim))
        q.put((rank, "ok"))
    except Exception as exc:   # pragma: no cover
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_rendezvous_and_assembly():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == [(0, "ok"), (1, "ok")], results
