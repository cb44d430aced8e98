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


def test_row_ranges_partition_every_table():
    for world in (2, 3, 4, 8):
        for n in (1, 5, 8, 9, 1000, 4000000, 28000):
            ranges = [shard.row_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            for r, (r0, r1) in enumerate(ranges):
                for row in {r0, r1 - 1} if r1 > r0 else ():
                    assert shard.row_owner(row, n, world) == r


def test_row_bucketing_is_order_preserving_and_lossless():
    """Integer reference of csrc/rowshard.cuh's k_row_bucket (the GPU is compared against this
    bit-for-bit in tests/_shard_worker.py): the per-rank buckets partition every segment's ids,
    keep their order and rebase them to the rank's slice."""
    from oracle import numpy_oracle as O
    rng = np.random.RandomState(5)
    n_rows, world = 1003, 4
    lengths = rng.randint(0, 30, size=41).astype(np.int32)
    lengths[7] = 0
    ids = rng.randint(0, n_rows, size=int(lengths.sum()))
    seg = np.repeat(np.arange(lengths.shape[0]), lengths)
    total = np.zeros_like(lengths)
    for r in range(world):
        row0, row1 = shard.row_range(n_rows, r, world)
        lid, llen = shard.bucket_rows(ids, lengths, row0, row1)
        assert lid.dtype == np.int32 and llen.dtype == np.int32
        assert lid.size == llen.sum() and (lid >= 0).all() and (lid < row1 - row0).all()
        total += llen
        pos = 0
        for b, k in enumerate(llen):                  # original order, rebased to the slice
            want = [v - row0 for v in ids[seg == b] if row0 <= v < row1]
            assert lid[pos:pos + k].tolist() == want
            pos += k
    assert np.array_equal(total, lengths)             # every id lands in exactly one bucket
    # and the row-wise pooled sum equals the sequential one up to fp32 reassociation
    W = rng.uniform(-1, 1, (n_rows, 8)).astype(np.float32)
    a = O.sparse_lengths_sum(W, ids, lengths)
    b = O.sparse_lengths_sum_rowwise(W, ids, lengths, world)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
    assert np.array_equal(O.sparse_lengths_sum_rowwise(W, ids, lengths, 1), a)


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
