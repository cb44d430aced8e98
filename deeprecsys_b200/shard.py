"""Host-side logic of the multi-GPU modes (one process per GPU).

replica  every rank holds all tables and serves its own queries: the path
         partitions by query, no data-path collective (weak scaling).
table    table-wise sharding: table t lives on rank t % world; for a
         query of B samples rank r runs the dense part for samples
         [r*S, min(B, (r+1)*S)), S = ceil(B / world), and gathers its tables for
         all B samples, storing pooled vectors straight into the owner rank's
         buffer through peer-mapped memory (the C library does that;
         torch.distributed is only the rendezvous that carries the IPC handles).

row      row-wise sharding: rank r holds rows [r*chunk, min(N, (r+1)*chunk)),
         chunk = ceil(N / world), of EVERY table.  Each rank buckets the query's ids
         by its row range (order kept, rebased to the slice), gathers partial pooled
         vectors for all B samples into the owner rank's partial slab, and the owner
         adds the world partials in rank order.  Useful when T < world (DIN: 4 tables)
         or one table outgrows a GPU; reassociates the fp32 sum (not bit-exact).

The integer bucketing here mirrors csrc/engine.cu / csrc/rowshard.cuh exactly (owns(),
enqueue_graph(), k_row_bucket) and is what the gloo tests check.
"""
from __future__ import absolute_import, division, print_function


def owned_tables(n_tables, rank, world):
    return [t for t in range(n_tables) if world == 1 or t % world == rank]


def table_owner(t, world):
    return 0 if world == 1 else t % world


def row_range(n_rows, rank, world):
    """[row0, row1) of a table's rows held by ``rank`` under row-wise sharding."""
    n_rows, world = int(n_rows), int(world)
    chunk = -(-n_rows // world)
    row0 = min(n_rows, rank * chunk)
    return row0, min(n_rows, row0 + chunk)


def row_owner(row, n_rows, world):
    return int(row) // (-(-int(n_rows) // int(world)))


def bucket_rows(ids, lengths, row0, row1):
    """Integer reference of k_row_bucket: keep, in order, the ids of every segment that fall
    into [row0, row1), rebased to the slice.  -> (local_ids int32, local_lengths int32)."""
    import numpy as np
    ids = np.asarray(ids).astype(np.int64)
    lengths = np.asarray(lengths).astype(np.int64)
    mine = (ids >= row0) & (ids < row1)
    seg = np.repeat(np.arange(lengths.shape[0]), lengths)
    local_len = np.bincount(seg[mine], minlength=lengths.shape[0]).astype(np.int32)
    return (ids[mine] - row0).astype(np.int32), local_len


def slice_size(batch, world):
    return -(-int(batch) // int(world))


def sample_slice(batch, rank, world):
    """[row0, row0 + rows) of the query's samples whose dense part / MLPs run on ``rank``."""
    S = slice_size(batch, world)
    row0 = rank * S
    rows = max(0, min(S, batch - row0))
    return row0, rows


def sample_owner(b, batch, world):
    return int(b) // slice_size(batch, world)


def exchange_blobs(dist, blob):
    """all-gather the per-rank IPC blobs (any torch.distributed backend)."""
    world = dist.get_world_size()
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    return blobs


def connect(engine, dist):
    """Export this rank's exchange slab, swap handles, map the peers."""
    rank, world = dist.get_rank(), dist.get_world_size()
    blobs = exchange_blobs(dist, engine.shard_export())
    engine.shard_connect(rank, world, blobs)
    dist.barrier()
    return rank, world


def assemble_outputs(dist, out_slice, batch, out_dim):
    """Gather every rank's output slice to all ranks -> [batch, out_dim] (host side, for tests)."""
    import numpy as np
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, np.asarray(out_slice))
    full = np.concatenate([p.reshape(-1, out_dim) for p in parts], axis=0)
    assert full.shape[0] == batch, (full.shape, batch)
    return full
