"""Host-side logic of the multi-GPU modes (one process per GPU).

replica  every rank holds all tables and serves its own queries: the path
         partitions by query, no data-path collective (weak scaling).
table    table-wise sharding: table t lives on rank t % world; for a
         query of B samples rank r runs the dense part for samples
         [r*S, min(B, (r+1)*S)), S = ceil(B / world), and gathers its tables for
         all B samples, storing pooled vectors straight into the owner rank's
         buffer through peer-mapped memory (the C library does that;
         torch.distributed is only the rendezvous that carries the IPC handles).

The integer bucketing here mirrors csrc/engine.cu exactly (owns(),
enqueue_graph()) and is what the gloo tests check.
"""
from __future__ import absolute_import, division, print_function


def owned_tables(n_tables, rank, world):
    return [t for t in range(n_tables) if world == 1 or t % world == rank]


def table_owner(t, world):
    return 0 if world == 1 else t % world


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
