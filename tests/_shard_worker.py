"""torchrun worker for tests/test_multigpu_gpu.py: table-sharded serving on
WORLD_SIZE GPUs, each rank checks its slice against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402

from deeprecsys_b200 import _lib, shard         # noqa: E402
from deeprecsys_b200.engine import AccelEngine  # noqa: E402
from oracle import numpy_oracle as O            # noqa: E402
from tests.helpers import make_case, request    # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    mode = sys.argv[2] if len(sys.argv) > 2 else "table"
    if mode == "row":
        return main_row(rank, world, local, precision)
    if mode == "nccl":
        return main_nccl(rank, world, local, precision)
    for name in ("dlrm_cat", "wnd", "ncf", "din_many", "dlrm_rm3_like", "mtwnd", "dlrm_rm2_like"):
        mb = 24 if name == "dlrm_rm3_like" else 96
        c = make_case(name, num_batches=2, max_batch=mb)       # same seed on every rank
        S = shard.slice_size(mb, world)
        eng = AccelEngine(args=c.args, device=local, max_batch=S, precision=precision,
                          max_lookups_per_table=mb * max(1, c.spec.L))
        shard.connect(eng, dist)
        for t in shard.owned_tables(c.spec.T, rank, world):
            eng.load_table(t, c.tables[t])
        for (mlp_id, layer), (W, b) in c.fcs.items():
            eng.load_fc(mlp_id, layer, W, b)
        for j in range(2):
            dense, ids, lengths = request(c, j, mb)
            eng.stage_batch(j, dense, ids, lengths)
        dist.barrier()
        for rep in range(3):                                   # exercises lane reuse + consumed flags
            for batch_id, B in [(0, mb), (1, mb - 5), (0, 3)]:
                dense, ids, lengths = request(c, batch_id, B)
                st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
                row0, rows = shard.sample_slice(B, rank, world)
                out, _ = eng.infer_staged(batch_id, B)
                out = out[:rows]
                np.testing.assert_allclose(out, st["out"][row0:row0 + rows], rtol=1e-4, atol=1e-5)
                if rows > 0:
                    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
                    want = st["pooled"][row0:row0 + rows]
                    assert np.array_equal(pooled.view(np.uint32), want.view(np.uint32)), \
                        "%s: pooled slice not bit-exact on rank %d" % (name, rank)
                if rep == 0 and B <= S:                        # host path too (slot capacity = S)
                    out_h, _ = eng.infer(dense, ids, lengths, B)
                    np.testing.assert_allclose(out_h[:rows], st["out"][row0:row0 + rows],
                                               rtol=1e-4, atol=1e-5)
                full = shard.assemble_outputs(dist, out, B, c.spec.out_dim)
                np.testing.assert_allclose(full, st["out"], rtol=1e-4, atol=1e-5)
        dist.barrier()
        eng.close()
        if rank == 0:
            print("sharded %s ok on %d GPUs (%s)" % (name, world, precision))
    dist.destroy_process_group()
    print("RANK %d DONE" % rank)


def main_nccl(rank, world, local, precision):
    """The UNFUSED exchange baseline (drs_exchange_begin / NCCL all_to_all_single /
    drs_exchange_end) must give what the fused peer-store exchange gives: pooled slices
    bit-exact, outputs equal to the oracle."""
    for name in ("dlrm_cat", "dlrm_rm3_like", "dlrm_rm2_like"):
        mb = 24 if name == "dlrm_rm3_like" else 96
        c = make_case(name, num_batches=2, max_batch=mb)
        S = shard.slice_size(mb, world)
        eng = AccelEngine(args=c.args, device=local, max_batch=S, precision=precision,
                          max_lookups_per_table=mb * max(1, c.spec.L))
        shard.connect(eng, dist)
        eng.set_option("exchange", 1)
        for t in shard.owned_tables(c.spec.T, rank, world):
            eng.load_table(t, c.tables[t])
        for (mlp_id, layer), (W, b) in c.fcs.items():
            eng.load_fc(mlp_id, layer, W, b)
        for j in range(2):
            dense, ids, lengths = request(c, j, mb)
            eng.stage_batch(j, dense, ids, lengths)
        dist.barrier()
        views = {}
        for rep in range(3):
            for batch_id, B in [(0, mb), (1, mb - 5), (0, 3)]:
                dense, ids, lengths = request(c, batch_id, B)
                st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
                row0, rows = shard.sample_slice(B, rank, world)
                send, recv, n, stream = eng.exchange_begin(batch_id, B)

                def view(ptr):
                    if (ptr, n) not in views:
                        class _Dev(object):
                            __cuda_array_interface__ = {"shape": (world * n,), "typestr": "<f4",
                                                        "data": (ptr, False), "version": 2}
                        views[(ptr, n)] = torch.as_tensor(_Dev(), device="cuda")
                    return views[(ptr, n)]
                with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                    dist.all_to_all_single(view(recv), view(send))
                eng.exchange_end()
                eng.sync()
                out = eng.fetch_stage(_lib.STAGE_OUT) if rows > 0 else np.zeros((0, c.spec.out_dim))
                np.testing.assert_allclose(out[:rows], st["out"][row0:row0 + rows], rtol=1e-4,
                                           atol=1e-5)
                if rows > 0:
                    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
                    want = st["pooled"][row0:row0 + rows]
                    assert np.array_equal(pooled.view(np.uint32), want.view(np.uint32)), \
                        "%s: pooled slice not bit-exact on rank %d (nccl exchange)" % (name, rank)
        dist.barrier()
        eng.close()
        if rank == 0:
            print("nccl-exchange %s ok on %d GPUs (%s)" % (name, world, precision))
    dist.destroy_process_group()
    print("RANK %d DONE" % rank)


def main_row(rank, world, local, precision):
    """Row-wise sharding: every rank holds a row slice of every table.  Checks, per rank:
    bucketed ids / local offsets bit-exact vs the integer reference, pooled slice bit-exact vs
    the oracle that emulates the rank-order reduction (and rtol 1e-6 vs the sequential
    SparseLengthsSum), outputs vs the oracle."""
    for name in ("dlrm_cat", "din", "ncf", "wnd", "dlrm_rm3_like", "mtwnd"):
        mb = 24 if name == "dlrm_rm3_like" else 96
        c = make_case(name, num_batches=2, max_batch=mb)
        spec = c.spec
        S = shard.slice_size(mb, world)
        eng = AccelEngine(args=c.args, device=local, max_batch=S, precision=precision,
                          max_lookups_per_table=mb * max(1, spec.L), shard_mode="row",
                          shard_world=world)
        shard.connect(eng, dist)
        for t in range(spec.T):
            eng.load_table(t, c.tables[t])                 # full table in, slice kept
        for (mlp_id, layer), (W, b) in c.fcs.items():
            eng.load_fc(mlp_id, layer, W, b)
        for j in range(2):
            dense, ids, lengths = request(c, j, mb)
            eng.stage_batch(j, dense, ids, lengths)
        dist.barrier()
        for rep in range(2):
            for batch_id, B in [(0, mb), (1, mb - 5), (0, 3)]:
                dense, ids, lengths = request(c, batch_id, B)
                st = O.forward(spec, c.tables, c.fcs, dense, ids, lengths)
                row0, rows = shard.sample_slice(B, rank, world)
                out, _ = eng.infer_staged(batch_id, B)
                np.testing.assert_allclose(out[:rows], st["out"][row0:row0 + rows], rtol=1e-4,
                                           atol=2e-5)
                # integer-exact bucketing
                off = eng.fetch_stage(_lib.STAGE_OFFSETS)
                for t in range(spec.T):
                    r0, r1 = shard.row_range(spec.table_rows[t], rank, world)
                    lid, llen = shard.bucket_rows(ids[t], lengths[t], r0, r1)
                    want_off = np.concatenate([[0], np.cumsum(llen)]).astype(np.int32)
                    assert np.array_equal(off[t], want_off), "%s: offsets table %d" % (name, t)
                    got = eng.fetch_stage(_lib.STAGE_IDS, t).ravel()
                    assert np.array_equal(got, lid), "%s: bucketed ids table %d" % (name, t)
                if rows > 0:
                    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
                    emu = np.concatenate(
                        [O.sparse_lengths_sum_rowwise(c.tables[t], ids[t], lengths[t], world)
                         for t in range(spec.T)], axis=1)[row0:row0 + rows]
                    assert np.array_equal(pooled.view(np.uint32), emu.view(np.uint32)), \
                        "%s: row-wise pooled slice != rank-order emulation on rank %d" % (name, rank)
                    np.testing.assert_allclose(pooled, st["pooled"][row0:row0 + rows], rtol=1e-5,
                                               atol=1e-6)
                if rep == 0 and B <= S:
                    out_h, _ = eng.infer(dense, ids, lengths, B)
                    np.testing.assert_allclose(out_h[:rows], st["out"][row0:row0 + rows],
                                               rtol=1e-4, atol=2e-5)
                full = shard.assemble_outputs(dist, out[:rows], B, spec.out_dim)
                np.testing.assert_allclose(full, st["out"], rtol=1e-4, atol=2e-5)
        dist.barrier()
        eng.close()
        if rank == 0:
            print("row-sharded %s ok on %d GPUs (%s)" % (name, world, precision))
    dist.destroy_process_group()
    print("RANK %d DONE" % rank)


if __name__ == "__main__":
    main()
