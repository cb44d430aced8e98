"""TEST INFRASTRUCTURE ONLY: an oracle-backed stand-in for deeprecsys_b200.engine.AccelEngine.

There is no GPU in the authoring container, and bench.py's host logic (the secondary block, the
sharded legs with their in-bench parity checks, the JSON line) would otherwise run for the first
time on the GPU box.  tests/test_bench_dryrun.py installs this class in place of the ctypes
engine and runs bench.run_b200 end to end on tiny tables — single process and world-size-2 over
gloo.  Never imported by the product or by bench.py itself.
"""
import time

import numpy as np

from deeprecsys_b200 import _lib, shard as shardlib
from deeprecsys_b200.spec import ModelSpec
from oracle import numpy_oracle as O

# sharded runs: a rank loads only the tables it owns; the fake asks this for the others
TABLE_PROVIDER = None


class FakeAccelEngine(object):
    instances = 0

    def __init__(self, args=None, spec=None, device=0, max_batch=None, precision=None, lanes=0,
                 max_lookups_per_table=None, shard_mode="table", shard_world=1):
        type(self).instances += 1
        self.spec = spec if spec is not None else ModelSpec(args)
        self.max_batch, self.precision = int(max_batch), precision
        self.out_dim = self.spec.out_dim
        self.shard_mode, self.rank, self.world = shard_mode, 0, 1
        self.tables, self.fcs, self.staged = {}, {}, {}
        self.launches, self.last, self.t0 = 0, None, None
        self.lanes = int(lanes) or 2
        self.prof_n, self.prof_cap = 0, 0

    # parameters / inputs
    def load_table(self, t, W):
        self.tables[int(t)] = np.array(W, dtype=np.float32)

    def load_fc(self, mlp_id, layer, W, b):
        self.fcs[(int(mlp_id), int(layer))] = (np.array(W, np.float32), np.array(b, np.float32))

    def stage_batch(self, batch_id, dense, ids, lengths, n_samples=None):
        self.staged[int(batch_id)] = (None if dense is None else np.array(dense),
                                      [np.array(i) for i in ids], [np.array(l) for l in lengths])

    # multi-GPU
    def shard_export(self):
        return bytes(_lib.SHARD_BLOB_BYTES)

    def shard_connect(self, rank, world, blobs):
        assert len(blobs) == world
        self.rank, self.world = int(rank), int(world)

    def _table(self, t):
        if t in self.tables:
            return self.tables[t]
        return TABLE_PROVIDER(self.spec, t)

    # serving
    def enqueue_staged(self, batch_id, batch_size):
        B = int(batch_size)
        assert 1 <= B <= self.max_batch * self.world, (B, self.max_batch, self.world)
        self.launches += 3
        self.last = (int(batch_id), B)
        if self.prof_n < self.prof_cap:
            self.prof_n += 1

    def enqueue_host(self, dense_ptr, ids_ptrs, is64, len_ptrs, batch_size, out_ptr):
        assert 1 <= int(batch_size) <= self.max_batch
        self.launches += 4

    def enqueue_host_packed(self, dense_ptr, packed_ptrs, bits, len_ptrs, batch_size, out_ptr):
        assert 1 <= int(batch_size) <= self.max_batch
        self.launches += 4 + self.spec.T

    def sync(self):
        pass

    def lane_count(self):
        return self.lanes

    def timer_start(self):
        self.t0 = time.perf_counter()

    def timer_stop(self):
        return max(1e-3, (time.perf_counter() - self.t0) * 1e3) + 0.5

    def profile_begin(self, max_steps):
        self.prof_cap, self.prof_n = int(max_steps), 0

    def profile_end(self):
        n = self.prof_n
        self.prof_cap = 0
        return n, 0.1 * n, 0.7 * n, 0.2 * n

    def launch_count(self):
        return self.launches

    def hbm_bytes(self):
        return sum(int(W.nbytes) for W in self.tables.values())

    def set_option(self, key, value):
        pass

    def fetch_stage(self, stage, layer=0):
        batch_id, B = self.last
        dense, ids, lengths = self.staged[batch_id]
        spec, L = self.spec, self.spec.L
        if self.world > 1:
            row0, rows = shardlib.sample_slice(B, self.rank, self.world)
        else:
            row0, rows = 0, B
        sl = slice(row0, row0 + rows)
        tables = [self._table(t) for t in range(spec.T)]
        ids_s = [np.asarray(i)[row0 * L:(row0 + rows) * L] for i in ids]
        len_s = [np.asarray(l)[sl] for l in lengths]
        st = O.forward(spec, tables, self.fcs, None if dense is None else dense[sl], ids_s, len_s)
        if stage == _lib.STAGE_OUT:
            return st["out"]
        if stage == _lib.STAGE_POOLED:
            if self.world > 1 and self.shard_mode == "row":
                return np.concatenate([O.sparse_lengths_sum_rowwise(W, i, l, self.world)
                                       for W, i, l in zip(tables, ids_s, len_s)], axis=1)
            return st["pooled"]
        raise NotImplementedError(stage)

    def close(self):
        pass


def fake_pinned_empty(shape, dtype):
    return np.empty(shape, dtype=dtype), None
