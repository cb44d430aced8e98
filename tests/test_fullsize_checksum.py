"""CPU tier: the integer checksum that tests/test_zfullsize_gpu.py uses to predict every pooled
element of a full-size RMC1 query is itself pinned against the oracle (small batch, compact copies
of the touched rows; same helper functions)."""
import numpy as np

from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
from deeprecsys_b200.spec import ModelSpec
from deeprecsys_b200.utils.utils import cli
from oracle import numpy_oracle as O
from tests.test_zfullsize_gpu import RM1, SCALE, row_hash, table_values


def test_checksum_formula_equals_the_sequential_oracle_bitwise():
    Bs = 96
    args = cli(["--config_file", RM1, "--num_batches", "1", "--max_mini_batch_size", str(Bs)])
    spec = ModelSpec(args)
    T, D, L = spec.T, spec.D, spec.L
    _, _, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    ids = lS_i[0]
    want = []
    for t in range(T):
        uniq, inv = np.unique(ids[t], return_inverse=True)
        want.append(O.sparse_lengths_sum(table_values(uniq.astype(np.int64), t, D),
                                         inv.astype(np.int64), lS_l[0][t]))
    want = np.concatenate(want, axis=1)
    h_sum = row_hash(ids).reshape(T, Bs, L).sum(axis=2)
    col = np.uint64(512 * L) * (np.arange(D, dtype=np.uint64)[None, None, :]
                                + np.uint64(32) * np.arange(T, dtype=np.uint64)[:, None, None])
    got = (h_sum[:, :, None] + col).astype(np.float32) * SCALE
    got = np.ascontiguousarray(got.transpose(1, 0, 2)).reshape(Bs, T * D)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the values tell rows, columns and tables apart
    a, b = table_values(np.array([7, 8]), 0, D), table_values(np.array([7, 8]), 1, D)
    assert len(np.unique(np.concatenate([a, b]).round(9))) >= 2 * D
    assert not np.array_equal(a[0], a[1]) or row_hash(np.array([7])) == row_hash(np.array([8]))


def test_generalised_checksum_equals_the_sequential_oracle_bitwise():
    """tests/fullsize_util.py (RMC2 / WnD / DIN-254 shapes): the integer prediction of every
    pooled element equals the oracle's sequential fp32 SparseLengthsSum on the same values."""
    import os
    from tests import fullsize_util as U
    from tests.test_zfullsize_more_gpu import CASES
    for name in ("rm2", "wnd", "din254"):
        argv, _, _ = CASES[name]
        Bs = 40
        args = cli(argv + ["--num_batches", "1", "--max_mini_batch_size", str(Bs)])
        spec = ModelSpec(args)
        T, D, L = spec.T, spec.D, spec.L
        hb = U.hash_bits(T, D, L)
        assert L * ((1 << hb) * T * D) < 2 ** 24
        _, _, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
        pick = np.arange(Bs)
        tables, ids, lengths = U.compact_tables(lS_i[0], lS_l[0], pick, spec, hb)
        want = np.concatenate([O.sparse_lengths_sum(W, i, l)
                               for W, i, l in zip(tables, ids, lengths)], axis=1)
        got = U.expected_pooled(lS_i[0], T, Bs, D, L, hb)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    assert U.hash_bits(8, 32, 80) == 9           # the RMC1 instance of tests/test_zfullsize_gpu.py
