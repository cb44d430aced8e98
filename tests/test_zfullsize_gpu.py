"""Full-size parity (RMC1 exactly as shipped: 8 tables x 4,000,000 rows x 32, L = 80,
16,384-sample queries — bench.py's workload) through properties that do not need the oracle to
walk 10 million lookups:

  * checksum of every pooled element: table rows hold small dyadic values
    (h(row) + 512 * (column + 32 * table)) * 2^-23, so every 80-row sum is exact in fp32 whatever
    the order, and numpy integer arithmetic on the ids predicts all 16,384 x 8 x 32 pooled
    values bit for bit — a wrong row, column, table, segment boundary or a dropped/duplicated
    lookup changes the sum;
  * partition invariance: 1,024-sample slices of the query served on their own (host path)
    give the same pooled bits and the same outputs as the rows of the whole 16,384-sample query
    (staged path);
  * sampled oracle: for 64 random samples the numpy oracle (on compact copies of just the rows
    they touch) reproduces pooled vectors bit-exactly and outputs within atol 1e-5.

Runs last in the GPU tier (file name) because it moves 4 GB of tables.
"""
import os

import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
from deeprecsys_b200.engine import AccelEngine
from deeprecsys_b200.spec import ModelSpec, draw_fc
from deeprecsys_b200.utils.utils import cli
from oracle import numpy_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RM1 = os.path.join(ROOT, "deeprecsys_b200", "models", "configs", "dlrm_rm1.json")
B = 16384
SCALE = np.float32(2.0 ** -23)


def row_hash(ids):
    """9-bit value of a row id (same for every table; the table enters through the column term)."""
    return ((ids.astype(np.uint64) * np.uint64(2654435761)) >> np.uint64(7)) & np.uint64(511)


def table_values(rows, t, D):
    """float32 [len(rows), D]: (h(row) + 512 * (c + 32 t)) * 2^-23 — numerators < 2^17, so a sum
    of 80 of them stays below 2^24 and is exact in fp32 in any order."""
    num = row_hash(rows)[:, None] + np.uint64(512) * (np.arange(D, dtype=np.uint64) + np.uint64(32 * t))
    return num.astype(np.float32) * SCALE


@pytest.fixture(scope="module")
def rm1():
    args = cli(["--config_file", RM1, "--num_batches", "1", "--max_mini_batch_size", str(B)])
    spec = ModelSpec(args)
    assert (spec.T, spec.D, spec.L) == (8, 32, 80) and spec.table_rows[0] == 4000000
    np.random.seed(123)
    _, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    fcs = {}
    for mlp_id, ln in spec.mlps:
        for layer in range(len(ln) - 1):
            fcs[(mlp_id, layer)] = draw_fc(ln[layer], ln[layer + 1])
    eng = AccelEngine(args=args, max_batch=B)
    all_rows = np.arange(spec.table_rows[0], dtype=np.int64)
    for t in range(spec.T):
        eng.load_table(t, table_values(all_rows, t, spec.D))
    for (mlp_id, layer), (W, b) in fcs.items():
        eng.load_fc(mlp_id, layer, W, b)
    eng.stage_batch(0, lX[0], [lS_i[0][t] for t in range(spec.T)],
                    [lS_l[0][t] for t in range(spec.T)])
    out, _ = eng.infer_staged(0, B)
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    yield {"spec": spec, "eng": eng, "fcs": fcs, "dense": lX[0], "ids": lS_i[0],
           "lengths": lS_l[0], "out": out, "pooled": pooled}
    eng.close()


def test_every_pooled_element_matches_the_integer_checksum(rm1):
    spec, ids = rm1["spec"], rm1["ids"]
    T, D, L = spec.T, spec.D, spec.L
    assert rm1["pooled"].shape == (B, T * D)
    h_sum = row_hash(ids).reshape(T, B, L).sum(axis=2)                        # [T, B] exact ints
    col = np.uint64(512 * L) * (np.arange(D, dtype=np.uint64)[None, None, :]
                                + np.uint64(32) * np.arange(T, dtype=np.uint64)[:, None, None])
    want = (h_sum[:, :, None] + col).astype(np.float32) * SCALE              # [T, B, D]
    want = np.ascontiguousarray(want.transpose(1, 0, 2)).reshape(B, T * D)
    assert int(h_sum.max()) + 512 * L * (D + 32 * T) < 2 ** 24               # exactness premise
    assert np.array_equal(rm1["pooled"].view(np.uint32), want.view(np.uint32))
    assert np.isfinite(rm1["out"]).all() and rm1["out"].shape == (B, 1)
    assert 0.0 <= float(rm1["out"].min()) and float(rm1["out"].max()) <= 1.0  # sigmoid
    assert float(rm1["out"].std()) > 1e-4                                     # not a constant


@pytest.mark.parametrize("start", [0, 5 * 1024, 15 * 1024])
def test_slices_served_alone_equal_the_rows_of_the_whole_query(rm1, start):
    spec, eng, L = rm1["spec"], rm1["eng"], rm1["spec"].L
    n = 1024
    ids = [np.ascontiguousarray(rm1["ids"][t][start * L:(start + n) * L]) for t in range(spec.T)]
    lengths = [np.ascontiguousarray(rm1["lengths"][t][start:start + n]) for t in range(spec.T)]
    out, _ = eng.infer(np.ascontiguousarray(rm1["dense"][start:start + n]), ids, lengths, n)
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    assert np.array_equal(pooled.view(np.uint32),
                          rm1["pooled"][start:start + n].view(np.uint32))
    np.testing.assert_allclose(out, rm1["out"][start:start + n], rtol=0, atol=1e-6)


@pytest.mark.parametrize("wide", [False, True], ids=["int32", "int64"])
def test_whole_query_through_the_host_packed_upload_equals_the_staged_one(rm1, wide):
    """drs_set_option "host_pack": all 10.5 M lookups of the 16,384-sample query travel as 22-bit
    fields in lane-interleaved blocks (8 host threads, many chunks per table) and are widened on
    the GPU — same ids, same pooled bits, same scores as the staged int32 path."""
    spec, eng = rm1["spec"], rm1["eng"]
    dtype = np.int64 if wide else np.int32
    ids = [np.ascontiguousarray(rm1["ids"][t], dtype=dtype) for t in range(spec.T)]
    lengths = [np.ascontiguousarray(rm1["lengths"][t]) for t in range(spec.T)]
    eng.set_option("host_pack", 8)
    try:
        l0 = eng.launch_count()
        out, _ = eng.infer(np.ascontiguousarray(rm1["dense"]), ids, lengths, B)
        pooled = eng.fetch_stage(_lib.STAGE_POOLED)
        assert eng.launch_count() - l0 >= spec.T           # one widening kernel per table ran
        for t in (0, spec.T - 1):
            assert np.array_equal(eng.fetch_stage(_lib.STAGE_IDS, t).ravel(), rm1["ids"][t])
    finally:
        eng.set_option("host_pack", 0)
    assert np.array_equal(pooled.view(np.uint32), rm1["pooled"].view(np.uint32))
    np.testing.assert_allclose(out, rm1["out"], rtol=0, atol=1e-6)


def test_sampled_rows_against_the_oracle(rm1):
    spec, L = rm1["spec"], rm1["spec"].L
    pick = np.sort(np.random.RandomState(5).choice(B, size=64, replace=False))
    ids, small_tables = [], []
    for t in range(spec.T):
        seg = rm1["ids"][t].reshape(B, L)[pick].reshape(-1)
        uniq, inv = np.unique(seg, return_inverse=True)
        small_tables.append(table_values(uniq.astype(np.int64), t, spec.D))
        ids.append(inv.astype(np.int64))
    lengths = [np.ascontiguousarray(rm1["lengths"][t][pick]) for t in range(spec.T)]
    st = O.forward(spec, small_tables, rm1["fcs"], rm1["dense"][pick], ids, lengths)
    assert np.array_equal(st["pooled"].view(np.uint32), rm1["pooled"][pick].view(np.uint32))
    np.testing.assert_allclose(rm1["out"][pick], st["out"], rtol=1e-4, atol=1e-5)
