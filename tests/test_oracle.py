"""Pin the CPU oracle: known answers, two independent
implementations agreeing bit-for-bit on SparseLengthsSum, and the numpy and
torch restatements agreeing on every model graph."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deeprecsys_b200.spec import ModelSpec
from deeprecsys_b200.utils.utils import cli
from oracle import numpy_oracle as O
from oracle.torch_cpu_engine import TorchCpuModel
from tests.helpers import SMALL, make_case, request


def test_tril_indices_known_answers():
    # hand-computed: F=3, offset 0 -> [3, 6, 7]; itself -> [0, 3, 4, 6, 7, 8]
    assert O.tril_indices(3, False).tolist() == [3, 6, 7]
    assert O.tril_indices(3, True).tolist() == [0, 3, 4, 6, 7, 8]
    assert len(O.tril_indices(9, False)) == 36


def test_lengths_to_offsets_known_answers():
    assert O.lengths_to_offsets([3, 0, 2, 5]).tolist() == [0, 3, 3, 5, 10]
    assert O.lengths_to_offsets([]).tolist() == [0]
    with pytest.raises(ValueError):
        O.lengths_to_offsets([1, -1])


def test_sls_known_answer_and_enforces():
    W = np.arange(12, dtype=np.float32).reshape(4, 3)
    out = O.sparse_lengths_sum(W, [0, 3, 1, 1, 2], [2, 0, 3])
    assert out.tolist() == [[9, 11, 13], [0, 0, 0], [12, 15, 18]]
    with pytest.raises(ValueError):
        O.sparse_lengths_sum(W, [0, 4], [2])            # idx out of range
    with pytest.raises(ValueError):
        O.sparse_lengths_sum(W, [0, 1, 2], [2])         # sum(lengths) != len(indices)


def test_sls_sequential_fp32_equals_embedding_bag_bitwise():
    """The property that makes a bit-exact GPU target possible."""
    rng = np.random.RandomState(0)
    for D in (32, 64, 12):
        W = rng.uniform(-5e-4, 5e-4, size=(5000, D)).astype(np.float32)
        lengths = rng.randint(0, 100, size=300).astype(np.int32)
        idx = rng.randint(0, 5000, size=int(lengths.sum())).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
        ref = F.embedding_bag(torch.from_numpy(idx), torch.from_numpy(W), torch.from_numpy(off),
                              mode="sum").numpy()
        got = O.sparse_lengths_sum(W, idx, lengths)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        # and equals an explicit python loop on a few segments
        for b in (0, 7, 299):
            acc = np.zeros(D, np.float32)
            for k in range(off[b], off[b] + lengths[b]):
                acc = acc + W[idx[k]]
            assert np.array_equal(acc, got[b])


def test_num_int_table_for_shipped_configs():
    import os
    cfg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "deeprecsys_b200", "models", "configs")
    want = {"dlrm_rm1.json": ([128, 64, 32], [288, 256, 64, 1], 1),
            "dlrm_rm2.json": ([256, 128, 64], [2112, 128, 64, 1], 1),
            "dlrm_rm3.json": ([2560, 1024, 256, 32], [352, 512, 256, 1], 1),
            "wide_and_deep.json": ([512], [1376, 1024, 512, 256, 1], 1),
            "ncf.json": ([], [128, 256, 256, 128, 64, 64], 64),
            "din.json": ([96, 1, 32], [128, 200, 80, 2], 2)}
    for name, (bot, top, out_dim) in want.items():
        s = ModelSpec(cli(["--config_file", os.path.join(cfg, name)]))
        assert s.ln_bot == bot and s.ln_top == top and s.out_dim == out_dim, name
    s = ModelSpec(cli(["--config_file", os.path.join(cfg, "dlrm_rm1.json")]))
    assert s.gather_bytes_per_sample(fused=True) == 84512
    assert s.gather_bytes_per_sample(fused=False) == 85536
    assert s.mlp_flops_per_sample() == 200832
    # dot variant of rm1: 9 features -> 36 pairs + 32
    a = cli(["--config_file", os.path.join(cfg, "dlrm_rm1.json")])
    a.arch_interaction_op = "dot"
    assert ModelSpec(a).ln_top[0] == 68
    a.arch_interaction_itself = True
    assert ModelSpec(a).ln_top[0] == 77


def test_spec_rejects_what_the_reference_rejects():
    with pytest.raises(ValueError):     # m_spa != ln_bot[-1]  (dlrm_s_caffe2.py:435)
        ModelSpec(cli(["--arch_sparse_feature_size", "8", "--arch_mlp_bot", "4-3",
                       "--arch_embedding_size", "5-5"]))
    with pytest.raises(ValueError):     # unsupported interaction
        ModelSpec(cli(["--arch_interaction_op", "sum"]))
    with pytest.raises(ValueError):     # NCF needs 4 tables
        ModelSpec(cli(["--model_type", "ncf", "--arch_embedding_size", "5-5-5",
                       "--arch_interaction_op", "cat", "--num_indices_per_lookup", "1"]))
    with pytest.raises(ValueError):     # not on the accel path
        ModelSpec(cli(["--model_type", "dien"]))


@pytest.mark.parametrize("name", sorted(SMALL))
def test_numpy_and_torch_restatements_agree(name):
    mb = 8 if name == "dlrm_rm3_like" else 24
    c = make_case(name, num_batches=1, max_batch=mb)
    dense, ids, lengths = request(c, 0, mb)
    st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
    tm = TorchCpuModel(c.spec, c.tables, c.fcs)
    out = tm.forward(dense, ids, lengths).numpy()
    np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=2e-6)
    pooled = torch.cat(tm.pool(ids, lengths), dim=1).numpy()
    assert np.array_equal(pooled.view(np.uint32), st["pooled"].view(np.uint32))
    assert st["out"].shape == (mb, c.spec.out_dim)


def test_parameter_draw_order_and_distribution():
    """tables first (uniform +-sqrt(1/n), float64 draw then cast), then per layer W then b."""
    c = make_case("dlrm_cat", num_batches=1, max_batch=4, seed=5)
    # replay the stream by hand
    from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
    args = cli(SMALL["dlrm_cat"] + ["--num_batches", "1", "--max_mini_batch_size", "4",
                                    "--numpy_rand_seed", "5"])
    g = DLRMDataGenerator(args)
    g.generate_input_data()
    g.generate_output_data()
    n = c.spec.table_rows[0]
    W0 = np.random.uniform(low=-np.sqrt(1 / n), high=np.sqrt(1 / n), size=(n, 32)).astype(np.float32)
    assert np.array_equal(W0, c.tables[0])
    assert abs(c.tables[0]).max() <= np.sqrt(1 / n)
    for t in range(1, 8):
        n = c.spec.table_rows[t]
        np.random.uniform(size=(n, 32))
    Wb, bb = c.fcs[(0, 0)]
    W = np.random.normal(0.0, np.sqrt(2 / (64 + 128)), size=(64, 128)).astype(np.float32)
    b = np.random.normal(0.0, np.sqrt(1 / 64), size=64).astype(np.float32)
    assert np.array_equal(W, Wb) and np.array_equal(b, bb)
