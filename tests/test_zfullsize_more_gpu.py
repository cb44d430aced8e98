"""Full-size parity beyond RMC1-cat: the kernels that round 1 only checked
on tables of a few thousand rows, at the shipped table sizes and 16,384-sample queries —

  rm2      D = 64 (LPR = 16 lanes per row), 32 tables x 500,000 rows, L = 120
  wnd      the SHORT gather instantiation (27 tables x 1,000,000 rows, L = 1) + k_fc_tc
  rm1_dot  the FUSE_DOT kernel (8 x 4,000,000 rows, L = 80, pairwise-dot interaction)
  din254   254 tables / 251 attention units (pure-CLI expansion), B = 4,096

through the integer checksum of tests/fullsize_util.py (every pooled element bit-exact), partition
invariance (slices served alone through the host path == rows of the whole query) and a sampled
comparison with the oracle (pooled bit-exact, scores atol 1e-5).  Runs after
tests/test_zfullsize_gpu.py (file name) because each case moves gigabytes of tables.
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
from tests import fullsize_util as U

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "deeprecsys_b200", "models", "configs")

CASES = {
    "rm2": (["--config_file", os.path.join(CFG, "dlrm_rm2.json")], 16384, "fp32"),
    "wnd": (["--config_file", os.path.join(CFG, "wide_and_deep.json")], 16384, "tf32x3"),
    "rm1_dot": (["--config_file", os.path.join(CFG, "dlrm_rm1.json")], 16384, "fp32"),
    "din254": (["--model_type", "din", "--arch_embedding_size", "1000000-100000-10000000-10000000",
                "--user_behavior_tables", "250", "--arch_sparse_feature_size", "32",
                "--arch_mlp_bot", "1", "--arch_mlp_top", "200-80-2", "--arch_interaction_op", "cat",
                "--num_indices_per_lookup", "3", "--num_indices_per_lookup_fixed", "True"],
               4096, "fp32"),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def served(request):
    name = request.param
    argv, B, precision = CASES[name]
    args = cli(argv + ["--num_batches", "1", "--max_mini_batch_size", str(B)])
    if name == "rm1_dot":
        args.arch_interaction_op = "dot"
    spec = ModelSpec(args)
    hb = U.hash_bits(spec.T, spec.D, spec.L)
    np.random.seed(123)
    _, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    fcs = {}
    for mlp_id, ln in spec.mlps:
        for layer in range(len(ln) - 1):
            fcs[(mlp_id, layer)] = draw_fc(ln[layer], ln[layer + 1])
    eng = AccelEngine(args=args, max_batch=B, precision=precision)
    for t in range(spec.T):
        rows = np.arange(spec.table_rows[t], dtype=np.int64)
        eng.load_table(t, U.table_values(rows, t, spec.D, hb))
    for (mlp_id, layer), (W, b) in fcs.items():
        eng.load_fc(mlp_id, layer, W, b)
    dense = lX[0] if spec.m_den > 0 else None
    eng.stage_batch(0, dense, [lS_i[0][t] for t in range(spec.T)],
                    [lS_l[0][t] for t in range(spec.T)])
    out, _ = eng.infer_staged(0, B)
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    yield {"name": name, "spec": spec, "eng": eng, "fcs": fcs, "dense": dense, "ids": lS_i[0],
           "lengths": lS_l[0], "out": out, "pooled": pooled, "B": B, "hb": hb,
           "precision": precision}
    eng.close()


def test_every_pooled_element_matches_the_integer_checksum(served):
    spec, B = served["spec"], served["B"]
    want = U.expected_pooled(served["ids"], spec.T, B, spec.D, spec.L, served["hb"])
    assert served["pooled"].shape == want.shape
    assert np.array_equal(served["pooled"].view(np.uint32), want.view(np.uint32))
    assert served["out"].shape == (B, spec.out_dim) and np.isfinite(served["out"]).all()


def test_a_slice_served_alone_equals_the_rows_of_the_whole_query(served):
    spec, eng, L, B = served["spec"], served["eng"], served["spec"].L, served["B"]
    n, start = 1024, B - 1024 - 7                     # unaligned start, reaches the last samples
    ids = [np.ascontiguousarray(served["ids"][t][start * L:(start + n) * L]) for t in range(spec.T)]
    lengths = [np.ascontiguousarray(served["lengths"][t][start:start + n]) for t in range(spec.T)]
    dense = (np.ascontiguousarray(served["dense"][start:start + n]) if spec.m_den > 0 else None)
    out, _ = eng.infer(dense, ids, lengths, n)
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    assert np.array_equal(pooled.view(np.uint32), served["pooled"][start:start + n].view(np.uint32))
    tol = 1e-6 if served["precision"] == "fp32" else 2e-5
    np.testing.assert_allclose(out, served["out"][start:start + n], rtol=0, atol=tol)


def test_sampled_rows_against_the_oracle(served):
    spec, B = served["spec"], served["B"]
    pick = np.sort(np.random.RandomState(7).choice(B, size=48, replace=False))
    tables, ids, lengths = U.compact_tables(served["ids"], served["lengths"], pick, spec,
                                            served["hb"])
    dense = served["dense"][pick] if spec.m_den > 0 else None
    st = O.forward(spec, tables, served["fcs"], dense, ids, lengths)
    assert np.array_equal(st["pooled"].view(np.uint32), served["pooled"][pick].view(np.uint32))
    atol = 1e-5 * max(1.0, float(np.abs(st["out"]).max()))
    np.testing.assert_allclose(served["out"][pick], st["out"], rtol=1e-4, atol=atol)
