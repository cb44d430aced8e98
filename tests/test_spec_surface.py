"""Spec surface: cli(), JSON configs, packets and the data
generator against golden fixtures produced from the reference's own python
(tests/golden/make_golden.py), and — when /root/reference is mounted (authoring
container only) — against the reference modules directly."""
import json
import os
import sys

import numpy as np
import pytest

from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator, slice_request
from deeprecsys_b200.utils import packets
from deeprecsys_b200.utils.utils import REFERENCE_FLAGS, cli, expand_din_tables

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CFG = os.path.join(os.path.dirname(HERE), "deeprecsys_b200", "models", "configs")
REF = "/root/reference"


def _argv(argv):
    return [a.replace("<CONFIGS>", CFG) for a in argv]


with open(os.path.join(GOLD, "cli_namespaces.json")) as f:
    CLI_GOLD = json.load(f)


@pytest.mark.parametrize("name", sorted(CLI_GOLD))
def test_cli_namespace_matches_reference_golden(name):
    case = CLI_GOLD[name]
    got = vars(cli(_argv(case["argv"])))
    want = {k: (v.replace("<CONFIGS>", CFG) if isinstance(v, str) else v)
            for k, v in case["namespace"].items()}
    for key, val in want.items():
        assert key in got, key
        assert got[key] == val and type(got[key]) is type(val), (key, got[key], val)
    # every reference flag exists; our additions keep old behaviour at their defaults
    assert set(want) == {n for n, _, _ in REFERENCE_FLAGS}
    assert got["accel_precision"] == "fp32" and got["accel_shard"] == "replica"


def test_cli_din_expansion_order():
    # CLI-only: 250 behaviour tables are *prepended* -> 254 tables
    a = cli(["--model_type", "din", "--arch_embedding_size", "1000-100-2000-3000",
             "--user_behavior_tables", "250"])
    sizes = a.arch_embedding_size.split("-")
    assert len(sizes) == 254 and sizes[0] == "1000" and sizes[1:252] == ["100"] * 251
    assert sizes[-2:] == ["2000", "3000"]
    # JSON wins over the expansion: din.json -> 4 tables
    b = cli(["--config_file", os.path.join(CFG, "din.json")])
    assert len(b.arch_embedding_size.split("-")) == 4
    assert expand_din_tables("5-6-7-8", 0) == "5-6-7-8"


def test_cli_json_unknown_key_raises(tmp_path):
    p = tmp_path / "bad.json"
    p.write_text(json.dumps({"no_such_flag": 1}))
    with pytest.raises(AttributeError):
        cli(["--config_file", str(p)])


def test_cli_seeds_numpy():
    cli(["--numpy_rand_seed", "77"])
    a = np.random.rand(3)
    np.random.seed(77)
    assert np.array_equal(a, np.random.rand(3))


def test_packets_attribute_surface():
    with open(os.path.join(GOLD, "packets.json")) as f:
        gold = json.load(f)
    assert sorted(vars(packets.ServiceRequest())) == gold["ServiceRequest"]
    assert sorted(vars(packets.ServiceResponse())) == gold["ServiceResponse"]
    r = packets.ServiceResponse(process_start_time=1.5, inference_end_time=2.0)
    assert r.queue_start_time == 1.5 and not hasattr(r, "process_start_time")
    str(r), str(packets.ServiceRequest(batch_id=1, epoch=0, batch_size=4, arrival_time=0.1))


def test_config_files_equal_reference_configs():
    if not os.path.isdir(REF):
        pytest.skip("reference not mounted")
    for name in os.listdir(os.path.join(REF, "models", "configs")):
        with open(os.path.join(REF, "models", "configs", name)) as f:
            ref = json.load(f)
        with open(os.path.join(CFG, name)) as f:
            ours = json.load(f)
        assert ref == ours, name


# ------------------------------------------------------------------ datagen
DG = np.load(os.path.join(GOLD, "datagen_cases.npz"))
DG_CASES = sorted({k.split("/")[0] for k in DG.files})


@pytest.mark.parametrize("name", DG_CASES)
def test_datagen_bit_identical_to_reference_golden(name):
    argv = [str(a) for a in DG[name + "/argv"]]
    args = cli(argv)                              # seeds numpy
    g = DLRMDataGenerator(args)
    nb, lX, lS_l, lS_i = g.generate_input_data()
    _, lT = g.generate_output_data()
    for j in range(nb):
        assert lX[j].dtype == np.float32 and np.array_equal(lX[j], DG["%s/X%d" % (name, j)])
        assert np.array_equal(lS_l[j], DG["%s/len%d" % (name, j)])
        assert np.array_equal(lS_i[j], DG["%s/idx%d" % (name, j)])
        assert np.array_equal(lT[j], DG["%s/T%d" % (name, j)])
        # ids sorted-unique inside each sample (np.unique in the reference)
        L = args.num_indices_per_lookup
        if L > 1:
            per = lS_i[j].reshape(lS_i[j].shape[0], -1, L)
            assert (np.diff(per, axis=2) > 0).all()
    # RNG left in the same state -> the parameters drawn next are the same
    assert np.array_equal(np.random.random_sample(4), DG[name + "/rng_tail"])


def test_datagen_rm1_probe_full_size_tables():
    with open(os.path.join(GOLD, "rm1_probe.json")) as f:
        probe = json.load(f)
    args = cli(_argv(probe["argv"]))
    _, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    assert [float(v) for v in lX[0][0, :4]] == probe["lX0_row0_first4"]
    assert [int(v) for v in lS_i[0][0][:8]] == probe["idx_table0_first8"]
    assert [int(v) for v in lS_i[0][7][-4:]] == probe["idx_table7_last4"]
    assert lS_l[0].dtype == np.int32 and (lS_l[0] == 80).all()


def test_datagen_against_reference_module_directly():
    if not os.path.isdir(REF):
        pytest.skip("reference not mounted")
    sys.path.insert(0, REF)
    try:
        from data_generator.dlrm_data_caffe2 import DLRMDataGenerator as RefGen
    finally:
        sys.path.remove(REF)
    argv = ["--arch_embedding_size", "3000-2000-400", "--arch_sparse_feature_size", "4",
            "--arch_mlp_bot", "9-4", "--arch_mlp_top", "8-1", "--num_indices_per_lookup", "30",
            "--num_batches", "3", "--max_mini_batch_size", "40", "--numpy_rand_seed", "2024"]
    args = cli(argv)
    _, X, Sl, Si = RefGen(args).generate_input_data()
    _, T = RefGen(args).generate_output_data()
    tail = np.random.rand(5)
    args = cli(argv)
    _, X2, Sl2, Si2 = DLRMDataGenerator(args).generate_input_data()
    _, T2 = DLRMDataGenerator(args).generate_output_data()
    for j in range(3):
        assert np.array_equal(X[j], X2[j]) and np.array_equal(np.array(Sl[j]), Sl2[j])
        assert np.array_equal(np.array(Si[j]), Si2[j]) and np.array_equal(T[j], T2[j])
    assert np.array_equal(tail, np.random.rand(5))


def test_slice_request_first_rows_only():
    args = cli(["--arch_embedding_size", "100-100", "--arch_sparse_feature_size", "4",
                "--arch_mlp_bot", "5-4", "--arch_mlp_top", "8-1", "--num_indices_per_lookup", "3",
                "--num_batches", "2", "--max_mini_batch_size", "10"])
    _, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    d, l, i = slice_request(lX, lS_l, lS_i, 1, 4, 3)
    assert d.shape == (4, 5) and l.shape == (2, 4) and i.shape == (2, 12)
    assert np.array_equal(i, lS_i[1][:, :12]) and np.array_equal(d, lX[1][:4])


from tests.helpers import CONFIG_DIR  # noqa: E402


def test_mtwnd_and_dien_dims_follow_the_reference_builders():
    """multi_task_wnd.py:340-362,396 and dien.py:386-433."""
    from deeprecsys_b200.spec import ModelSpec
    a = cli(["--config_file", os.path.join(CONFIG_DIR, "mtwnd.json")])
    s = ModelSpec(a)
    assert s.T == 43 and s.ln_top == [43 * 32 + 512, 1024, 512] and s.ln_task == [512, 256, 128]
    assert s.n_tasks == a.num_multi_tasks == 1
    assert s.task_sigmoid_layer == 2          # sigmoid_top = len(ln_top) - 1, applied to the heads
    assert s.out_dim == s.n_tasks * 128
    a = cli(["--model_type", "mtwnd", "--arch_embedding_size", "9-9", "--arch_sparse_feature_size", "4",
             "--arch_mlp_bot", "3", "--arch_mlp_top", "8-8-8", "--arch_mlp_tasks", "8-2-1",
             "--arch_interaction_op", "cat", "--num_multi_tasks", "2"])
    s = ModelSpec(a)
    assert s.task_sigmoid_layer == 0          # layer 3 of a 2-layer head: never reached, all ReLU
    with pytest.raises(ValueError):            # multi_task_wnd.py:362
        ModelSpec(cli(["--model_type", "mtwnd", "--arch_embedding_size", "9-9",
                       "--arch_sparse_feature_size", "4", "--arch_mlp_bot", "3",
                       "--arch_mlp_top", "8-4", "--arch_mlp_tasks", "8-1",
                       "--arch_interaction_op", "cat"]))
    a = cli(["--config_file", os.path.join(CONFIG_DIR, "dien.json")])
    s = ModelSpec(a)
    assert s.T == 43 and s.n_steps == 40 and s.hidden_size == 64
    assert s.ln_top == [64 + 3 * 32, 200, 80, 2] and s.out_dim == 2 and s.m_den == 0
    assert s.rnn == [(32, 64), (64, 64)]
