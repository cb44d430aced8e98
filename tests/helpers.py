"""Shared builders for the parity tests: a small seeded case per model graph,
generated exactly the way an engine process does it (seed -> datagen inputs ->
datagen targets -> parameters)."""
import os

import numpy as np

from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator, slice_request
from deeprecsys_b200.spec import ModelSpec, draw_all_parameters
from deeprecsys_b200.utils.utils import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_DIR = os.path.join(ROOT, "deeprecsys_b200", "models", "configs")

SMALL = {
    # name -> argv  (small tables so the oracle finishes in seconds)
    "dlrm_cat": ["--arch_embedding_size", "5000-6000-7000-8000-5000-6000-7000-8000",
                 "--arch_sparse_feature_size", "32", "--arch_mlp_bot", "128-64-32",
                 "--arch_mlp_top", "256-64-1", "--arch_interaction_op", "cat",
                 "--num_indices_per_lookup", "80"],
    "dlrm_dot": ["--arch_embedding_size", "5000-6000-7000-8000-5000-6000-7000-8000",
                 "--arch_sparse_feature_size", "32", "--arch_mlp_bot", "128-64-32",
                 "--arch_mlp_top", "256-64-1", "--arch_interaction_op", "dot",
                 "--num_indices_per_lookup", "80"],
    "dlrm_dot_itself": ["--arch_embedding_size", "500-600-700", "--arch_sparse_feature_size", "16",
                        "--arch_mlp_bot", "13-16", "--arch_mlp_top", "32-1",
                        "--arch_interaction_op", "dot", "--arch_interaction_itself",
                        "--num_indices_per_lookup", "5"],
    "dlrm_rm2_like": ["--arch_embedding_size", "-".join(["3000"] * 32),
                      "--arch_sparse_feature_size", "64", "--arch_mlp_bot", "256-128-64",
                      "--arch_mlp_top", "128-64-1", "--arch_interaction_op", "cat",
                      "--num_indices_per_lookup", "120"],
    "dlrm_rm3_like": ["--arch_embedding_size", "-".join(["4000"] * 10),
                      "--arch_sparse_feature_size", "32", "--arch_mlp_bot", "2560-1024-256-32",
                      "--arch_mlp_top", "512-256-1", "--arch_interaction_op", "cat",
                      "--num_indices_per_lookup", "20"],
    "dlrm_odd": ["--arch_embedding_size", "50-60-70", "--arch_sparse_feature_size", "6",
                 "--arch_mlp_bot", "7-9-6", "--arch_mlp_top", "11-5-1",
                 "--arch_interaction_op", "dot", "--num_indices_per_lookup", "3"],
    "wnd": ["--model_type", "wnd", "--arch_embedding_size", "-".join(["2000"] * 27),
            "--arch_sparse_feature_size", "32", "--arch_mlp_bot", "512",
            "--arch_mlp_top", "1024-512-256-1", "--arch_interaction_op", "cat",
            "--num_indices_per_lookup", "1"],
    "ncf": ["--model_type", "ncf", "--arch_embedding_size", "1400-1400-280-280",
            "--arch_sparse_feature_size", "64", "--arch_mlp_bot", "512",
            "--arch_mlp_top", "256-256-128-64-64", "--arch_interaction_op", "cat",
            "--num_indices_per_lookup", "1"],
    "din": ["--model_type", "din", "--arch_embedding_size", "1000-100-2000-2000",
            "--user_behavior_tables", "0", "--arch_sparse_feature_size", "32",
            "--arch_mlp_bot", "1", "--arch_mlp_top", "200-80-2", "--arch_interaction_op", "cat",
            "--num_indices_per_lookup", "3"],
    "mtwnd": ["--model_type", "mtwnd", "--arch_embedding_size", "-".join(["300"] * 5),
              "--arch_sparse_feature_size", "16", "--arch_mlp_bot", "24",
              "--arch_mlp_top", "64-32", "--arch_mlp_tasks", "32-16-1",
              "--arch_interaction_op", "cat", "--num_indices_per_lookup", "1",
              "--num_multi_tasks", "3"],
    "dien": ["--model_type", "dien", "--arch_embedding_size",
             "300-" + "-".join(["200"] * 5) + "-400-500", "--arch_sparse_feature_size", "16",
             "--arch_mlp_bot", "24", "--arch_mlp_top", "40-20-2", "--arch_interaction_op", "cat",
             "--num_indices_per_lookup", "1", "--hidden_size", "32"],
    "din_many": ["--model_type", "din", "--arch_embedding_size", "1000-100-2000-2000",
                 "--user_behavior_tables", "10", "--arch_sparse_feature_size", "32",
                 "--arch_mlp_bot", "8-4", "--arch_mlp_top", "200-80-2",
                 "--arch_interaction_op", "cat", "--num_indices_per_lookup", "3"],
    # DIN as the reference's cli() expands it when din.json's values arrive as FLAGS rather than
    # through the JSON (utils/utils.py:132-149): profile + 251 behaviour tables + ad + context
    # = 254 tables, 251 attention units [3D, 1, D] ("pure-CLI variant")
    "din254": ["--model_type", "din", "--arch_embedding_size", "1000-100-2000-2000",
               "--user_behavior_tables", "250", "--arch_sparse_feature_size", "32",
               "--arch_mlp_bot", "1", "--arch_mlp_top", "200-80-2",
               "--arch_interaction_op", "cat", "--num_indices_per_lookup", "3"],
}


class Case(object):
    pass


def make_case(name, num_batches=2, max_batch=64, seed=123, extra=()):
    argv = list(SMALL[name]) + ["--num_batches", str(num_batches),
                                "--max_mini_batch_size", str(max_batch),
                                "--numpy_rand_seed", str(seed)] + list(extra)
    c = Case()
    c.name = name
    c.args = cli(argv)                                   # seeds numpy
    gen = DLRMDataGenerator(c.args)
    _, c.lX, c.lS_l, c.lS_i = gen.generate_input_data()
    _, c.lT = gen.generate_output_data()
    c.spec = ModelSpec(c.args)
    c.tables, c.fcs = draw_all_parameters(c.spec)
    return c


def request(c, batch_id, batch_size):
    d, l, i = slice_request(c.lX, c.lS_l, c.lS_i, batch_id, batch_size, c.spec.L)
    dense = d if c.spec.m_den > 0 else None
    return dense, [np.ascontiguousarray(x) for x in i], [np.ascontiguousarray(x) for x in l]


# ---------------------------------------------------------------------------------------------
# Golden vectors recorded from the REFERENCE's own model code (tests/golden/make_ref_models.py:
# the unmodified /root/reference/models/*.py builders driven like inferenceEngine.py drives
# them, over oracle/caffe2_shim).  Blob names are the reference's tags.
GOLDEN_MAX_BATCH = 12
GOLDEN_REQUESTS = [(1, 7), (0, 12), (1, 1)]


def load_ref_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))


# cases recorded through the reference's standalone path (whole batches, no queues): the
# reference's MT-WnD queue mode is broken (multi_task_wnd.py:111)
GOLDEN_STANDALONE = {"mtwnd": [(0, GOLDEN_MAX_BATCH), (1, GOLDEN_MAX_BATCH)]}


def golden_requests(name):
    return GOLDEN_STANDALONE.get(name, GOLDEN_REQUESTS)


def apply_golden_params(c, z):
    """DIEN: the recurrent cells' weights are created by the framework's initialiser when the
    reference runs model.param_init_net (not from the numpy stream), so the golden file records
    them; use exactly those for the comparison.  No-op for every other graph."""
    if c.spec.model_type != "dien":
        return c
    from deeprecsys_b200 import _lib
    for cell in range(2):
        pre = "%s/params/rnn_%d/" % (c.name, cell)
        c.fcs[(_lib.MLP_RNN_BASE + cell, 0)] = (z[pre + "i2h_w"], z[pre + "i2h_b"])
        c.fcs[(_lib.MLP_RNN_BASE + cell, 1)] = (z[pre + "gates_t_w"], z[pre + "gates_t_b"])
    return c


def golden_stages(z, name, batch_id, B, spec):
    """Map the reference's blob names to the stage dict of oracle.numpy_oracle.forward."""
    tag = name + "@standalone" if name in GOLDEN_STANDALONE else name
    pre = "%s/req_%d_%d/" % (tag, batch_id, B)

    def g(tag):
        return z[pre + tag]

    def mlp_tags(layer_tag, n, out_tag):
        return [g("%s:::fc%d_z" % (layer_tag, i)) if i < n else g(out_tag) for i in range(1, n + 1)]

    T = spec.T
    if spec.model_type == "ncf":                       # ncf.py:208-297 table tags
        pooled = [g("emb:::mf_sls0_z"), g("emb:::mf_sls1_z"), g("emb:::mlp_sls0_z"),
                  g("emb:::mlp_sls1_z")]
    else:
        pooled = [g("emb:::sls%d_z" % t) for t in range(T)]
    st = {"pooled": np.concatenate(pooled, axis=1), "bot": [], "top": []}
    mlps = dict(spec.mlps)
    if spec.model_type == "dlrm":
        st["bot"] = mlp_tags("bot", len(mlps[0]) - 1, "dense_out")
        st["interact"] = g("interaction")
        st["top"] = mlp_tags("top", len(mlps[1]) - 1, "prob_click")
    elif spec.model_type == "wnd":
        st["interact"] = g("interaction")
        st["top"] = mlp_tags("top", len(mlps[1]) - 1, "prob_click")
    elif spec.model_type == "mtwnd":                   # multi_task_wnd.py:407-420 tags
        st["interact"] = g("interaction")
        st["top"] = mlp_tags("top", len(mlps[1]) - 1, "shared_dense")
        n_task_layers = len(spec.ln_task) - 1
        st["tasks"] = [mlp_tags("task_%d" % k, n_task_layers, "task_outs_%d" % k)
                       for k in range(spec.n_tasks)]
        st["out"] = np.concatenate([t[-1] for t in st["tasks"]], axis=1)
        return st
    elif spec.model_type == "ncf":
        st["top"] = mlp_tags("mlpfc", len(mlps[1]) - 1, "mlpfc_out")
        st["interact"] = g("feat_int")
        st["out"] = g("prob_click")
        return st
    elif spec.model_type == "dien":
        st["bot"] = [g("rnn_1/hidden_t_last").reshape(B, -1)]
        st["interact"] = g("gru:::concat")
        st["top"] = mlp_tags("top", len(mlps[1]) - 1, "prob_click")
    elif spec.model_type == "din":
        st["bot"] = [g("atten_out")]
        st["interact"] = g("top_fc_in")
        st["top"] = mlp_tags("top", len(mlps[1]) - 1, "prob_click")
    else:
        raise ValueError(spec.model_type)
    st["out"] = g("prob_click")
    return st
