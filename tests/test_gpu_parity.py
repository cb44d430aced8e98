"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Tolerances:
  index/offset bucketing ............ bit-exact (int32)
  pooled vectors (SparseLengthsSum) . bit-exact (fp32, sequential order)
  cat interaction ................... bit-exact
  dot interaction ................... rtol 1e-5 / atol 1e-6
  FC stack, fp32 FFMA path .......... atol 1e-5 on outputs (rtol 1e-4 on hidden layers)
"""
import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine
from oracle import numpy_oracle as O
from tests.helpers import SMALL, make_case, request

pytestmark = pytest.mark.gpu


def _engine(c, precision="fp32", chain=True):
    eng = AccelEngine(args=c.args, max_batch=c.args.max_mini_batch_size, precision=precision)
    eng.set_option("mlp_chain", int(chain))
    eng.set_option("tap_hidden", 1)          # stage-level parity needs the hidden activations
    eng.load_parameters(c.tables, c.fcs)
    return eng


def _check_stages(c, eng, st, B, fc_atol=1e-5, fc_rtol=1e-4):
    spec = c.spec
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    assert pooled.shape == st["pooled"].shape
    assert np.array_equal(pooled.view(np.uint32), st["pooled"].view(np.uint32)), "pooled not bit-exact"
    off = eng.fetch_stage(_lib.STAGE_OFFSETS)
    assert off.dtype == np.int32 and off.shape == (spec.T, B + 1)
    inter = eng.fetch_stage(_lib.STAGE_INTERACT)
    assert inter.shape == st["interact"].shape
    if spec.model_type in ("dlrm",) and spec.interaction_op == "cat":
        # embedding columns are copies of pooled: exact; dense block is FC output
        D = spec.D
        assert np.array_equal(inter[:, D:], st["interact"][:, D:])
        np.testing.assert_allclose(inter[:, :D], st["interact"][:, :D], rtol=fc_rtol, atol=fc_atol)
    elif spec.model_type in ("wnd", "mtwnd"):
        assert np.array_equal(inter, st["interact"]), "wnd cat must be exact"
    else:
        np.testing.assert_allclose(inter, st["interact"], rtol=fc_rtol, atol=fc_atol)
    if spec.model_type == "dlrm":
        for k, ref in enumerate(st["bot"]):
            np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_BOT, k), ref, rtol=fc_rtol,
                                       atol=fc_atol)
    for k, ref in enumerate(st["top"]):
        np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_TOP, k), ref, rtol=fc_rtol,
                                   atol=fc_atol)
    for tk, head in enumerate(st.get("tasks", [])):
        for k, ref in enumerate(head):
            np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_TASK, tk * 256 + k), ref,
                                       rtol=fc_rtol, atol=fc_atol)
    out = eng.fetch_stage(_lib.STAGE_OUT)
    np.testing.assert_allclose(out, st["out"], rtol=fc_rtol, atol=fc_atol)


@pytest.mark.parametrize("name", sorted(SMALL))
def test_model_stage_parity_host_path(name):
    mb = 16 if name == "dlrm_rm3_like" else 48
    c = make_case(name, num_batches=2, max_batch=mb)
    with _engine(c) as eng:
        for batch_id, B in [(0, mb), (1, 5), (0, 1)]:
            dense, ids, lengths = request(c, batch_id, B)
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            out, tm = eng.infer(dense, ids, lengths, B)
            assert out.shape == (B, c.spec.out_dim)
            np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=1e-5)
            _check_stages(c, eng, st, B)
            assert tm.gpu_ms > 0
        assert eng.launch_count() > 0


@pytest.mark.parametrize("name", ["dlrm_cat", "dlrm_dot", "wnd", "ncf", "din_many"])
def test_staged_path_matches_host_path_bitwise(name):
    c = make_case(name, num_batches=2, max_batch=40)
    with _engine(c) as eng:
        for j in range(2):
            dense, ids, lengths = request(c, j, 40)
            eng.stage_batch(j, dense, [i.astype(np.int64) for i in ids], lengths)   # int64 as np.array(list)
        for batch_id, B in [(1, 40), (0, 17), (1, 1)]:
            dense, ids, lengths = request(c, batch_id, B)
            out_h, _ = eng.infer(dense, ids, lengths, B)
            out_s, tm = eng.infer_staged(batch_id, B)
            assert np.array_equal(out_h, out_s)
            assert tm.h2d_ms == 0.0
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            _check_stages(c, eng, st, B)


@pytest.mark.parametrize("D", [2, 4, 6, 8, 16, 32, 64, 128, 12, 200])
def test_sls_operator_bitexact_ragged(D):
    rng = np.random.RandomState(D)
    rows = 3001
    W = rng.uniform(-1, 1, size=(rows, D)).astype(np.float32)
    # ragged segments incl. empty ones and a long one
    lengths = rng.randint(0, 40, size=97).astype(np.int32)
    lengths[3] = 0
    lengths[50] = 333
    idx = rng.randint(0, rows, size=int(lengths.sum())).astype(np.int64)
    argv = ["--arch_embedding_size", str(rows), "--arch_sparse_feature_size", str(D),
            "--arch_mlp_bot", "4-%d" % D, "--arch_mlp_top", "4-1", "--arch_interaction_op", "cat"]
    from deeprecsys_b200.utils.utils import cli
    with AccelEngine(args=cli(argv), max_batch=8) as eng:
        eng.load_table(0, W)
        ref = O.sparse_lengths_sum(W, idx, lengths)
        got = eng.sparse_lengths_sum(0, idx, lengths)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        got32 = eng.sparse_lengths_sum(0, idx.astype(np.int32), lengths)
        assert np.array_equal(got32.view(np.uint32), ref.view(np.uint32))
        # all-empty and single-segment edge cases
        z = eng.sparse_lengths_sum(0, np.zeros(0, np.int64), np.zeros(5, np.int32))
        assert np.array_equal(z, np.zeros((5, D), np.float32))
        one = eng.sparse_lengths_sum(0, np.array([7], np.int64), np.array([1], np.int32))
        assert np.array_equal(one, W[7:8])


def test_sls_enforce_errors():
    from deeprecsys_b200.utils.utils import cli
    argv = ["--arch_embedding_size", "100", "--arch_sparse_feature_size", "32",
            "--arch_mlp_bot", "4-32", "--arch_mlp_top", "4-1", "--arch_interaction_op", "cat"]
    W = np.ones((100, 32), np.float32)
    with AccelEngine(args=cli(argv), max_batch=8) as eng:
        eng.load_table(0, W)
        with pytest.raises(_lib.DrsError) as ei:          # index out of range
            eng.sparse_lengths_sum(0, np.array([1, 100], np.int64), np.array([2], np.int32))
        assert ei.value.code == _lib.ERR_INDEX
        with pytest.raises(_lib.DrsError) as ei:          # sum(lengths) != len(indices)
            eng.sparse_lengths_sum(0, np.array([1, 2, 3], np.int64), np.array([2], np.int32))
        assert ei.value.code == _lib.ERR_INDEX
        # engine still usable afterwards
        ok = eng.sparse_lengths_sum(0, np.array([1, 2], np.int64), np.array([2], np.int32))
        assert np.array_equal(ok, 2 * np.ones((1, 32), np.float32))
        with pytest.raises(_lib.DrsError) as ei:          # weights not loaded yet
            eng.infer(np.zeros((1, 4), np.float32), [np.array([0], np.int32)],
                      [np.array([1], np.int32)], 1)
        assert ei.value.code == _lib.ERR_STATE


def test_offsets_bitexact_variable_lengths():
    """lengths -> offsets on the GPU vs numpy cumsum, ragged lengths, B > one scan chunk."""
    from deeprecsys_b200.utils.utils import cli
    rng = np.random.RandomState(7)
    T, D, B = 3, 32, 1000
    argv = ["--arch_embedding_size", "-".join(["500"] * T), "--arch_sparse_feature_size", str(D),
            "--arch_mlp_bot", "8-32", "--arch_mlp_top", "16-1", "--arch_interaction_op", "cat",
            "--num_indices_per_lookup", "12"]
    args = cli(argv)
    spec_tables = [rng.uniform(-1, 1, (500, D)).astype(np.float32) for _ in range(T)]
    from deeprecsys_b200.spec import ModelSpec, draw_all_parameters
    spec = ModelSpec(args)
    _, fcs = draw_all_parameters(spec)
    lengths = [rng.randint(0, 13, size=B).astype(np.int32) for _ in range(T)]
    ids = [rng.randint(0, 500, size=int(l.sum())).astype(np.int32) for l in lengths]
    dense = rng.rand(B, 8).astype(np.float32)
    with AccelEngine(args=args, max_batch=B) as eng:
        eng.load_parameters(spec_tables, fcs)
        out, _ = eng.infer(dense, ids, lengths, B)
        off = eng.fetch_stage(_lib.STAGE_OFFSETS)
        for t in range(T):
            assert np.array_equal(off[t], O.lengths_to_offsets(lengths[t]))
        st = O.forward(spec, spec_tables, fcs, dense, ids, lengths)
        pooled = eng.fetch_stage(_lib.STAGE_POOLED)
        assert np.array_equal(pooled.view(np.uint32), st["pooled"].view(np.uint32))
        np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------- tensor-core MLP path
@pytest.mark.parametrize("chain", [True, False], ids=["chain", "per_layer"])
@pytest.mark.parametrize("name", ["dlrm_cat", "dlrm_dot", "dlrm_rm2_like", "dlrm_rm3_like", "wnd",
                                  "ncf", "din", "din_many", "dlrm_odd"])
def test_tf32x3_tensor_core_path_fp32_class_accuracy(name, chain):
    """tcgen05 kind::tf32 with hi/lo split operands (3 MMAs per product).  Final
    outputs: same tolerance as the fp32 FFMA path (atol 1e-5).  Hidden layers:
    atol 5e-5 / rtol 2e-4 — the tensor core accumulates in fp32 with
    round-toward-zero, which leaves a bias of ~K/8*3 half-ulps (2e-5 at K=2560),
    still ~50x tighter than single-pass TF32."""
    mb = 16 if name == "dlrm_rm3_like" else 200
    c = make_case(name, num_batches=1, max_batch=mb)
    with _engine(c, precision="tf32x3", chain=chain) as eng:
        for B in (mb, 3):
            dense, ids, lengths = request(c, 0, B)
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            out, _ = eng.infer(dense, ids, lengths, B)
            np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=1e-5)
            _check_stages(c, eng, st, B, fc_atol=5e-5, fc_rtol=2e-4)


@pytest.mark.parametrize("name", ["dlrm_cat", "wnd", "dlrm_rm3_like"])
def test_tf32_single_pass_throughput_mode(name):
    """single-pass TF32 (10-bit mantissa operands): atol 2e-3 on post-sigmoid outputs."""
    mb = 16 if name == "dlrm_rm3_like" else 200
    c = make_case(name, num_batches=1, max_batch=mb)
    with _engine(c, precision="tf32") as eng:
        dense, ids, lengths = request(c, 0, mb)
        st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
        out, _ = eng.infer(dense, ids, lengths, mb)
        np.testing.assert_allclose(out, st["out"], rtol=0, atol=2e-3)
        pooled = eng.fetch_stage(_lib.STAGE_POOLED)
        assert np.array_equal(pooled.view(np.uint32), st["pooled"].view(np.uint32))


@pytest.mark.parametrize("precision,atol", [("tf32x3", 2e-5), ("tf32", 5e-3)])
@pytest.mark.parametrize("name,ctas", [("wnd", 5), ("wnd", 148), ("dlrm_rm3_like", 3)])
def test_persistent_tensor_core_fc_many_tiles_per_cta(name, ctas, precision, atol):
    """k_fc_tc_persist (two TMEM accumulators, ring running across tile boundaries): forced on,
    with few CTAs so that each one walks many tiles and both accumulator phases wrap several
    times; outputs and the wide layers' activations against the oracle."""
    B = 2048
    c = make_case(name, num_batches=1, max_batch=B)
    with AccelEngine(args=c.args, max_batch=B, precision=precision) as eng:
        eng.set_option("tc_persistent", ctas)
        eng.set_option("mlp_chain", 0)                 # every layer through the FC kernel
        eng.load_parameters(c.tables, c.fcs)
        dense, ids, lengths = request(c, 0, B)
        out, _ = eng.infer(dense, ids, lengths, B)
        st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
        scale = max(1.0, float(np.abs(st["out"]).max()))
        np.testing.assert_allclose(out, st["out"], rtol=0, atol=atol * scale)
        for k, want in enumerate(st["top"][:-1]):
            got = eng.fetch_stage(_lib.STAGE_TOP, k)
            np.testing.assert_allclose(got, want, rtol=0,
                                       atol=(5e-5 if precision == "tf32x3" else 2e-2)
                                       * max(1.0, float(np.abs(want).max())))
        eng.set_option("tc_persistent", 0)             # one CTA per tile: same numbers
        out0, _ = eng.infer(dense, ids, lengths, B)
        assert np.array_equal(out.view(np.uint32), out0.view(np.uint32))


@pytest.mark.parametrize("name", ["din", "din_many", "din254"])
def test_fused_din_attention_equals_the_unfused_graph(name):
    """k_din_attention (attention units + Sum + Concat in one pass over the pooled vectors)
    against the per-layer graph it replaces (k_din_concat -> batched FCs -> k_sum_units ->
    copies) and against the oracle, ragged batch sizes included."""
    c = make_case(name, num_batches=1, max_batch=64)
    with AccelEngine(args=c.args, max_batch=64) as eng:
        eng.load_parameters(c.tables, c.fcs)
        for B in (64, 33, 1):
            dense, ids, lengths = request(c, 0, B)
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            eng.set_option("din_fused", 1)
            out_f, _ = eng.infer(dense, ids, lengths, B)
            att_f = eng.fetch_stage(_lib.STAGE_BOT)
            inter_f = eng.fetch_stage(_lib.STAGE_INTERACT)
            eng.set_option("din_fused", 0)
            out_u, _ = eng.infer(dense, ids, lengths, B)
            att_u = eng.fetch_stage(_lib.STAGE_BOT)
            np.testing.assert_allclose(att_f, st["bot"][0], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(att_f, att_u, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(inter_f, st["interact"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out_f, st["out"], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(out_f, out_u, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("precision,atol", [("tf32x3", 2e-5), ("tf32", 2e-2)])
@pytest.mark.parametrize("shape", ["small", "shipped"])
def test_tensor_core_recurrence_matches_the_oracle_and_the_ffma_kernel(shape, precision, atol):
    """k_rnn_tc (DIEN's two tanh cells on tcgen05: x_t.Wi^T + h_{t-1}.Wg^T into one TMEM
    accumulator, h_t kept on the SM) against the oracle and against the FFMA recurrence it
    replaces; 300 samples = two full 128-row tiles and a ragged one.  "shipped": D = 32, H = 64,
    40 behaviour tables as in dien.json (small tables)."""
    B = 300
    extra = [] if shape == "small" else [
        "--arch_sparse_feature_size", "32", "--hidden_size", "64",
        "--arch_embedding_size", "300-" + "-".join(["200"] * 40) + "-400-500"]
    c = make_case("dien", num_batches=1, max_batch=B, extra=extra)
    with AccelEngine(args=c.args, max_batch=B, precision=precision) as eng:
        eng.load_parameters(c.tables, c.fcs)
        dense, ids, lengths = request(c, 0, B)
        st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
        scale = max(1.0, float(np.abs(st["out"]).max()))
        out_tc, _ = eng.infer(dense, ids, lengths, B)
        h_tc = eng.fetch_stage(_lib.STAGE_BOT)
        eng.set_option("rnn_tc", 0)
        out_ff, _ = eng.infer(dense, ids, lengths, B)
        h_ff = eng.fetch_stage(_lib.STAGE_BOT)
        # hidden state after 2 x N dependent steps: single-pass TF32 products (10-bit mantissa)
        # accumulate ~1e-2 over 80 steps; the 3xTF32 path stays at fp32 level
        h_atol = 5e-5 if precision == "tf32x3" else 3e-2
        np.testing.assert_allclose(h_ff, st["bot"][0], rtol=0, atol=h_atol)
        np.testing.assert_allclose(h_tc, st["bot"][0], rtol=0, atol=h_atol)
        np.testing.assert_allclose(out_tc, st["out"], rtol=0, atol=atol * scale)
        np.testing.assert_allclose(out_tc, out_ff, rtol=0, atol=atol * scale)
        for n in (1, 129):                                  # other tile counts, same engine
            eng.set_option("rnn_tc", 1)
            d2, i2, l2 = request(c, 0, n)
            o2, _ = eng.infer(d2, i2, l2, n)
            want = O.forward(c.spec, c.tables, c.fcs, d2, i2, l2)["out"]
            np.testing.assert_allclose(o2, want, rtol=0, atol=atol * scale)


@pytest.mark.parametrize("name", ["wnd", "ncf", "din", "din254", "mtwnd", "dien"])
def test_fixed_length_gather_equals_the_tiled_gather(name):
    """k_sls_fixed (flat (table, 32 samples) work items, no offsets) against the tiled persistent
    kernel it replaces for fixed-length short segments: pooled vectors bit for bit, staged and
    host path, ragged batch sizes; and a query with ONE different length falls back to the tiled
    kernel and still matches the oracle."""
    c = make_case(name, num_batches=1, max_batch=200)
    with AccelEngine(args=c.args, max_batch=200) as eng:
        eng.load_parameters(c.tables, c.fcs)
        dense, ids, lengths = request(c, 0, 200)
        eng.stage_batch(0, dense, ids, lengths)
        for B in (200, 97, 32, 1):
            got = {}
            for flag in (1, 0):
                eng.set_option("short_fixed", flag)
                out_s, _ = eng.infer_staged(0, B)
                pooled_s = eng.fetch_stage(_lib.STAGE_POOLED)
                d, i, l = request(c, 0, B)
                out_h, _ = eng.infer(d, i, l, B)
                pooled_h = eng.fetch_stage(_lib.STAGE_POOLED)
                assert np.array_equal(pooled_s.view(np.uint32), pooled_h.view(np.uint32))
                got[flag] = (pooled_s, out_s)
            assert np.array_equal(got[1][0].view(np.uint32), got[0][0].view(np.uint32))
            if name != "dien":                   # (DIEN's reshape needs the whole batch)
                st = O.forward(c.spec, c.tables, c.fcs, *request(c, 0, B))
                assert np.array_equal(got[1][0].view(np.uint32), st["pooled"].view(np.uint32))
        if name == "wnd":                        # one shorter segment: not fixed-length any more
            eng.set_option("short_fixed", 1)
            d, i, l = request(c, 0, 64)
            l = [x.copy() for x in l]
            i = [x.copy() for x in i]
            l[3][10] = 0
            i[3] = np.delete(i[3], 10)
            out, _ = eng.infer(d, i, l, 64)
            st = O.forward(c.spec, c.tables, c.fcs, d, i, l)
            assert np.array_equal(eng.fetch_stage(_lib.STAGE_POOLED).view(np.uint32),
                                  st["pooled"].view(np.uint32))
            np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=1e-5)
