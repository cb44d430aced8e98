"""Locality-aware traces: the LRU stack-distance generator and the trace
profiler against golden fixtures recorded from the reference's own functions
(tests/golden/make_golden.py -> trace_cases.npz), against the reference modules directly when
/root/reference is mounted (authoring container only), through size-independent properties, and
wired into DLRMDataGenerator's ``--data_generation synthetic`` mode."""
import os
import random

import numpy as np
import pytest

from deeprecsys_b200.data_generator import trace_generator as tg
from deeprecsys_b200.data_generator import trace_profile as tp
from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
from deeprecsys_b200.utils.utils import cli

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "trace_cases.npz"))
REF = "/root/reference"
GEN_CASES = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("gen/")})
PROF_CASES = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("prof/")})


@pytest.mark.parametrize("name", GEN_CASES)
def test_trace_generate_lru_matches_reference_golden(name):
    rows, k, pad, seed = (int(v) for v in GOLD["gen/%s/params" % name])
    random.seed(seed)
    np.random.seed(seed)
    got = tg.trace_generate_lru(rows, GOLD["gen/%s/list_sd" % name].tolist(),
                                GOLD["gen/%s/cumm_sd" % name].tolist(), k, bool(pad))
    assert got.dtype == np.uint64
    assert np.array_equal(got, GOLD["gen/%s/trace" % name])
    # both random streams were consumed exactly as the reference consumes them
    assert np.array_equal(np.random.random_sample(2), GOLD["gen/%s/rng_tail" % name])


@pytest.mark.parametrize("name", GEN_CASES)
def test_resumable_generator_equals_one_shot(name):
    rows, k, pad, seed = (int(v) for v in GOLD["gen/%s/params" % name])
    random.seed(seed)
    order = random.sample(range(rows), rows)
    np.random.seed(seed)
    gen = tg.LruTraceGenerator(order, GOLD["gen/%s/list_sd" % name].tolist(),
                               GOLD["gen/%s/cumm_sd" % name].tolist(), bool(pad))
    parts = [gen.generate(n) for n in (1, 0, 7, k // 3, k - k // 3 - 8)]
    assert np.array_equal(np.concatenate(parts), GOLD["gen/%s/trace" % name])


@pytest.mark.parametrize("name", PROF_CASES)
def test_trace_profile_matches_reference_golden(name):
    hi, n, window, seed = (int(v) for v in GOLD["prof/%s/params" % name])
    trace = np.random.RandomState(seed).randint(0, hi, n).astype(np.int32)
    sd, la = tp.trace_profile(trace, window)
    assert sd == GOLD["prof/%s/stack_distances" % name].tolist()
    assert [int(v) for v in la] == GOLD["prof/%s/line_accesses" % name].tolist()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_against_reference_modules_directly():
    import importlib.util

    def load(name):
        spec = importlib.util.spec_from_file_location(
            "ref_" + name, os.path.join(REF, "data_generator", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ref_g, ref_p = load("trace_generator"), load("trace_profile")
    shipped = os.path.join(REF, "data_generator", "profile", "sd_cumm")
    list_sd, cumm_sd = tg.read_dist_from_file(shipped)
    assert (list_sd, cumm_sd) == tuple(ref_g.read_dist_from_file(shipped))
    for rows, k, pad, seed in ((20000, 15000, False, 1), (12000, 15000, True, 2)):
        random.seed(seed)
        np.random.seed(seed)
        want = np.array(ref_g.trace_generate_lru(rows, list_sd, cumm_sd, k, pad), dtype=np.uint64)
        random.seed(seed)
        np.random.seed(seed)
        assert np.array_equal(tg.trace_generate_lru(rows, list_sd, cumm_sd, k, pad), want)
    trace = np.random.RandomState(4).zipf(1.3, 2500) % 300
    want_sd, want_la = ref_p.trace_profile(trace, 64)
    got_sd, got_la = tp.trace_profile(trace, 64)
    assert got_sd == want_sd and [int(v) for v in got_la] == [int(v) for v in want_la]


def test_generate_stack_distance_single_draw_semantics():
    list_sd, cumm = [0, 2, 5], [0.5, 0.8, 1.0]
    np.random.seed(0)
    u = np.random.rand(3)
    np.random.seed(0)
    # i = 0: only distance 0 is possible (support cut to distances <= first references so far)
    assert tg.generate_stack_distance(list_sd, cumm, 5, 0) == 0
    # i >= max: the plain inverse cdf
    want = list_sd[int(np.searchsorted(cumm, u[1], side="left"))]
    assert tg.generate_stack_distance(list_sd, cumm, 5, 9) == want
    # padding: distance 0 is excluded once every distance is reachable
    assert tg.generate_stack_distance(list_sd, cumm, 5, 9, enable_padding=True) in (2, 5)


def test_profile_of_generated_trace_round_trip():
    """profile -> distribution -> generate -> profile: the generated trace only contains stack
    distances of the distribution, and their frequencies follow it."""
    rng = np.random.RandomState(12)
    source = rng.zipf(1.2, 6000) % 500
    sd, _ = tp.trace_profile(source, 200)
    list_sd, prob_sd, cumm_sd = tp.stack_distance_distribution(sd)
    assert list_sd[0] == 0 and abs(cumm_sd[-1] - 1.0) < 1e-9 and abs(sum(prob_sd) - 1.0) < 1e-9
    cumm_sd[-1] = 1.0
    np.random.seed(5)
    trace = tg.LruTraceGenerator(rng.permutation(100000), list_sd, cumm_sd).generate(20000)
    # window = whole trace: the profiler's window counts references, the distances count ids
    sd2, first = tp.trace_profile(trace.astype(np.int64), len(trace))
    assert set(sd2) <= set(list_sd)
    assert len(first) == len(set(trace.tolist()))          # every first reference is a new id
    reuse_src = 1.0 - prob_sd[0]
    reuse_gen = 1.0 - sd2.count(0) / float(len(sd2))
    assert abs(reuse_gen - reuse_src) < 0.03


def test_dist_and_trace_files_round_trip(tmp_path):
    list_sd, cumm_sd = [0, 1, 4, 9], [0.25, 0.5, 0.75, 1.0]
    p2, p3, pt = (str(tmp_path / n) for n in ("two", "three", "trace"))
    tg.write_dist_to_file(p2, list_sd, cumm_sd)
    assert tg.read_dist_from_file(p2) == (list_sd, cumm_sd)
    tg.write_dist_to_file(p3, list_sd, cumm_sd, unique_accesses=[7, 3, 5])
    assert tg.read_dist_from_file(p3) == ([7, 3, 5], list_sd, cumm_sd)
    tg.write_trace_to_file(pt, np.array([5, 1, 5, 9], dtype=np.uint64))
    assert open(pt).read() == "5, 1, 5, 9\n"                # the reference's one-line format
    assert tg.read_trace_from_file(pt).tolist() == [5, 1, 5, 9]
    (tmp_path / "bad").write_text("1, 2\n")
    with pytest.raises(ValueError):
        tg.read_dist_from_file(str(tmp_path / "bad"))


def test_generator_rejects_distances_deeper_than_the_stack():
    np.random.seed(1)
    gen = tg.LruTraceGenerator([3, 1, 2], [0, 50], [0.01, 1.0])
    with pytest.raises(ValueError):
        gen.generate(200)                                   # table of 3 ids, distance 50


def _synthetic_args(tmp_path, extra=()):
    prof = tmp_path / "sd_cumm"
    tg.write_dist_to_file(str(prof), [0, 1, 2, 4, 8, 16], [0.4, 0.55, 0.7, 0.8, 0.9, 1.0])
    return cli(["--arch_embedding_size", "300-200-5000", "--arch_sparse_feature_size", "8",
                "--arch_mlp_bot", "6-8", "--arch_mlp_top", "16-1", "--num_indices_per_lookup", "5",
                "--num_indices_per_lookup_fixed", "True", "--num_batches", "3",
                "--max_mini_batch_size", "16", "--data_generation", "synthetic",
                "--data_trace_file", str(prof)] + list(extra))


def test_synthetic_mode_layout_locality_and_determinism(tmp_path):
    args = _synthetic_args(tmp_path)
    np.random.seed(args.numpy_rand_seed)
    nb, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
    assert nb == 3 and len(lX) == len(lS_l) == len(lS_i) == 3
    for j in range(nb):
        assert lX[j].shape == (16, 6) and lX[j].dtype == np.float32
        assert lS_l[j].shape == (3, 16) and lS_l[j].dtype == np.int32 and (lS_l[j] == 5).all()
        assert lS_i[j].shape == (3, 80) and lS_i[j].dtype == np.int32
        for t, rows in enumerate((300, 200, 5000)):
            assert lS_i[j][t].min() >= 0 and lS_i[j][t].max() < rows
    # reuse: ~60 % of the references of a table hit an id already seen (uniform-random ids over
    # 5000 rows would repeat a handful of times in 240 draws)
    whole = np.concatenate([lS_i[j][2] for j in range(nb)])
    assert len(np.unique(whole)) < 0.55 * whole.size
    # the same seed gives the same dataset; the RNG state afterwards is reproducible too
    tail = np.random.random_sample(2)
    np.random.seed(args.numpy_rand_seed)
    _, lX2, _, lS_i2 = DLRMDataGenerator(args).generate_input_data()
    assert all(np.array_equal(a, b) for a, b in zip(lS_i, lS_i2))
    assert all(np.array_equal(a, b) for a, b in zip(lX, lX2))
    assert np.array_equal(tail, np.random.random_sample(2))


def test_synthetic_mode_per_table_profiles_and_first_reference_order(tmp_path):
    # "j" in the path is replaced by the table index (reference convention); a three-line
    # profile carries the order of first references itself
    for t, order in enumerate(([4, 2, 0, 1, 3], [9, 8, 7, 6, 5, 4, 3, 2, 1, 0])):
        tg.write_dist_to_file(str(tmp_path / ("dist_emb_%d.log" % t)), [0], [1.0],
                              unique_accesses=order)
    args = cli(["--arch_embedding_size", "5-10", "--arch_sparse_feature_size", "4",
                "--arch_mlp_bot", "3-4", "--arch_mlp_top", "8-1", "--num_indices_per_lookup", "2",
                "--num_indices_per_lookup_fixed", "True", "--num_batches", "1",
                "--max_mini_batch_size", "2", "--data_generation", "synthetic",
                "--data_trace_file", str(tmp_path / "dist_emb_j.log")])
    np.random.seed(0)
    _, _, _, lS_i = DLRMDataGenerator(args).generate_input_data()
    assert lS_i[0][0].tolist() == [4, 2, 0, 1] and lS_i[0][1].tolist() == [9, 8, 7, 6]


def test_synthetic_mode_needs_fixed_lookups(tmp_path):
    args = _synthetic_args(tmp_path)
    args.num_indices_per_lookup_fixed = False
    with pytest.raises(SystemExit):
        DLRMDataGenerator(args).generate_input_data()
    args = _synthetic_args(tmp_path)
    args.data_generation = "dataset"
    with pytest.raises(SystemExit):
        DLRMDataGenerator(args).generate_input_data()


def _locality_case(tmp_path, name):
    from tests.helpers import make_case
    prof = tmp_path / "sd_cumm"
    tg.write_dist_to_file(str(prof), [0, 1, 2, 4, 8, 16, 64], [0.3, 0.45, 0.6, 0.7, 0.8, 0.9, 1.0])
    return make_case(name, num_batches=2, max_batch=48,
                     extra=["--data_generation", "synthetic", "--data_trace_file", str(prof),
                            "--num_indices_per_lookup_fixed", "True"])


def test_oracle_runs_on_locality_traces(tmp_path):
    """Lookups of a locality trace repeat ids inside a sample and are not sorted; the oracle's
    sequential SparseLengthsSum takes them as they come (Caffe2 semantics)."""
    from oracle import numpy_oracle as O
    from tests.helpers import request
    c = _locality_case(tmp_path, "dlrm_cat")
    dense, ids, lengths = request(c, 1, 48)
    L = c.spec.L
    groups = ids[0].reshape(48, L)
    assert any(len(set(g.tolist())) < L for g in groups)             # repeats inside a lookup
    st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
    want = np.zeros(c.spec.D, np.float32)
    for i in groups[5]:
        want = want + c.tables[0][i]
    assert np.array_equal(st["pooled"][5, :c.spec.D], want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dlrm_cat", "dlrm_dot"])
def test_gpu_parity_on_locality_traces(tmp_path, name):
    """The CUDA path on a dataset with reuse (repeated, unsorted ids inside a lookup), staged and
    host paths: pooled vectors bit-exact against the oracle, outputs atol 1e-5."""
    from deeprecsys_b200 import _lib
    from deeprecsys_b200.engine import AccelEngine
    from oracle import numpy_oracle as O
    from tests.helpers import request
    c = _locality_case(tmp_path, name)
    with AccelEngine(args=c.args, max_batch=48) as eng:
        eng.load_parameters(c.tables, c.fcs)
        for j in range(2):
            dense, ids, lengths = request(c, j, 48)
            eng.stage_batch(j, dense, ids, lengths)
        for batch_id, B in [(0, 48), (1, 9)]:
            dense, ids, lengths = request(c, batch_id, B)
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            for out in (eng.infer(dense, ids, lengths, B)[0], eng.infer_staged(batch_id, B)[0]):
                pooled = eng.fetch_stage(_lib.STAGE_POOLED)
                assert np.array_equal(pooled.view(np.uint32), st["pooled"].view(np.uint32))
                np.testing.assert_allclose(out, st["out"], rtol=1e-4, atol=1e-5)


def test_sorted_lookups_keep_every_samples_multiset_of_rows(tmp_path):
    """--data_trace_sort_lookups: same ids per sample as the unsorted trace, ascending inside
    each lookup (what np.unique gives the random generator's lookups); RNG use is unchanged."""
    plain = _synthetic_args(tmp_path)
    np.random.seed(plain.numpy_rand_seed)
    _, _, _, a = DLRMDataGenerator(plain).generate_input_data()
    tail = np.random.random_sample(2)
    srt = _synthetic_args(tmp_path, extra=["--data_trace_sort_lookups"])
    np.random.seed(srt.numpy_rand_seed)
    _, _, _, b = DLRMDataGenerator(srt).generate_input_data()
    assert np.array_equal(tail, np.random.random_sample(2))
    for x, y in zip(a, b):
        assert np.array_equal(np.sort(x.reshape(3, 16, 5), axis=2), y.reshape(3, 16, 5))
        assert (np.diff(y.reshape(3, 16, 5), axis=2) >= 0).all()
    assert any(not np.array_equal(x, y) for x, y in zip(a, b))
