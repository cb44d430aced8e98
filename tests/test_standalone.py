"""Standalone characterization entry: the reference's ``python models/<model>.py``
main block on the accel backend — six "***" lines that the reference's scrapers parse."""
import numpy as np
import pytest

from deeprecsys_b200.models import standalone
from deeprecsys_b200.utils.utils import cli

SMALL = ["--arch_embedding_size", "500-600-700", "--arch_sparse_feature_size", "16",
         "--arch_mlp_bot", "13-24-16", "--arch_mlp_top", "32-8-1", "--arch_interaction_op", "cat",
         "--num_indices_per_lookup", "6", "--num_batches", "3", "--max_mini_batch_size", "32",
         "--nepochs", "2"]


def parse_like_the_reference(lines):
    """accelerator/predict_execution.py:17-27 (GPU_Data.parse_gpu): every line containing '***'
    yields float(line[rindex('*') + 1 : rindex('ms')]); results are grouped in tuples of six."""
    results = [float(l[l.rindex("*") + 1:l.rindex("ms")]) for l in lines if "***" in l]
    return [results[i:i + 6] for i in range(0, len(results), 6)]


class FakeTiming(object):
    h2d_ms, gpu_ms, d2h_ms, total_ms = 0.25, 1.0, 0.125, 1.5


class FakeEngine(object):
    """Records what the main block does with the engine (no GPU in the CPU tier)."""
    log = []

    def __init__(self, args=None, max_batch=None):
        from deeprecsys_b200.spec import ModelSpec
        self.spec = ModelSpec(args)
        self.max_batch = max_batch
        type(self).log = [("rng_at_init", float(np.random.random_sample()))]

    def create(self):
        self.log.append(("create",))

    def profile_begin(self, n):
        self.log.append(("profile_begin", n))

    def profile_end(self):
        return 1, 0.1, 0.7, 0.2

    def infer(self, dense, ids, lengths, n):
        self.log.append(("infer", dense.shape, len(ids), ids[0].shape[0], lengths[0].shape[0], n))
        return np.zeros((n, 1), np.float32), FakeTiming()

    def close(self):
        self.log.append(("close",))


def test_six_star_lines_in_the_reference_format_and_order():
    lines = []
    res = standalone.run(cli(SMALL + ["--use_accel", "--enable_profiling"]),
                         engine_factory=FakeEngine,
                         out=lambda *a: lines.append(" ".join(str(x) for x in a)))
    starred = [l for l in lines if "***" in l]
    assert [l.split(":")[0] for l in starred] == [
        "Total data loading time", "Total data loading time", "Total computation time",
        "Total computation time", "Total execution time", "Total execution time"]
    assert [l.rstrip().endswith("ms/iter") for l in starred] == [False, True] * 3
    (tup,) = parse_like_the_reference(lines)
    n_iter = 2 * 3
    assert res["iterations"] == n_iter
    assert tup[0] == pytest.approx(0.25 * n_iter) and tup[1] == pytest.approx(0.25)
    assert tup[4] == pytest.approx(res["total_ms"]) and tup[5] == pytest.approx(tup[4] / n_iter)
    assert tup[2] == pytest.approx(tup[4] - tup[0])
    assert res["device_compute_ms"] == pytest.approx(1.125 * n_iter)
    assert sum("ms. per iter" in l for l in lines) == 3          # the profiling split
    # whole batches (max_mini_batch_size samples, L ids each), every epoch, then the engine closes
    infers = [e for e in FakeEngine.log if e[0] == "infer"]
    assert infers == [("infer", (32, 13), 3, 32 * 6, 32, 32)] * n_iter
    assert FakeEngine.log[1] == ("create",) and FakeEngine.log[-1] == ("close",)


def test_rng_is_consumed_like_an_engine_process_before_the_parameters():
    """datagen inputs -> datagen targets -> parameters (inferenceEngine.py:72-88): the engine is
    built when the stream stands exactly where the CPU engines' parameter draws start."""
    from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
    args = cli(SMALL + ["--use_accel"])
    standalone.run(args, engine_factory=FakeEngine, out=lambda *a: None)
    got = dict(FakeEngine.log[:1])["rng_at_init"]
    args = cli(SMALL)                                           # seeds numpy
    gen = DLRMDataGenerator(args)
    gen.generate_input_data()
    gen.generate_output_data()
    assert got == float(np.random.random_sample())


def test_no_cpu_mode():
    with pytest.raises(SystemExit):
        standalone.run(cli(SMALL), engine_factory=FakeEngine, out=lambda *a: None)
