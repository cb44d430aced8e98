"""GPU parity against golden vectors recorded from the REFERENCE's own model code
(tests/golden/ref_models.npz, see tests/golden/make_ref_models.py): the CUDA
path, called through the C-ABI, must reproduce the blobs the reference's
unmodified builders produce for the same seed and request.

  pooled vectors (emb:::sls*_z) ........ bit-exact
  cat interaction ...................... exact in the embedding columns
  FC outputs, fp32 FFMA path ........... rtol 1e-4 / atol 1e-5
  FC outputs, tf32x3 tcgen05 path ...... atol 1e-5 outputs / 5e-5 hidden
"""
import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine
from tests.helpers import (GOLDEN_MAX_BATCH, SMALL, apply_golden_params, golden_requests,
                           golden_stages, load_ref_golden, make_case, request)

pytestmark = pytest.mark.gpu

Z = load_ref_golden()


def _check(c, eng, ref, B, atol_out, atol_hidden, rtol):
    spec = c.spec
    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
    assert np.array_equal(pooled.view(np.uint32), ref["pooled"].view(np.uint32)), \
        "pooled vectors are not bit-identical to the reference's SparseLengthsSum blobs"
    inter = eng.fetch_stage(_lib.STAGE_INTERACT)
    assert inter.shape == ref["interact"].shape
    if spec.model_type in ("wnd", "mtwnd"):
        assert np.array_equal(inter, ref["interact"])
    elif spec.model_type == "dlrm" and spec.interaction_op == "cat":
        assert np.array_equal(inter[:, spec.D:], ref["interact"][:, spec.D:])
        np.testing.assert_allclose(inter[:, :spec.D], ref["interact"][:, :spec.D], rtol=rtol,
                                   atol=atol_hidden)
    else:
        np.testing.assert_allclose(inter, ref["interact"], rtol=rtol, atol=atol_hidden)
    if spec.model_type == "dien":                 # last hidden state of the second recurrent cell
        np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_BOT), ref["bot"][0], rtol=rtol,
                                   atol=atol_hidden)
    for tk, head in enumerate(ref.get("tasks", [])):
        for k, want in enumerate(head):
            np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_TASK, tk * 256 + k), want,
                                       rtol=rtol, atol=atol_hidden)
    for k, want in enumerate(ref["top"]):
        last = (k == len(ref["top"]) - 1) and spec.model_type not in ("ncf", "mtwnd")
        np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_TOP, k), want, rtol=rtol,
                                   atol=atol_out if last else atol_hidden)
    np.testing.assert_allclose(eng.fetch_stage(_lib.STAGE_OUT), ref["out"], rtol=rtol,
                               atol=atol_out)


@pytest.mark.parametrize("precision,atol_out,atol_hidden", [("fp32", 1e-5, 1e-5),
                                                            ("tf32x3", 1e-5, 5e-5)])
@pytest.mark.parametrize("name", sorted(SMALL))
def test_cuda_path_reproduces_reference_blobs(name, precision, atol_out, atol_hidden):
    c = apply_golden_params(make_case(name, num_batches=2, max_batch=GOLDEN_MAX_BATCH, seed=123), Z)
    with AccelEngine(args=c.args, max_batch=GOLDEN_MAX_BATCH, precision=precision) as eng:
        eng.set_option("tap_hidden", 1)
        eng.load_parameters(c.tables, c.fcs)
        for j in range(2):                                   # staged like the engine process
            dense, ids, lengths = request(c, j, GOLDEN_MAX_BATCH)
            eng.stage_batch(j, dense, [i.astype(np.int64) for i in ids], lengths)
        for batch_id, B in golden_requests(name):
            ref = golden_stages(Z, name, batch_id, B, c.spec)
            # host path (run_queues equivalent)
            dense, ids, lengths = request(c, batch_id, B)
            out, _ = eng.infer(dense, ids, lengths, B)
            np.testing.assert_allclose(out, ref["out"], rtol=1e-4, atol=atol_out)
            _check(c, eng, ref, B, atol_out, atol_hidden, 1e-4)
            # staged path (what accelInferenceEngine serves)
            out_s, _ = eng.infer_staged(batch_id, B)
            assert np.array_equal(out, out_s)
            _check(c, eng, ref, B, atol_out, atol_hidden, 1e-4)
        assert eng.launch_count() > 0
