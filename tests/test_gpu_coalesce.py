"""drs_submit_many (coalesced serving): several queued requests — request k = the first size_k
samples of staged batch b_k, as inferenceEngine.py:200-215 slices them — served by ONE stacked
launch give exactly what the same requests give one by one, and what the oracle computes."""
import ctypes as C

import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine, pinned_empty
from oracle import numpy_oracle as O
from tests.helpers import make_case, request

pytestmark = pytest.mark.gpu

REQUESTS = [(0, 13), (1, 64), (0, 1), (1, 40), (0, 64), (1, 7)]


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", ["dlrm_cat", "dlrm_rm2_like", "wnd", "ncf", "din_many", "mtwnd",
                                  "dlrm_dot", "dien"])
def test_coalesced_requests_equal_single_requests_and_the_oracle(name, precision):
    c = make_case(name, num_batches=2, max_batch=64)
    total = sum(n for _, n in REQUESTS)
    with AccelEngine(args=c.args, max_batch=256, precision=precision) as eng:
        eng.load_parameters(c.tables, c.fcs)
        for j in range(2):
            dense, ids, lengths = request(c, j, 64)
            eng.stage_batch(j, dense, ids, lengths)
        singles = [eng.infer_staged(b, n)[0] for b, n in REQUESTS]
        buf, _owner = pinned_empty((256, c.spec.out_dim), np.float32)
        buf[:] = np.nan
        for rep in range(3):                                   # lanes are reused
            ticket = eng.submit_many([b for b, _ in REQUESTS], [n for _, n in REQUESTS],
                                     C.c_void_p(buf.ctypes.data))
            eng.wait(ticket)
        got = np.array(buf[:total])
        want = np.concatenate(singles, axis=0)
        if precision == "fp32":
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        else:
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
        stacked = name not in ("dlrm_dot", "dien")             # those run back to back on one lane
        row = 0
        pooled = eng.fetch_stage(_lib.STAGE_POOLED) if stacked else None
        for b, n in REQUESTS:
            dense, ids, lengths = request(c, b, n)
            st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
            np.testing.assert_allclose(got[row:row + n], st["out"], rtol=1e-4, atol=1e-5)
            if stacked:
                assert np.array_equal(pooled[row:row + n].view(np.uint32),
                                      st["pooled"].view(np.uint32)), (name, b, n)
            row += n
        assert np.isnan(buf[total:]).all()                     # nothing written past the stack


def test_coalesced_requests_are_validated():
    c = make_case("dlrm_cat", num_batches=1, max_batch=64)
    with AccelEngine(args=c.args, max_batch=100) as eng:
        eng.load_parameters(c.tables, c.fcs)
        dense, ids, lengths = request(c, 0, 64)
        eng.stage_batch(0, dense, ids, lengths)
        with pytest.raises(_lib.DrsError):
            eng.submit_many([0, 0], [64, 64])                  # 128 stacked rows > max_batch 100
        with pytest.raises(_lib.DrsError):
            eng.submit_many([0, 5], [8, 8])                    # batch 5 was never staged
        with pytest.raises(_lib.DrsError):
            eng.submit_many([0], [65])                         # more samples than the batch holds
        eng.wait(eng.submit_many([0, 0], [50, 50]))            # exactly max_batch is fine
