"""GPU: the asynchronous halves of the C-ABI (drs_submit_staged / drs_submit_host / drs_wait,
drs_enqueue_* / drs_sync) — what the serving loop and the bench drive — give the same bits as
the synchronous calls, with several queries in flight on the engine lanes, and report a
SparseLengthsSum enforce failure on the ticket of the offending query only."""
import ctypes as C

import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine, pinned_empty
from tests.helpers import make_case, request

pytestmark = pytest.mark.gpu


def _pinned_query(c, batch_id, n):
    T, L = c.spec.T, c.spec.L
    dense, ids, lengths = request(c, batch_id, n)
    p_ids, k1 = pinned_empty((T, n * L), np.int32)
    p_len, k2 = pinned_empty((T, n), np.int32)
    for t in range(T):
        p_ids[t] = ids[t]
        p_len[t] = lengths[t]
    p_den, k3 = pinned_empty((n, max(1, c.spec.m_den)), np.float32)
    if c.spec.m_den > 0:
        p_den[:] = dense
    keep = (k1, k2, k3, p_ids, p_len, p_den)
    ids_ptrs = (C.c_void_p * T)(*[p_ids[t].ctypes.data for t in range(T)])
    len_ptrs = (C.c_void_p * T)(*[p_len[t].ctypes.data for t in range(T)])
    den_ptr = C.c_void_p(p_den.ctypes.data) if c.spec.m_den > 0 else None
    return den_ptr, ids_ptrs, len_ptrs, keep


@pytest.mark.parametrize("name", ["dlrm_cat", "wnd", "din_many"])
def test_submit_wait_matches_synchronous_calls(name):
    n = 40
    c = make_case(name, num_batches=3, max_batch=n)
    with AccelEngine(args=c.args, max_batch=n, precision="fp32", lanes=3) as eng:
        eng.load_parameters(c.tables, c.fcs)
        want_full, want_part = [], []
        for j in range(3):
            dense, ids, lengths = request(c, j, n)
            eng.stage_batch(j, dense, ids, lengths)
            want_full.append(eng.infer_staged(j, n)[0])
            want_part.append(eng.infer_staged(j, 7 + j)[0])
        assert eng.lane_count() == 3
        # staged submits: three in flight, waited out of order
        outs = [pinned_empty((n, c.spec.out_dim), np.float32) for _ in range(3)]
        tickets = [eng.submit_staged(j, 7 + j, C.c_void_p(outs[j][0].ctypes.data)) for j in range(3)]
        for j in (2, 0, 1):
            eng.wait(tickets[j])
            assert np.array_equal(outs[j][0][:7 + j], want_part[j])
        # host submits from pinned buffers
        queries = [_pinned_query(c, j, n) for j in range(3)]
        tickets = [eng.submit_host(q[0], q[1], 0, q[2], n, C.c_void_p(outs[j][0].ctypes.data))
                   for j, q in enumerate(queries)]
        for j in range(3):
            eng.wait(tickets[j])
            assert np.array_equal(outs[j][0], want_full[j])
        # enqueue / sync pair (bench loop)
        for j in range(3):
            eng.enqueue_host(queries[j][0], queries[j][1], 0, queries[j][2], n,
                             C.c_void_p(outs[j][0].ctypes.data))
        eng.sync()
        for j in range(3):
            assert np.array_equal(outs[j][0], want_full[j])


def test_wait_reports_the_enforce_failure_of_its_own_query_only():
    n = 16
    c = make_case("dlrm_cat", num_batches=2, max_batch=n)
    with AccelEngine(args=c.args, max_batch=n, precision="fp32", lanes=2) as eng:
        eng.load_parameters(c.tables, c.fcs)
        good = _pinned_query(c, 0, n)
        bad = _pinned_query(c, 1, n)
        bad[3][3][0, 5] = c.spec.table_rows[0] + 123          # out-of-range id in table 0
        out_g = pinned_empty((n, c.spec.out_dim), np.float32)
        out_b = pinned_empty((n, c.spec.out_dim), np.float32)
        t_bad = eng.submit_host(bad[0], bad[1], 0, bad[2], n, C.c_void_p(out_b[0].ctypes.data))
        t_good = eng.submit_host(good[0], good[1], 0, good[2], n, C.c_void_p(out_g[0].ctypes.data))
        eng.wait(t_good)                                       # other lane: unaffected
        with pytest.raises(_lib.DrsError) as err:
            eng.wait(t_bad)
        assert err.value.code == _lib.ERR_INDEX
        dense, ids, lengths = request(c, 0, n)                 # and the engine keeps serving
        again, _ = eng.infer(dense, ids, lengths, n)
        assert np.array_equal(again, out_g[0])
