"""utils/response_ring.py: the responder helper turns the engine's binary records back into the
reference's ServiceResponse objects — same attribute values and types as an in-process response,
in order, with the final sentinel last."""
import multiprocessing as mp

import numpy as np
import pytest

from deeprecsys_b200.utils import response_ring as R
from deeprecsys_b200.utils.packets import ServiceRequest, ServiceResponse


def _req(i, **kw):
    f = dict(batch_id=i % 5, epoch=i, batch_size=100 + i, sub_id=0, total_sub_batches=1,
             exp_packet=False, arrival_time=1000.0 + i)
    f.update(kw)
    return ServiceRequest(**f)


def test_encode_request_accepts_ints_bools_none_and_refuses_the_rest():
    vals, kinds, at = R.encode_request(_req(3, batch_size=np.int64(77), exp_packet=True))
    assert vals == (3, 3, 77, 0, 1, 1) and at == 1003.0 and kinds != 0
    vals, kinds, at = R.encode_request(_req(4, sub_id=None, arrival_time=None))
    assert vals[3] == 0 and at == 0.0
    assert R.encode_request(_req(5, batch_size="12")) is None
    assert R.encode_request(_req(6, batch_size=12.5)) is None


def test_ring_wraps_and_keeps_order():
    ring = R.ResponseRing(capacity=8)
    seen = []
    for start in range(0, 40, 5):
        rows = [(i, 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0) for i in range(start, start + 5)]
        ring.push(rows)
        seen.extend(int(r["epoch"]) for r in ring.pop_all())
    assert seen == list(range(40)) and ring.pop_all() is None


@pytest.mark.parametrize("n", [1, 300])
def test_responder_process_rebuilds_the_reference_responses(n):
    ctx = mp.get_context("fork")
    resp_q = ctx.Queue()
    ring = R.ResponseRing(capacity=64)                 # smaller than n: the producer must wait
    proc = ctx.Process(target=R.responder_main, args=(ring, resp_q, 9, ServiceResponse))
    proc.start()
    reqs = [_req(i) for i in range(n)]
    reqs[0].exp_packet = True
    if n > 2:
        reqs[2].sub_id, reqs[2].total_sub_batches, reqs[2].arrival_time = None, None, None
    # the parent reads while it produces (as DeepRecSys.py's loop does): 300 responses are more
    # than the response pipe + the ring hold, a producer that only reads afterwards would wait
    # on a full ring for a responder that waits on a full pipe
    import threading
    got = []

    def collect():
        while True:
            item = resp_q.get(timeout=60)
            if item is None:
                break
            got.append(item)

    reader = threading.Thread(target=collect)
    reader.start()
    for i in range(0, n, 7):
        rows = []
        for r in reqs[i:i + 7]:
            vals, kinds, at = R.encode_request(r)
            rows.append(vals[:5] + (int(r.batch_size) - 1, vals[5], kinds, at, 5.0, 6.0))
        ring.push(rows)
    ring.push([(0, 0, 0, 0, 0, 0, 0, R.KIND_SENTINEL, 0.0, 0.0, 0.0)])
    reader.join(timeout=90)
    assert not reader.is_alive()
    proc.join(timeout=30)
    assert not proc.is_alive() and len(got) == n
    for r, q in zip(got, reqs):
        assert isinstance(r, ServiceResponse) and r.consumer_id == 9
        for name in ("epoch", "batch_id", "batch_size", "sub_id", "total_sub_batches",
                     "exp_packet", "arrival_time"):
            a, b = getattr(r, name), getattr(q, name)
            assert a == b and type(a) is type(b), (name, a, b)
        assert r.out_batch_size == q.batch_size - 1
        assert (r.queue_start_time, r.queue_end_time, r.inference_end_time) == (5.0, 6.0, 6.0)
