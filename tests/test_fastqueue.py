"""utils/fastqueue.py: batched ends of a stock multiprocessing.Queue speak the Queue's own
protocol — a plain get()/put() on the other side sees the same objects, and the bounded queue's
slot accounting stays exact."""
import multiprocessing as mp
import queue as _queue

import pytest

from deeprecsys_b200.utils.fastqueue import FrameReader, FrameWriter
from deeprecsys_b200.utils.packets import ServiceRequest


def _req(i):
    return ServiceRequest(batch_id=i % 7, epoch=i, batch_size=100 + i, sub_id=0,
                          total_sub_batches=1, exp_packet=False, arrival_time=float(i))


def _producer(q, n):
    for i in range(n):
        q.put(_req(i))                       # stock put (feeder thread)
    q.put(None)


def test_reader_drains_what_stock_puts_send_and_frees_the_slots():
    ctx = mp.get_context("fork")
    q = ctx.Queue(maxsize=8)                 # bounded: the producer depends on released slots
    p = ctx.Process(target=_producer, args=(q, 200))
    p.start()
    reader = FrameReader(q)
    assert reader.fast
    got = []
    while not got or got[-1] is not None:
        batch = reader.drain(block=True, timeout=10.0)
        assert batch, "producer stalled: slots were not released"
        got.extend(batch)
    p.join(timeout=10)
    assert [r.epoch for r in got[:-1]] == list(range(200))
    assert reader.drain(block=False) == []
    assert not reader.waiting()


def _consumer(q, out_q):
    items = []
    while True:
        r = q.get()                          # stock get
        if r is None:
            break
        items.append((r.epoch, r.batch_size))
    out_q.put(items)


def test_writer_batches_are_plain_queue_traffic_even_on_a_small_bounded_queue():
    ctx = mp.get_context("fork")
    q, out_q = ctx.Queue(maxsize=4), ctx.Queue()
    p = ctx.Process(target=_consumer, args=(q, out_q))
    p.start()
    w = FrameWriter(q)
    assert w.fast
    w.put_many([_req(i) for i in range(50)])          # 50 packets through 4 slots: no deadlock
    w.put(_req(50))
    w.put(None)
    items = out_q.get(timeout=20)
    p.join(timeout=10)
    assert items == [(i, 100 + i) for i in range(51)]
    with pytest.raises(_queue.Empty):
        q.get(timeout=0.05)


def test_limit_and_partial_frames_are_kept_between_calls():
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    w = FrameWriter(q)
    w.put_many([_req(i) for i in range(10)])
    reader = FrameReader(q, max_bytes=300)            # reads end in the middle of a frame
    seen = []
    while len(seen) < 10:
        seen.extend(reader.drain(block=True, limit=3, timeout=5.0))
    assert [r.epoch for r in seen] == list(range(10))
