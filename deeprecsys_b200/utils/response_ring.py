"""Responder helper of the accel engine: builds and sends ServiceResponses in its own process.

Why: with coalesced serving the GPU answers a 256-sample RMC1 query in a few microseconds, and
what is left per packet in the engine's (single-threaded, GIL-bound) loop is mostly the
response: construct ``ServiceResponse`` (/root/reference/utils/packets.py:32-59), pickle it,
write it to the response Queue's pipe — ~10 us, about half of the loop.  The engine therefore
writes one fixed-size binary record per finished request into a single-producer /
single-consumer ring in shared memory (an anonymous mmap created before the fork), and a helper
process turns records into the reference's ServiceResponse objects and writes them to
``responseQueue`` (utils/fastqueue.FrameWriter: same bytes as ``Queue.put``).  The timestamps
are taken by the ENGINE when the result has landed in host memory (``inference_end_time``,
what DeepRecSys.py:105-123 measures latency with); the helper only does the packaging.

The final sentinel travels through the ring too, so every response precedes the ``None`` the
parent counts (DeepRecSys.py:96-98).
"""
from __future__ import absolute_import, division, print_function

import mmap
import os
import time

import numpy as np

REC = np.dtype([("epoch", "<i8"), ("batch_id", "<i8"), ("batch_size", "<i8"), ("sub_id", "<i8"),
                ("total_sub_batches", "<i8"), ("out_batch_size", "<i8"), ("exp_packet", "<i8"),
                ("kinds", "<i8"), ("arrival_time", "<f8"), ("start_time", "<f8"),
                ("end_time", "<f8")])
_INT_FIELDS = ("epoch", "batch_id", "batch_size", "sub_id", "total_sub_batches", "exp_packet")
_KIND_INT, _KIND_BOOL, _KIND_NONE = 0, 1, 2
KIND_SENTINEL = -1                     # `kinds` of the record that ends the stream
_COMMON = _KIND_BOOL << (2 * 5)        # every field an int, exp_packet a bool


_COMMON_TYPES = (int, int, int, int, int, bool, float)


def encode_request(request):
    """-> (values tuple for the int fields, kinds word, arrival_time) or None when a field is
    not an int / bool / None (such a request is answered in-process instead)."""
    try:                                 # what loadGenerator.py sends: one type check for all
        vals = (request.epoch, request.batch_id, request.batch_size, request.sub_id,
                request.total_sub_batches, request.exp_packet, request.arrival_time)
        if tuple(map(type, vals)) == _COMMON_TYPES:
            return vals[:5] + (int(vals[5]),), _COMMON, vals[6]
    except AttributeError:
        pass
    vals, kinds = [], 0
    for i, name in enumerate(_INT_FIELDS):
        v = getattr(request, name, None)
        if v is None:
            k, v = _KIND_NONE, 0
        elif isinstance(v, (bool, np.bool_)):
            k, v = _KIND_BOOL, int(v)
        elif isinstance(v, (int, np.integer)):
            k, v = _KIND_INT, int(v)
        else:
            return None
        kinds |= k << (2 * i)
        vals.append(v)
    at = getattr(request, "arrival_time", None)
    if at is None:
        kinds |= _KIND_NONE << (2 * len(_INT_FIELDS))
        at = 0.0
    elif not isinstance(at, (float, int, np.floating)):
        return None
    return tuple(vals), kinds, float(at)


class ResponseRing(object):
    """SPSC ring of REC records in an anonymous shared mapping (create BEFORE forking)."""

    def __init__(self, capacity=8192):
        self.capacity = int(capacity)
        self._map = mmap.mmap(-1, 128 + self.capacity * REC.itemsize)
        self._hdr = np.frombuffer(self._map, dtype=np.uint64, count=16)     # [0] written, [8] read
        self._recs = np.frombuffer(self._map, dtype=REC, count=self.capacity, offset=128)
        self._w = 0                     # producer's private copy of hdr[0]
        self._r = 0                     # consumer's private copy of hdr[8]

    # ---- producer (engine process)
    def push(self, rows):
        """rows: list of REC-shaped tuples.  Blocks (spins) while the ring is full."""
        n = len(rows)
        while self._w - int(self._hdr[8]) + n > self.capacity:
            time.sleep(0)
        arr = np.array(rows, dtype=REC)
        i = self._w % self.capacity
        first = min(n, self.capacity - i)
        self._recs[i:i + first] = arr[:first]
        if first < n:
            self._recs[:n - first] = arr[first:]
        self._w += n
        self._hdr[0] = self._w          # publish after the records (x86: stores stay in order)

    # ---- consumer (helper process)
    def pop_all(self):
        w = int(self._hdr[0])
        if w == self._r:
            return None
        i, n = self._r % self.capacity, w - self._r
        first = min(n, self.capacity - i)
        out = self._recs[i:i + first].copy()
        if first < n:
            out = np.concatenate([out, self._recs[:n - first]])
        self._r = w
        self._hdr[8] = w
        return out


def responder_main(ring, responseQueue, engine_id, response_cls):
    """Body of the helper process: records -> ServiceResponse -> responseQueue, until the
    sentinel record; then the final ``None``."""
    from deeprecsys_b200.utils.fastqueue import FrameWriter
    try:
        import ctypes
        import signal
        ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, signal.SIGTERM)   # die with the engine
    except Exception:
        pass
    writer = FrameWriter(responseQueue)
    frame = writer.frame
    idle = 0
    while True:
        recs = ring.pop_all()
        if recs is None:
            idle += 1
            if idle > 2000:
                time.sleep(0.0002)      # nothing for a while: stop burning the core
            elif idle > 50:
                time.sleep(0)
            continue
        idle = 0
        frames, done = [], False
        n_int = len(_INT_FIELDS)
        for (epoch, batch_id, batch_size, sub_id, tsb, out_bs, exp, kinds, at, st,
             end) in recs.tolist():
            if kinds == KIND_SENTINEL:
                done = True
                break
            if kinds == _COMMON:        # what loadGenerator.py:165-170 sends: ints + a bool
                exp = bool(exp)
            elif kinds != 0:            # some other field is a bool or None: restore its type
                def dec(v, i):
                    k = (kinds >> (2 * i)) & 3
                    return None if k == _KIND_NONE else (bool(v) if k == _KIND_BOOL else v)
                epoch, batch_id, batch_size = dec(epoch, 0), dec(batch_id, 1), dec(batch_size, 2)
                sub_id, tsb, exp = dec(sub_id, 3), dec(tsb, 4), dec(exp, 5)
                if ((kinds >> (2 * n_int)) & 3) == _KIND_NONE:
                    at = None
            frames.append(frame(response_cls(
                consumer_id=engine_id, epoch=epoch, batch_id=batch_id, batch_size=batch_size,
                arrival_time=at, process_start_time=st, queue_end_time=end,
                inference_end_time=end, out_batch_size=out_bs, total_sub_batches=tsb,
                exp_packet=exp, sub_id=sub_id)))
        if frames:
            if writer.fast:
                writer.write_frames(frames)
            else:
                import pickle
                for f in frames:
                    responseQueue.put(pickle.loads(f[4:]))
        if done:
            writer.put(None)
            if not writer.fast:         # stock put: let the feeder thread drain before leaving
                responseQueue.close()
                responseQueue.join_thread()
            os._exit(0)
