"""Cheaper ends for a stock ``multiprocessing.Queue`` — same bytes on the pipe, same semaphore
accounting, fewer system calls and no feeder thread.

The reference's serving protocol is pickled packets over ``multiprocessing.Queue``s
(/root/reference/DeepRecSys.py:44-66, loadGenerator.py:162-177, accelInferenceEngine.py:46-83).
A B200 needs ~4 us for a 256-sample RMC1 query, while one ``Queue.get()`` costs ~12 us and one
``Queue.put()`` ~10 us plus a hand-off to the feeder thread, so the queue ends decide the
system-level throughput.  What a ``Queue`` puts on its pipe is a sequence of
frames ``struct.pack("!i", len(payload)) + payload`` with ``payload = pickle.dumps(obj)``
(multiprocessing/connection.py ``_send_bytes``; frames below 16 KiB are written with ONE
``write``, atomically for payloads up to PIPE_BUF), and a bounded queue counts free slots in
``Queue._sem`` (acquired by ``put``, released by ``get``).  The helpers here keep exactly that
protocol — a plain ``Queue.get()`` / ``Queue.put()`` on the other side cannot tell the
difference — but move several packets per system call:

* ``FrameReader(q).drain(block)``: ONE ``read`` for everything that is waiting, split into
  frames, unpickled, one ``_sem.release()`` per packet.  Sole-consumer queues only (a partial
  frame is kept in this process between calls).
* ``FrameWriter(q).put_many(objs)``: pickle in the calling thread, one ``write`` per batch under
  the queue's cross-process write lock (several processes may write to the same queue).

Both fall back to ``Queue.get`` / ``Queue.put`` for anything that is not a stock Queue, or when
DRS_ACCEL_QUEUE_FEEDER=1.
"""
from __future__ import absolute_import, division, print_function

import os
import pickle
import select
import struct

_HDR = struct.Struct("!i")
_MAX_ATOMIC = 4096                    # PIPE_BUF: larger batches are split so frames never interleave


def _stock(q):
    try:
        ok = (hasattr(q, "_reader") and hasattr(q, "_writer") and hasattr(q, "_sem")
              and hasattr(q, "_rlock") and getattr(q, "_joincancelled", None) is not None)
    except Exception:
        ok = False
    return ok and os.environ.get("DRS_ACCEL_QUEUE_FEEDER") != "1"


class FrameReader(object):
    """Sole consumer of ``q``: ``drain(block)`` -> list of objects (possibly empty when not
    blocking).  ``waiting()`` -> is at least one byte readable right now."""

    def __init__(self, q, max_bytes=1 << 16):
        self.q = q
        self.fast = _stock(q)
        self.max_bytes = int(max_bytes)
        self._buf = b""
        if self.fast:
            self._fd = q._reader.fileno()
            self._poll = select.poll()
            self._poll.register(self._fd, select.POLLIN)

    def waiting(self):
        if not self.fast:
            return not self.q.empty()
        return bool(self._buf) or bool(self._poll.poll(0))

    def _parse(self, limit):
        buf, pos, out = self._buf, 0, []
        n, size_of, loads = len(buf), _HDR.size, pickle.loads
        release = self.q._sem.release
        while n - pos >= size_of and len(out) < limit:
            (sz,) = _HDR.unpack_from(buf, pos)
            if sz < 0:
                raise ValueError("frames above 2 GiB are not supported on this queue end")
            if n - pos - size_of < sz:
                break
            out.append(loads(buf[pos + size_of:pos + size_of + sz]))
            pos += size_of + sz
            release()                  # Queue.get: one free slot per packet taken
        self._buf = buf[pos:] if pos < n else b""
        return out

    def drain(self, block=True, limit=1 << 30, timeout=None):
        """Everything that is waiting, at most ``limit`` packets; with ``block`` waits for the
        first packet (``timeout`` seconds at most, then returns [])."""
        if not self.fast:
            import queue as _queue
            out = []
            try:
                out.append(self.q.get(block, timeout) if block else self.q.get(False))
                while len(out) < limit:
                    out.append(self.q.get(False))
            except _queue.Empty:
                pass
            return out
        out = self._parse(limit) if self._buf else []
        while not out:
            if not block and not self._poll.poll(0):
                return out
            if block and timeout is not None and not self._poll.poll(int(timeout * 1000)):
                return out
            with self.q._rlock:
                chunk = os.read(self._fd, self.max_bytes)
            if not chunk:
                raise EOFError("queue pipe closed")
            self._buf = self._buf + chunk if self._buf else chunk
            out = self._parse(limit)
            if not block:
                break
        return out


class FrameWriter(object):
    """``put(obj)`` / ``put_many(objs)`` on ``q`` without the feeder thread."""

    def __init__(self, q):
        self.q = q
        self.fast = _stock(q) and getattr(q, "_thread", None) is None   # feeder running: keep order
        if self.fast:
            self._fd = q._writer.fileno()
            self._wlock = q._wlock

    @staticmethod
    def frame(obj):
        data = pickle.dumps(obj)
        return _HDR.pack(len(data)) + data

    def put(self, obj):
        if not self.fast:
            return self.q.put(obj)
        self.write_frames([self.frame(obj)])

    def put_many(self, objs):
        if not self.fast:
            for o in objs:
                self.q.put(o)
            return
        self.write_frames([self.frame(o) for o in objs])

    def write_frames(self, frames):
        """Pre-framed packets -> the pipe, as few ``write`` calls as atomicity allows."""
        acquire, wlock, fd = self.q._sem.acquire, self._wlock, self._fd
        batch, size = [], 0

        def flush():
            data = b"".join(batch)
            if wlock is None:
                _write_all(fd, data)
            else:
                with wlock:
                    _write_all(fd, data)

        for f in frames:
            # Queue.put: takes a free slot.  Never block on a full queue while holding unsent
            # frames — the consumer can only free slots for packets it has received.
            if not acquire(False):
                if batch:
                    flush()
                    batch, size = [], 0
                acquire()
            if batch and size + len(f) > _MAX_ATOMIC:
                flush()
                batch, size = [], 0
            batch.append(f)
            size += len(f)
        if batch:
            flush()


def _write_all(fd, data):
    view = memoryview(data)
    while view:
        n = os.write(fd, view)
        view = view[n:]
