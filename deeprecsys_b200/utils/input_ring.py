"""Shared-memory ring that carries a query's tensors next to its ServiceRequest.

The reference's packets name a pre-generated batch (``batch_id``, ``batch_size``) and carry no
data: every engine process regenerates the whole dataset from the seed
(/root/reference/inferenceEngine.py:72-88, :200-215; utils/packets.py:6-22).  That is enough to
time a sleep-based accelerator model, not to serve queries a client actually sends.  This ring is
the transport for real inputs: the producer (load generator, front end)
writes ``dense [B, m_den] f32``, per-table ``ids int32`` and ``lengths [T, B] int32`` into a slot
of a memory-mapped file under /dev/shm and sends a ``TensorServiceRequest`` that names the slot
(a few dozen bytes on the Queue instead of ~0.8 MB of pickled arrays for a 256-sample RMC1
query); the accel engine maps the same file once, page-locks it (``drs_host_register``) and
hands slot pointers straight to ``drs_submit_host`` — the H2D copies read the producer's bytes,
there is no host-side copy in the engine.

Slot reuse is explicit: the mapping starts with one ``busy`` byte per slot.  The producer
``acquire``s a slot (waits while it is busy, then marks it), writes the tensors and ``put``s the
request; the engine ``release``s the slot once the query that reads it has completed (its H2D
copies are done by then).  Nothing depends on how many requests the queue, the engine lanes,
a coalescing window or several ``--accel_gpus`` consumers may hold at once; ``slots_needed`` is
only a sizing hint that keeps the producer from ever having to wait.

Layout: ``busy uint8 [n_slots]`` (padded to 64 bytes), then the slots.  One slot (all offsets
64-byte aligned), ids of table t at ``t * max_lookups``:
    dense   float32 [max_batch, m_den]
    lengths int32   [T, max_batch]
    ids     int32   [T, max_lookups]
"""
from __future__ import absolute_import, division, print_function

import mmap
import os

import numpy as np

_ALIGN = 64


def _up(n):
    return (int(n) + _ALIGN - 1) // _ALIGN * _ALIGN


def slots_needed(queue_maxsize, depth, consumers=1):
    """Slots so that the producer never waits: what the bounded queue holds + what every consumer
    process may hold in flight (``depth`` each, + 1 being handed over) + 1 being written."""
    return int(queue_maxsize) + int(consumers) * (int(depth) + 1) + 1


class SharedInputRing(object):
    """``SharedInputRing.create(...)`` in the producer, ``SharedInputRing.attach(desc)`` in the
    engine; ``desc`` (a small dict, picklable) travels inside the requests."""

    def __init__(self, path, n_slots, n_tables, m_den, max_batch, max_lookups, owner):
        self.path, self.owner = path, bool(owner)
        self.n_slots, self.T, self.m_den = int(n_slots), int(n_tables), int(m_den)
        self.max_batch, self.max_lookups = int(max_batch), int(max_lookups)
        self.off_dense = 0
        self.off_lengths = _up(self.max_batch * self.m_den * 4)
        self.off_ids = self.off_lengths + _up(self.T * self.max_batch * 4)
        self.slot_bytes = self.off_ids + _up(self.T * self.max_lookups * 4)
        self.header_bytes = _up(self.n_slots)          # busy flags, one byte per slot
        self.nbytes = self.header_bytes + self.slot_bytes * self.n_slots
        flags = os.O_RDWR | (os.O_CREAT | os.O_EXCL if owner else 0)
        fd = os.open(path, flags, 0o600)
        try:
            if owner:
                os.ftruncate(fd, self.nbytes)
            elif os.fstat(fd).st_size < self.nbytes:
                raise ValueError("%s holds %d bytes, the descriptor needs %d"
                                 % (path, os.fstat(fd).st_size, self.nbytes))
            self._map = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self._bytes = np.frombuffer(self._map, dtype=np.uint8)
        self._busy = self._bytes[:self.n_slots]
        if owner:
            self._busy[:] = 0
        self.registered = False

    # ------------------------------------------------------------------ construction
    @classmethod
    def create(cls, n_slots, n_tables, m_den, max_batch, max_lookups, directory="/dev/shm",
               name=None):
        name = name or "drs_ring_%d_%x" % (os.getpid(), int.from_bytes(os.urandom(4), "little"))
        return cls(os.path.join(directory, name), n_slots, n_tables, m_den, max_batch,
                   max_lookups, owner=True)

    @classmethod
    def attach(cls, desc):
        return cls(desc["path"], desc["n_slots"], desc["n_tables"], desc["m_den"],
                   desc["max_batch"], desc["max_lookups"], owner=False)

    def descriptor(self):
        return {"path": self.path, "n_slots": self.n_slots, "n_tables": self.T,
                "m_den": self.m_den, "max_batch": self.max_batch,
                "max_lookups": self.max_lookups}

    # ------------------------------------------------------------------ views
    def base_address(self):
        return self._bytes.ctypes.data

    def _check_slot(self, slot):
        slot = int(slot)
        if not 0 <= slot < self.n_slots:
            raise IndexError("slot %d outside [0, %d)" % (slot, self.n_slots))
        return slot

    def _view(self, slot, off, dtype, shape):
        slot = self._check_slot(slot)
        start = self.header_bytes + slot * self.slot_bytes + off
        count = int(np.prod(shape))
        return self._bytes[start:start + count * 4].view(dtype).reshape(shape)

    def dense(self, slot):
        return self._view(slot, self.off_dense, np.float32, (self.max_batch, self.m_den))

    def lengths(self, slot):
        return self._view(slot, self.off_lengths, np.int32, (self.T, self.max_batch))

    def ids(self, slot):
        return self._view(slot, self.off_ids, np.int32, (self.T, self.max_lookups))

    # ------------------------------------------------------------------ slot ownership
    def acquire(self, slot, timeout_s=30.0):
        """Producer: wait until the engine has released ``slot`` (its previous query completed),
        then mark it busy.  Raises TimeoutError when the slot never comes back."""
        import time
        slot = self._check_slot(slot)
        if self._busy[slot]:
            deadline = time.time() + timeout_s
            while self._busy[slot]:
                if time.time() > deadline:
                    raise TimeoutError("ring slot %d was never released" % slot)
                time.sleep(0)
        self._busy[slot] = 1

    def release(self, slot):
        """Consumer: the query that read ``slot`` is complete; the producer may overwrite it."""
        self._busy[self._check_slot(slot)] = 0

    def is_busy(self, slot):
        return bool(self._busy[self._check_slot(slot)])

    # ------------------------------------------------------------------ producer
    def write(self, slot, dense, ids, lengths, batch_size):
        """Copy one query into ``slot``; ids may be int32 or int64 (stored as int32, the dtype
        the reference's Cast produces, dlrm_s_caffe2.py:308).  Returns the number of ids per
        table.  Raises ValueError when the query does not fit a slot."""
        B = int(batch_size)
        if not 0 <= B <= self.max_batch:
            raise ValueError("batch_size %d outside [0, %d]" % (B, self.max_batch))
        if len(ids) != self.T or len(lengths) != self.T:
            raise ValueError("expected %d tables" % self.T)
        if self.m_den > 0:
            self.dense(slot)[:B] = np.asarray(dense, dtype=np.float32)[:B]
        L, I = self.lengths(slot), self.ids(slot)
        if (isinstance(ids, np.ndarray) and isinstance(lengths, np.ndarray) and ids.ndim == 2
                and lengths.ndim == 2):
            # [T, n] arrays (the generator's layout): one 2-D copy when every table has the
            # same number of lookups
            counts = lengths[:, :B].sum(axis=1)
            n = int(counts[0]) if self.T > 0 else 0
            if (counts == n).all() and n <= self.max_lookups and ids.shape[1] >= n:
                L[:, :B] = lengths[:, :B]
                I[:, :n] = ids[:, :n]
                return [n] * self.T
        counts = []
        for t in range(self.T):
            lt = np.asarray(lengths[t])[:B]
            n = int(lt.sum())
            if n > self.max_lookups:
                raise ValueError("table %d: %d lookups exceed the slot's %d"
                                 % (t, n, self.max_lookups))
            it = np.asarray(ids[t])
            if it.shape[0] < n:
                raise ValueError("table %d: sum(lengths)=%d > len(indices)=%d" % (t, n, it.shape[0]))
            L[t, :B] = lt
            I[t, :n] = it[:n]               # int64 -> int32 narrowing like the reference's Cast
            counts.append(n)
        return counts

    # ------------------------------------------------------------------ consumer
    def read(self, slot, batch_size, counts=None):
        """(dense [B, m_den] | None, ids[t] [n_t], lengths[t] [B]) as views on the slot."""
        B = int(batch_size)
        L, I = self.lengths(slot), self.ids(slot)
        lengths = [L[t, :B] for t in range(self.T)]
        if counts is None:
            counts = [int(l.sum()) for l in lengths]
        ids = [I[t, :int(counts[t])] for t in range(self.T)]
        return (self.dense(slot)[:B] if self.m_den > 0 else None), ids, lengths

    def pointers(self, slot):
        """(dense address | None, [ids address per table], [lengths address per table])."""
        base = self.base_address() + self.header_bytes + self._check_slot(slot) * self.slot_bytes
        return (base + self.off_dense if self.m_den > 0 else None,
                [base + self.off_ids + t * self.max_lookups * 4 for t in range(self.T)],
                [base + self.off_lengths + t * self.max_batch * 4 for t in range(self.T)])

    # ------------------------------------------------------------------ lifetime
    def close(self):
        self._bytes = self._busy = None
        try:
            self._map.close()
        except (BufferError, ValueError):
            pass                              # views still alive: unmapped when they go
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass
